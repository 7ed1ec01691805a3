from . import regnet  # noqa: F401
