"""tdeed_b200 — B200-native (sm_100a) kernels and host runtime for T-DEED's per-clip hot path."""
from . import _lib  # noqa: F401
from .engine import EngineConfig, InferenceEngine  # noqa: F401

__all__ = ['EngineConfig', 'InferenceEngine']
