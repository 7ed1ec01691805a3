"""TEST INFRASTRUCTURE ONLY — minimal stand-in for `timm==1.0.3` (requirements.txt:39 of the
reference), which is neither vendored under /root/reference nor installable offline.

It restates the published RegNetY-200MF / RegNetY-800MF definition (pycls / timm `regnet.py`)
with timm's attribute names, so that the reference's `model/model.py:38-46` and
`model/shift.py:46,72-73` import and run unchanged in this container.  It is used solely by
`oracle/gen_golden.py` and by tests that pin `oracle/tdeed_oracle.py` against the reference.
Nothing in the product path (t-deed_b200/) may import it.

Parity status: the shim is pinned only by parameter counts (regnety_002: 3 162 996,
regnety_008: 6 263 168 including the 1000-way fc, as published by timm) — "parity unpinned"
with respect to timm's own numerics, since timm's source is absent.
"""
from .models.regnet import create_model  # noqa: F401
from . import models, layers  # noqa: F401
