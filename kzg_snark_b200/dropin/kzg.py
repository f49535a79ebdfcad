"""`from kzg import KZG` (main.py:7, plonk/*.py:1, marlin/*.py:3) resolves here when this
directory precedes the reference checkout on sys.path.  See INTEGRATION.md."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from kzg_snark_b200.kzg import KZG, CommitmentKey  # noqa: E402,F401
