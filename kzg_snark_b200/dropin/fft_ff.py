"""`from fft_ff import fft_ff, fft_ff_interpolation` (plonk/encoder.py:3, plonk/prover.py:2,
marlin/encoder.py:3, marlin/prover.py:4) resolves here when this directory precedes the
reference checkout on sys.path.  See INTEGRATION.md."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from kzg_snark_b200.fft_ff import (  # noqa: E402,F401
    fft_ff, ifft_ff, fft_ff_interpolation, coset_fft_ff, coset_ifft_ff)
