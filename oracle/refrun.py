"""Run the reference's OWN modules (kzg.py, fft_ff.py, transcript.py, plonk/*, marlin/*) from
/root/reference, unmodified, inside this image -- and record every call that crosses the
hot-path boundary as a trace.

Oracle / test infrastructure only (see oracle/__init__.py).  /root/reference exists only in the
build container; `__graft_entry__.build()` stages a byte-identical, git-ignored copy under
baseline/_ref/ (oracle/refstage.py) that travels to the GPU box.  This module is used (a) by
tests/golden/make_traces.py to generate the committed fixtures tests/golden/ref_trace_*.json,
(b) by `-m "not gpu"` tests, (c) by the `-m gpu` test that runs the reference's UNMODIFIED
main.py / plonk / marlin on top of the GPU drop-in (`ReferenceRun(gpu_dropin=True)`: `kzg` and
`fft_ff` then resolve to kzg_snark_b200/dropin/, everything else to the reference), and (d) by
`bench.py --impl reference`, which times the reference's own kzg.py commit loop.

What is real and what is a stand-in when the reference runs here:
  real      kzg.py (setup/commit/open/check/batch_check), fft_ff.py, transcript.py,
            plonk/{encoder,indexer,prover,verifier}.py, marlin/{...}.py -- imported from
            /root/reference by file path, byte-for-byte the reference's code;
  stand-in  `sage.all` (GF, PolynomialRing, vector, prod, matrix) -> kzg_snark_b200/sageshim.py,
            `py_ecc.optimized_bn128` -> oracle/pyecc_standin.py, `py_ecc.optimized_bls12_381` -> oracle/pyecc_standin_bls.py.  Both third-party packages are
            un-pinned pip/conda dependencies of the reference that cannot be installed here.

The trace therefore pins the oracle (and the CUDA path) against the reference's own control
flow -- coercions, zero skipping, xi powers, the recursion and ordering of fft_ff -- with exact
modular arithmetic underneath; it does not pin py_ecc's or Sage's internals, which are exact
integer arithmetic with a canonical result.
"""

import importlib
import os
import random
import sys
import types

from . import refstage

REFERENCE_ROOT = refstage.root() or refstage.SOURCE
DROPIN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kzg_snark_b200", "dropin")
_REF_MODULES = ("main", "kzg", "fft_ff", "transcript", "plonk", "plonk.encoder", "plonk.indexer", "plonk.prover",
                "plonk.verifier", "marlin", "marlin.encoder", "marlin.indexer", "marlin.prover", "marlin.verifier")


def available():
    global REFERENCE_ROOT
    REFERENCE_ROOT = refstage.root() or refstage.SOURCE
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "kzg.py"))


def _install_standins():
    from kzg_snark_b200 import sageshim
    from . import pyecc_standin, pyecc_standin_bls

    sage = types.ModuleType("sage")
    sage_all = types.ModuleType("sage.all")
    sage_all.GF = sageshim.GF
    sage_all.PolynomialRing = sageshim.PolynomialRing
    sage_all.vector = sageshim.vector
    sage_all.matrix = sageshim.matrix
    sage_all.prod = sageshim.prod
    sage.all = sage_all
    py_ecc = types.ModuleType("py_ecc")
    py_ecc.optimized_bn128 = pyecc_standin
    py_ecc.optimized_bls12_381 = pyecc_standin_bls
    # py_ecc.fields: the FQ classes a drop-in builds its result points from (kzg_snark_b200/points.py:fq_class)
    fields = types.ModuleType("py_ecc.fields")
    fields.optimized_bn128_FQ = pyecc_standin.FQ
    fields.optimized_bls12_381_FQ = pyecc_standin_bls.FQ
    py_ecc.fields = fields
    names = ("sage", "sage.all", "py_ecc", "py_ecc.optimized_bn128", "py_ecc.optimized_bls12_381", "py_ecc.fields")
    saved = {k: sys.modules.get(k) for k in names}
    sys.modules.update({"sage": sage, "sage.all": sage_all, "py_ecc": py_ecc, "py_ecc.fields": fields,
                        "py_ecc.optimized_bn128": pyecc_standin, "py_ecc.optimized_bls12_381": pyecc_standin_bls})
    from . import sagepickle
    sagepickle.install()                                 # main.py:43-44,68-69 un-pickle Sage objects
    return saved


class ReferenceRun:
    """Context manager: inside it, `self.kzg`, `self.fft_ff`, `self.plonk_*`, `self.marlin_*` are
    the reference's modules; `self.trace` collects boundary calls.  On exit sys.modules and
    sys.path are restored so the repo's own drop-in modules named `kzg` / `fft_ff` are not
    shadowed for other tests."""

    def __init__(self, seed=0, record=True, gpu_dropin=False):
        self.seed = seed
        self.record = record
        self.gpu_dropin = gpu_dropin          # `kzg` / `fft_ff` = the GPU drop-in instead of the reference's
        self.trace = []
        self._depth = 0

    # -- value encoders (JSON-friendly, canonical)
    @staticmethod
    def enc_scalar(x):
        return hex(int(x))

    @staticmethod
    def enc_poly(p):
        if hasattr(p, "list"):
            p = p.list()
        return [hex(int(c)) for c in p]

    @staticmethod
    def enc_point(pt):
        from . import pyecc_standin as E
        if E.is_inf(pt):
            return None
        x, y = E.normalize(pt)
        return [hex(int(x)), hex(int(y))]

    def __enter__(self):
        if not available():
            raise RuntimeError("reference tree not present")
        self._saved_std = _install_standins()
        self._saved_ref = {k: sys.modules.pop(k, None) for k in _REF_MODULES}
        self._saved_path = list(sys.path)
        self._saved_cwd = os.getcwd()
        sys.path.insert(0, REFERENCE_ROOT)
        hot_root = REFERENCE_ROOT
        if self.gpu_dropin:                       # the zero-edit substitution: the drop-in directory precedes the reference
            sys.path.insert(0, DROPIN_DIR)
            hot_root = DROPIN_DIR
            # the drop-in builds its points from py_ecc's FQ class when py_ecc is importable -- re-resolve it now that
            # the stand-in is installed
            importlib.invalidate_caches()
        from kzg_snark_b200 import sageshim
        sageshim.seed(self.seed)
        random.seed(self.seed)

        self.fft_ff = importlib.import_module("fft_ff")
        assert self.fft_ff.__file__.startswith(hot_root), self.fft_ff.__file__
        if self.record:
            self._wrap_fft()
        self.kzg = importlib.import_module("kzg")
        assert self.kzg.__file__.startswith(hot_root), self.kzg.__file__
        if self.record:
            self._wrap_kzg()
        self.transcript = importlib.import_module("transcript")
        assert self.transcript.__file__.startswith(REFERENCE_ROOT)
        return self

    def main(self):
        """The reference's main.py, imported unmodified; its demos open the fixtures with cwd-relative paths
        (main.py:43,68), so the working directory is the reference root until the context exits."""
        os.chdir(REFERENCE_ROOT)
        return self.load("main")

    def load(self, name):
        m = importlib.import_module(name)
        assert m.__file__.startswith(REFERENCE_ROOT), m.__file__
        return m

    def __exit__(self, *exc):
        for k in _REF_MODULES:
            sys.modules.pop(k, None)
        for k, v in self._saved_ref.items():
            if v is not None:
                sys.modules[k] = v
        for k, v in self._saved_std.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        sys.path[:] = self._saved_path
        os.chdir(self._saved_cwd)
        from . import sagepickle
        sagepickle.uninstall()
        return False

    # -- tracing wrappers: only outermost boundary calls are recorded (fft_ff_interpolation calls
    #    ifft_ff calls fft_ff; KZG.open calls KZG.commit)
    def _wrap_fft(self):
        mod, run = self.fft_ff, self
        orig_fft, orig_ifft, orig_interp = mod.fft_ff, mod.ifft_ff, mod.fft_ff_interpolation

        def rec(fn, vals, w, out):
            run.trace.append({"fn": fn, "n": len(vals), "w": run.enc_scalar(w),
                              "in": [run.enc_scalar(v) for v in vals], "out": run.enc_poly(out)})

        def fft_ff(coeffs, w, F):
            top = run._depth == 0
            run._depth += 1
            try:
                out = orig_fft(coeffs, w, F)
            finally:
                run._depth -= 1
            if top:
                rec("fft_ff", coeffs, w, out)
            return out

        def ifft_ff(values, w, F):
            top = run._depth == 0
            run._depth += 1
            try:
                out = orig_ifft(values, w, F)
            finally:
                run._depth -= 1
            if top:
                rec("ifft_ff", values, w, out)
            return out

        def fft_ff_interpolation(values, g, F):
            top = run._depth == 0
            run._depth += 1
            try:
                out = orig_interp(values, g, F)
            finally:
                run._depth -= 1
            if top:
                rec("fft_ff_interpolation", values, g, out)
            return out

        # fft_ff recurses through its module global: nested calls see depth > 0 and pass through
        mod.fft_ff, mod.ifft_ff, mod.fft_ff_interpolation = fft_ff, ifft_ff, fft_ff_interpolation

    def _wrap_kzg(self):
        KZG, run = self.kzg.KZG, self
        orig_commit, orig_open, orig_setup = KZG.commit, KZG.open, KZG.setup

        def setup(kself, max_degree, *a, **kw):
            run._depth += 1
            try:
                ck, rk = orig_setup(kself, max_degree, *a, **kw)
            finally:
                run._depth -= 1
            return ck, rk

        def commit(kself, ck, polynomials):
            top = run._depth == 0
            run._depth += 1
            try:
                out = orig_commit(kself, ck, polynomials)
            finally:
                run._depth -= 1
            if top:
                run.trace.append({"fn": "commit", "ck_len": len(ck), "ck_id": run._ck_id(ck),
                                  "polys": [run.enc_poly(p) for p in polynomials],
                                  "out": [run.enc_point(c) for c in out]})
            return out

        def open_(kself, ck, polynomials, z, xi):
            top = run._depth == 0
            run._depth += 1
            try:
                out = orig_open(kself, ck, polynomials, z, xi)
            finally:
                run._depth -= 1
            if top:
                run.trace.append({"fn": "open", "ck_len": len(ck), "ck_id": run._ck_id(ck),
                                  "polys": [run.enc_poly(p) for p in polynomials],
                                  "z": run.enc_scalar(z), "xi": run.enc_scalar(xi), "out": run.enc_point(out)})
            return out

        KZG.setup, KZG.commit, KZG.open = setup, commit, open_

    def _ck_id(self, ck):
        """Commitment keys are recorded once (affine) in self.keys and referred to by index."""
        if not hasattr(self, "keys"):
            self.keys, self._key_ids = [], {}
        k = id(ck)
        if k not in self._key_ids:
            enc = [self.enc_point(p) for p in ck]
            for i, e in enumerate(self.keys):
                if e == enc:
                    self._key_ids[k] = i
                    break
            else:
                self._key_ids[k] = len(self.keys)
                self.keys.append(enc)
            self._keep = getattr(self, "_keep", []) + [ck]      # keep ids alive
        return self._key_ids[k]
