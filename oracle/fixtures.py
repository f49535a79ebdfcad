"""Sage-free loader for the reference's two pickled fixtures.

Oracle / test infrastructure only (see oracle/__init__.py).

constraint-system/PLONK_ARITHMETIZATION_INSTANCE.pkl and R1CS_INSTANCE.pkl are
pickled Sage objects (main.py:43-44,68-69 load them with Sage present).  A stub
`Unpickler.find_class` maps the Sage constructors to plain Python values
(SURVEY.md section 8c).  tests/golden/make_fixtures.py uses this to write the
JSON copies the tests read (the pickles themselves stay in /root/reference).
"""

import io
import pickle


class _Field:
    def __init__(self, q):
        self.q = q


class _Opaque:
    """Placeholder for Sage parents we do not need (MatrixSpace, FreeModule ...)."""

    def __init__(self, name, args=()):
        self.name = name
        self.args = args
        self.state = None

    def __call__(self, *a, **k):
        return _Opaque(self.name + "()", a)

    def __setstate__(self, st):
        self.state = st


class _Matrix:
    def __init__(self):
        self.parent = None
        self.rows = None

    def __setstate__(self, st):
        self.state = st


def _make_integer(s):
    return int(s, 32)                    # sage.rings.integer.make_integer: base-32 string


def _mod(*args):
    # sage.rings.finite_rings.integer_mod.Mod(value, modulus-or-parent, ...): keep the residue
    for a in args:
        if isinstance(a, int):
            return a
    raise ValueError("Mod() pickle without an integer value")


def _generic_factory_unpickle(factory, *args):
    # FiniteField factory: args carry (order, name, modulus, impl, ...) in a key tuple
    ints = [x for a in args if isinstance(a, tuple) for x in a if isinstance(x, int)]
    if ints and max(ints) > 2 ** 64:
        return _Field(max(ints))
    return _Opaque("factory", args)


def _lookup_global(name):
    return _Opaque(name if isinstance(name, str) else name.decode())


def _matrix_unpickle(cls, parent, immutability, cache, data, *version):
    m = _Matrix()
    m.parent, m.cache, m.data = parent, cache, data
    return m


def _make_vec(parent, entries, *rest):
    return list(entries)


class _StubUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        key = f"{module}.{name}"
        table = {
            "sage.rings.integer.make_integer": _make_integer,
            "sage.rings.finite_rings.integer_mod.Mod": _mod,
            "sage.structure.factory.generic_factory_unpickle": _generic_factory_unpickle,
            "sage.structure.factory.lookup_global": _lookup_global,
            "sage.matrix.matrix0.unpickle": _matrix_unpickle,
            "sage.modules.free_module_element.make_FreeModuleElement_generic_dense_v1": _make_vec,
        }
        if key in table:
            return table[key]
        if module.startswith("sage."):
            return _Opaque(key)
        return super().find_class(module, name)


def _matrix_to_rows(m):
    """Dense row list from whatever the Matrix_generic_dense pickle carried."""
    data = m.data
    if isinstance(data, list) and data and isinstance(data[0], list):
        return data
    cache = m.cache if isinstance(m.cache, dict) else {}
    if isinstance(data, list) and data:
        n = int(round(len(data) ** 0.5))
        if n * n == len(data):
            return [data[i * n:(i + 1) * n] for i in range(n)]
    for key in ("dense_columns", "columns"):
        if key in cache:
            cols = [list(c) for c in cache[key]]
            return [[cols[j][i] for j in range(len(cols))] for i in range(len(cols[0]))]
    raise ValueError("cannot recover matrix entries from pickle state")


def load_pickle(path):
    with open(path, "rb") as f:
        return _StubUnpickler(io.BytesIO(f.read())).load()


def load_plonk_instance(path):
    """dict with qM,qL,qR,qO,qC (len n), perm (len 3n), w (len 3n) as ints (main.py:68-79)."""
    d = load_pickle(path)
    return {k: [int(x) for x in v] for k, v in d.items()}


def load_r1cs_instance(path):
    """dict with A,B,C as dense row lists and z as list (main.py:43-48)."""
    d = load_pickle(path)
    out = {}
    for k, v in d.items():
        if isinstance(v, _Matrix):
            out[k] = [[int(x) for x in row] for row in _matrix_to_rows(v)]
        else:
            out[k] = [int(x) for x in v]
    return out
