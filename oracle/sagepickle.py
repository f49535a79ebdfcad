"""Let the stock `pickle.load` of main.py:43-44,68-69 decode the reference's two Sage pickles without Sage.

Oracle / test infrastructure only (see oracle/__init__.py).  The fixtures under constraint-system/ are pickled Sage
objects; un-pickling resolves names such as `sage.rings.finite_rings.integer_mod.Mod` by importing their modules.
install() registers an import hook that serves every `sage.<...>` module (except the `sage.all` stand-in) as a stub
whose attributes rebuild the values on kzg_snark_b200/sageshim.py types -- field elements of GF(q), `matrix`, lists
for vectors -- i.e. on the same stand-in algebra the reference's modules already run on here (oracle/refrun.py).
oracle/fixtures.py decodes the same files to plain ints for the JSON copies under tests/golden/.
"""
import importlib.abc
import importlib.machinery
import sys
import types


class _Opaque:
    """Sage parents / categories the values do not need (MatrixSpace, FreeModule, ...)."""

    def __init__(self, name, args=()):
        self.name, self.args, self.state = name, args, None

    def __call__(self, *a, **k):
        return _Opaque(self.name + "()", a)

    def __setstate__(self, st):
        self.state = st


class _MatrixState:
    def __init__(self, cache, data):
        self.cache, self.data = cache, data


def _field_of(args):
    from kzg_snark_b200 import sageshim
    ints = [x for a in args if isinstance(a, tuple) for x in a if isinstance(x, int)]
    return sageshim.GF(max(ints)) if ints and max(ints) > 2 ** 64 else None


def _make_integer(s):
    return int(s, 32)                                  # sage.rings.integer.make_integer: base-32 digits


def _mod(*args):
    """integer_mod.Mod(value, parent): an element of the shim field the parent decoded to."""
    F = next((a for a in args if hasattr(a, "random_element")), None)
    v = next(a for a in args if isinstance(a, int))
    return F(v) if F is not None else v


def _generic_factory_unpickle(factory, *args):
    F = _field_of(args)
    return F if F is not None else _Opaque("factory", args)


def _matrix_unpickle(cls, parent, immutability, cache, data, *version):
    from kzg_snark_b200 import sageshim
    st = _MatrixState(cache if isinstance(cache, dict) else {}, data)
    rows = None
    if isinstance(data, list) and data and isinstance(data[0], list):
        rows = data
    elif isinstance(data, list) and data and int(round(len(data) ** 0.5)) ** 2 == len(data):
        n = int(round(len(data) ** 0.5))
        rows = [data[i * n:(i + 1) * n] for i in range(n)]
    else:
        for key in ("dense_columns", "columns"):
            if key in st.cache:
                cols = [list(c) for c in st.cache[key]]
                rows = [[cols[j][i] for j in range(len(cols))] for i in range(len(cols[0]))]
                break
    if rows is None:
        raise ValueError("cannot recover matrix entries from the pickle state")
    F = rows[0][0].parent()
    return sageshim.matrix(F, rows)


def _make_vec(parent, entries, *rest):
    return list(entries)


_TABLE = {
    "sage.rings.integer.make_integer": _make_integer,
    "sage.rings.finite_rings.integer_mod.Mod": _mod,
    "sage.structure.factory.generic_factory_unpickle": _generic_factory_unpickle,
    "sage.structure.factory.lookup_global": lambda name: _Opaque(name if isinstance(name, str) else name.decode()),
    "sage.matrix.matrix0.unpickle": _matrix_unpickle,
    "sage.modules.free_module_element.make_FreeModuleElement_generic_dense_v1": _make_vec,
}


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        key = f"{self.__name__}.{name}"
        return _TABLE.get(key) or _Opaque(key)


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.startswith("sage.") and fullname != "sage.all":
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_finder = _Finder()


def install():
    if _finder not in sys.meta_path:
        sys.meta_path.insert(0, _finder)
    pkg = sys.modules.get("sage")
    if pkg is not None and not hasattr(pkg, "__path__"):
        pkg.__path__ = []                              # so that `import sage.rings...` treats the stand-in as a package


def uninstall():
    if _finder in sys.meta_path:
        sys.meta_path.remove(_finder)
    for k in [k for k in sys.modules if k.startswith("sage.") and k != "sage.all" and isinstance(sys.modules[k], _StubModule)]:
        del sys.modules[k]
