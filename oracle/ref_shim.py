"""Import shim for the Python reference (TEST / BASELINE INFRASTRUCTURE ONLY).

The reference (robin-janssen/CODES-Benchmark) star-imports optional third-party packages
(optuna, h5py, matplotlib, torchode, schedulefree, psycopg2, ...) that are not installed in
this image.  None of them is touched by the surrogate compute path except torchode
(LatentNeuralODE integration), so inert stand-ins are registered in ``sys.modules`` before
``codes`` is imported.

Where the reference is found, in this order:
  1. ``/root/reference``           -- the read-only checkout in the build container
  2. ``oracle/_ref``               -- the git-ignored vendored copy made by ``oracle/make_ref.py``
                                      (travels to the GPU box with the snapshot, like a built .so)
Users: ``tests/golden/make_golden.py`` (fixtures), ``bench.py --impl reference`` and the
``reference_gpu`` comparator leg of ``bench.py``.  Never imported by the product package.
"""
import importlib.machinery
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATE_ROOTS = ("/root/reference", os.path.join(HERE, "_ref"))


class _Anything:
    """Callable/attribute sink used for every symbol of a stubbed package."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _StubFinder:
    """Meta-path finder serving stub modules for the missing roots (only those that do not import)."""

    ROOTS = ("optuna", "h5py", "matplotlib", "torchode", "schedulefree", "psycopg2",
             "torchtyping", "seaborn", "sqlalchemy", "mpl_toolkits", "plotly")

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        if module.__name__ == "optuna":
            class TrialPruned(Exception):
                pass
            module.TrialPruned = TrialPruned


def reference_root():
    """Directory that holds the reference's ``codes`` package, or None."""
    for root in CANDIDATE_ROOTS:
        if os.path.isfile(os.path.join(root, "codes", "__init__.py")):
            return root
    return None


def import_reference(root=None):
    """Return the reference's ``codes.surrogates`` module (the live oracle); raises ImportError if absent."""
    root = root or reference_root()
    if root is None:
        raise ImportError("reference package not found (neither /root/reference nor oracle/_ref; "
                          "run `python oracle/make_ref.py` in the build container)")
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder())     # appended: real packages, when installed, win
    if root not in sys.path:
        sys.path.insert(0, root)
    import codes.surrogates as surrogates  # noqa: E402
    return surrogates
