"""Import shim for the (read-only) Python reference at /root/reference.

TEST INFRASTRUCTURE ONLY.  The reference star-imports optional third-party
packages (optuna, h5py, matplotlib, torchode, schedulefree, psycopg2, ...)
that are not installed in this container.  None of them is touched by the
surrogate compute path except torchode (LatentNeuralODE integration), so we
register inert stand-ins in ``sys.modules`` before importing ``codes``.

Used by ``make_golden.py`` (here, in the build container) to produce the
fixtures under ``tests/golden/*.npz``.  It is never imported on the GPU box:
``/root/reference`` does not exist there.
"""
import importlib.machinery
import sys
import types

REFERENCE_ROOT = "/root/reference"


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
    """Meta-path finder serving stub modules for the missing roots."""

    ROOTS = (
        "optuna", "h5py", "matplotlib", "torchode", "schedulefree", "psycopg2",
        "torchtyping", "seaborn", "sqlalchemy", "mpl_toolkits", "plotly",
    )

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


def import_reference():
    """Return the reference's ``codes.surrogates`` module (live oracle)."""
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import codes.surrogates as surrogates  # noqa: E402
    return surrogates
