"""Import the UNMODIFIED reference (`trphysx`).

TEST / BASELINE INFRASTRUCTURE.  Two locations are tried, in this order:
  1. /root/reference (the read-only source tree; exists in the build container only): used by
     tests/golden/make_golden.py (fixture generation), tests/test_oracle_vs_reference.py and tests/test_host_api.py;
  2. baseline/_ref (git-ignored `pip install --no-deps --target baseline/_ref` of that tree, made by
     __graft_entry__.build(); it travels to the GPU box with the snapshot): used ONLY by bench.py's reference arm
     (`--impl reference`, `cpu_baseline`), which times the reference's own modules on the host cores.
Nothing under `-m gpu` or smoke() calls this.

`trphysx/__init__.py:5-10` eagerly imports data_utils (h5py) and utils -> viz (matplotlib); neither is
installed and neither is touched by the hot path, so empty stub packages are served for them
(SURVEY.md Appendix A).
"""
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = [os.environ.get("TRPX_REFERENCE_ROOT", "/root/reference"), os.path.join(_REPO, "baseline", "_ref")]
REF_ROOT = next((c for c in _CANDIDATES if os.path.isdir(os.path.join(c, "trphysx"))), _CANDIDATES[0])


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "trphysx"))


def source_tree() -> bool:
    """True when the reference is the full source tree (has examples/: the Rossler module lives there)."""
    return os.path.isdir(os.path.join(REF_ROOT, "examples", "rossler"))


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        v = type(k, (), {"__init__": lambda s, *a, **kw: None, "__call__": lambda s, *a, **kw: s,
                         "__getattr__": lambda s, k2: s})
        setattr(self, k, v)
        return v


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    NAMES = ("h5py", "matplotlib", "mpl_toolkits")

    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in self.NAMES:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Stub(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_loaded = None


def load():
    """Returns the imported `trphysx` package (reference tree is read-only: no bytecode is written)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    sys.dont_write_bytecode = True
    missing = [n for n in _Finder.NAMES if importlib.util.find_spec(n) is None]
    if missing:
        sys.meta_path.insert(0, _Finder())
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    ros = os.path.join(REF_ROOT, "examples", "rossler")
    if ros not in sys.path:
        sys.path.append(ros)
    import trphysx  # noqa
    _loaded = trphysx
    return trphysx


def rossler_modules():
    load()
    from rossler_module.configuration_rossler import RosslerConfig
    from rossler_module.embedding_rossler import RosslerEmbedding, RosslerEmbeddingTrainer
    return RosslerConfig, RosslerEmbedding, RosslerEmbeddingTrainer
