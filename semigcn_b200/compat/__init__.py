"""``torch_geometric`` shim: makes ``from torch_geometric.nn import GCNConv, ChebConv,
Sequential`` and ``from torch_geometric.data import Data`` (reference util/networks.py:4,
util/meshnet.py:6, util/datamaker.py:9) resolve to semigcn_b200.

    import semigcn_b200.compat as compat; compat.install()     # before importing the reference
or  PYTHONPATH=$(python -c 'import semigcn_b200.compat as c; print(c.path())') python sgcn.py ...
"""
import os
import sys


def path() -> str:
    return os.path.dirname(os.path.abspath(__file__))


def install(force: bool = False) -> None:
    """Put the shim package first on sys.path (refuses to shadow a real torch_geometric unless forced)."""
    if not force:
        try:
            import importlib.util
            spec = importlib.util.find_spec("torch_geometric")
            if spec is not None and path() not in (spec.origin or ""):
                raise RuntimeError("a real torch_geometric is installed; pass force=True to shadow it")
        except (ImportError, ValueError):
            pass
    if path() not in sys.path:
        sys.path.insert(0, path())
