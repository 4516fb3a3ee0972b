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


def patch_meshnet(meshnet_module=None) -> None:
    """Replace the reference's ``MeshPool`` / ``MeshUnpool`` (util/meshnet.py:9-27: torch.sparse.mm plus a dense
    [n_coarse, n_fine] row-sum on every forward) by the SpMM-kernel drop-ins of the same constructor / forward:

        import util.meshnet as meshnet; compat.patch_meshnet(meshnet)      # before MGCN(...) is constructed

    With no argument, patches ``util.meshnet`` if it is already imported."""
    from semigcn_b200.nn import MeshPool, MeshUnpool
    if meshnet_module is None:
        meshnet_module = sys.modules.get("util.meshnet")
        if meshnet_module is None:
            raise RuntimeError("patch_meshnet: import util.meshnet first or pass the module")
    meshnet_module.MeshPool, meshnet_module.MeshUnpool = MeshPool, MeshUnpool
