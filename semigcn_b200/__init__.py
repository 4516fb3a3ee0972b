"""semigcn_b200 -- B200-native (sm_100a) graph-convolution hot path of SeMIGCN.

Public surface:
  * ``semigcn_b200.nn.{GCNConv, ChebConv, Sequential}`` -- drop-ins for the torch_geometric
    symbols the reference imports (util/networks.py:4, util/meshnet.py:6);
  * ``semigcn_b200.data.Data`` -- the attribute bag util/datamaker.py:9,105 needs;
  * ``semigcn_b200.compat`` -- a ``torch_geometric`` shim package: put
    ``semigcn_b200.compat.path()`` on ``sys.path`` (or call ``compat.install()``) and the
    reference's sgcn.py / mgcn.py import our modules unchanged;
  * ``semigcn_b200.networks.SingleScaleGCN`` -- host-side mirror of util/networks.py;
  * ``semigcn_b200.ops`` -- the C-ABI wrappers / autograd Functions;
  * ``include/semigcn_b200.h`` -- the C-ABI itself (libsemigcn_b200.so).
"""
from ._lib import SgbError, LIB_PATH, build_library  # noqa: F401

__all__ = ["SgbError", "LIB_PATH", "build_library"]
__version__ = "0.1.0"
