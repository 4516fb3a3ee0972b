"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares, the Python modules mirror PyG's constructor / parameter layout, CPU tensors
are refused (no fallback), and the torch_geometric shim resolves to our classes."""
import os
import re
import subprocess
import sys

import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "semigcn_b200.h")).read()
    return sorted(set(re.findall(r"SGB_API[^;(]*?\b(sgb_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from semigcn_b200 import _lib as L
    syms = header_symbols()
    assert len(syms) >= 20
    lib = L.load()
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/semigcn_b200.h but not exported"
        assert s in L.SIGNATURES, f"{s} has no ctypes signature"
    assert set(L.SIGNATURES) == set(syms)
    assert lib.sgb_version() == 100
    out = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (\w+)", out))
    assert set(syms) <= exported


def test_header_compiles_as_c():
    src = '#include "semigcn_b200.h"\nint main(void){return SGB_VERSION==100?0:1;}\n'
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"],
                       input=src, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_size_queries_need_no_gpu():
    from semigcn_b200 import _lib as L
    lib = L.load()
    assert lib.sgb_graph_build_workspace_bytes(6000, 1000) > 6000 * 4
    assert lib.sgb_gemm_stat_rows(1000) == 4 * 8          # one row per epilogue warp and m-tile group, capped at 4 x #SMs
    assert lib.sgb_gemm_stat_rows(10 ** 6) <= 4 * 1024    # no longer grows with m (a GPU-less process assumes 148 SMs)
    assert lib.sgb_gemm_tn_workspace_bytes(100000, 256, 256) >= 256 * 256 * 4


def test_module_layout_matches_pyg():
    from semigcn_b200.nn import ChebConv, GCNConv, Sequential
    from oracle import pyg_ref as O
    torch.manual_seed(314)
    ours = GCNConv(4, 16)
    torch.manual_seed(314)
    ref = O.GCNConv(4, 16)
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys()) == ["bias", "lin.weight"]
    assert torch.equal(ours.lin.weight, ref.lin.weight) and torch.equal(ours.bias, ref.bias)
    torch.manual_seed(314)
    ours = ChebConv(4, 16, K=3)
    torch.manual_seed(314)
    ref = O.ChebConv(4, 16, K=3)
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    for a, b in zip(ours.state_dict().values(), ref.state_dict().values()):
        assert torch.equal(a, b)
    blk = Sequential("x, edge_index", [(GCNConv(4, 8), "x, edge_index -> x"), nn.BatchNorm1d(8), nn.LeakyReLU(),
                                       (nn.Linear(8, 3), "x -> x")])
    keys = list(blk.state_dict().keys())
    assert "module_0.lin.weight" in keys and "module_1.running_var" in keys and "module_3.bias" in keys
    assert blk._fusable(0) == (3, pytest.approx(0.01))


def test_network_state_dict_interchanges_with_oracle():
    from semigcn_b200.networks import SingleScaleGCN
    from oracle import pyg_ref as O
    for conv in ("gcnconv", "chebconv"):
        torch.manual_seed(314)
        ours = SingleScaleGCN("cpu", conv=conv)
        torch.manual_seed(314)
        ref = O.SingleScaleGCN(conv)
        sd_o, sd_r = ours.state_dict(), ref.state_dict()
        assert list(sd_o.keys()) == list(sd_r.keys())
        for k in sd_o:
            assert torch.equal(sd_o[k], sd_r[k]), k       # same seed -> same init (RNG consumption, A.4)
        ours.load_state_dict(sd_r)


def test_cpu_tensors_are_refused():
    from semigcn_b200 import SgbError
    from semigcn_b200.nn import GCNConv
    conv = GCNConv(4, 8)
    with pytest.raises(SgbError):
        conv(torch.randn(5, 4), torch.tensor([[0, 1], [1, 0]]))


def test_unsupported_arguments_raise():
    from semigcn_b200 import SgbError
    from semigcn_b200.nn import ChebConv, GCNConv
    with pytest.raises(SgbError):
        GCNConv(4, 8, improved=True)
    with pytest.raises(SgbError):
        ChebConv(4, 8, K=3, normalization="rw")
    with pytest.raises(SgbError):
        ChebConv(4, 8, K=0)


def test_torch_geometric_shim_resolves_to_our_classes():
    code = ("import semigcn_b200.compat as c; c.install();"
            "from torch_geometric.nn import GCNConv, ChebConv, Sequential;"
            "from torch_geometric.data import Data;"
            "import semigcn_b200.nn as n; assert GCNConv is n.GCNConv and ChebConv is n.ChebConv and Sequential is n.Sequential;"
            "import torch; d = Data(x=torch.zeros(3,2), edge_index=torch.tensor([[0,1],[1,0]]), z1=torch.ones(3,2));"
            "assert d.num_nodes == 3 and d.num_edges == 2 and d.num_node_features == 2 and d.has_isolated_nodes() and not d.has_self_loops();"
            "assert d['z1'].sum() == 6 and 'z1' in d.keys; print('ok')")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "semigcn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"
