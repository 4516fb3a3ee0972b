from semigcn_b200.nn import GCNConv, ChebConv, Sequential  # noqa: F401
