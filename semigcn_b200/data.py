"""``torch_geometric.data.Data`` stand-in: exactly the surface util/datamaker.py:13-25,105
touches (attribute bag, ``keys``, ``num_nodes``, ``num_edges``, ``num_node_features``,
``has_isolated_nodes()``, ``has_self_loops()``, ``__getitem__``, ``to``)."""
from __future__ import annotations

import torch


class Data:
    def __init__(self, x=None, edge_index=None, **kwargs):
        self._store = {}
        if x is not None:
            self._store["x"] = x
        if edge_index is not None:
            self._store["edge_index"] = edge_index
        self._store.update(kwargs)

    def __getattr__(self, name):
        store = self.__dict__.get("_store", {})
        if name in store:
            return store[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name == "_store":
            object.__setattr__(self, name, value)
        else:
            self._store[name] = value

    def __getitem__(self, key):
        return self._store[key]

    def __setitem__(self, key, value):
        self._store[key] = value

    def __contains__(self, key):
        return key in self._store

    @property
    def keys(self):
        return [k for k, v in self._store.items() if v is not None]

    @property
    def num_nodes(self):
        x = self._store.get("x")
        if x is not None:
            return int(x.shape[0])
        ei = self._store.get("edge_index")
        return int(ei.max()) + 1 if ei is not None and ei.numel() else 0

    @property
    def num_edges(self):
        ei = self._store.get("edge_index")
        return int(ei.shape[1]) if ei is not None else 0

    @property
    def num_node_features(self):
        x = self._store.get("x")
        return 0 if x is None else (1 if x.dim() == 1 else int(x.shape[1]))

    num_features = num_node_features

    def has_isolated_nodes(self) -> bool:
        ei = self._store.get("edge_index")
        n = self.num_nodes
        if ei is None or n == 0:
            return n > 0
        keep = ei[0] != ei[1]
        seen = torch.zeros(n, dtype=torch.bool, device=ei.device)
        seen[ei[0][keep]] = True
        seen[ei[1][keep]] = True
        return bool((~seen).any())

    contains_isolated_nodes = has_isolated_nodes

    def has_self_loops(self) -> bool:
        ei = self._store.get("edge_index")
        return bool((ei[0] == ei[1]).any()) if ei is not None else False

    contains_self_loops = has_self_loops

    def to(self, device, *args, **kwargs):
        for k, v in list(self._store.items()):
            if torch.is_tensor(v):
                self._store[k] = v.to(device, *args, **kwargs)
        return self

    def __repr__(self):
        parts = [f"{k}={list(v.shape)}" if torch.is_tensor(v) else f"{k}={type(v).__name__}" for k, v in self._store.items()]
        return "Data(" + ", ".join(parts) + ")"
