"""Shim package (see semigcn_b200/compat/__init__.py)."""
from . import nn, data  # noqa: F401

__version__ = "2.2.0+semigcn_b200"
