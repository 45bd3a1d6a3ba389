"""Import shim: the package directory is named `duckdb-parquet-parser_b200` (with dashes),
which the `import` statement cannot spell.  `import pqb200` gives the same module."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_m = importlib.import_module("duckdb-parquet-parser_b200")
globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
module = _m
