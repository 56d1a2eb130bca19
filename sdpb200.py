"""Import shim: the package directory is named `stochastic-inventory_b200` (not a Python identifier),
so `import sdpb200` loads it through importlib and re-exports its namespace."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("stochastic-inventory_b200")
globals().update({k: v for k, v in vars(_pkg).items() if not k.startswith("__")})
package = _pkg
