"""Import alias: the package directory is named `obmd-lammps-extension_b200` (not a valid
Python identifier), so `import obmd_b200` loads it through importlib."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("obmd-lammps-extension_b200")
globals().update({k: getattr(_pkg, k) for k in _pkg.__all__})
package = _pkg
