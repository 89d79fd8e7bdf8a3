"""Import alias: `import b200mosaic` == the package in `real-time-video-mosaic_b200/` (whose directory name, fixed by
the project layout, is not a Python identifier)."""
import importlib
import sys
from pathlib import Path

_root = str(Path(__file__).resolve().parent.parent)
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("real-time-video-mosaic_b200")
sys.modules[__name__] = _pkg
