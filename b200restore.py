"""Import alias: `import b200restore` loads the package that lives in
`image-restoration-for-road-sign-recognition-in-autonomous-driving_b200/` (a directory name with hyphens cannot be
written in an `import` statement)."""
import importlib
import sys
from pathlib import Path

_ROOT = Path(__file__).resolve().parent
if str(_ROOT) not in sys.path:
    sys.path.insert(0, str(_ROOT))
_pkg = importlib.import_module("image-restoration-for-road-sign-recognition-in-autonomous-driving_b200")
sys.modules[__name__] = _pkg
