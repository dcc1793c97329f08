"""Import shim: the product package lives in `lightweight-human-pose-estimation.pytorch_b200/`, a directory
name Python cannot import directly.  `import lwpose_b200` loads that directory as the package
`lwpose_b200` (sub-modules `lwpose_b200.models.with_mobilenet`, `lwpose_b200.modules.keypoints`, ...)."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lightweight-human-pose-estimation.pytorch_b200")
_spec = importlib.util.spec_from_file_location(
    "lwpose_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_pkg = importlib.util.module_from_spec(_spec)
sys.modules["lwpose_b200"] = _pkg
_spec.loader.exec_module(_pkg)
