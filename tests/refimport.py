"""Import the real reference (read-only at /root/reference) for golden generation / live cross-checks.

Only usable in the build container; on the GPU box `/root/reference` does not exist and
`available()` is False.  `pycocotools` is absent offline, so stubs are injected before `val` / `demo`
are imported."""
import os
import sys
import types

REF = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF, "modules"))


def load():
    """Returns a namespace with the reference's hot-path callables."""
    if not available():
        raise RuntimeError("reference not available")
    sys.dont_write_bytecode = True
    for name in ("pycocotools", "pycocotools.coco", "pycocotools.cocoeval", "pycocotools.mask"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.COCO = object
            m.COCOeval = object
            sys.modules[name] = m
    # the reference's top-level package names (models, modules, val, demo) must not collide with ours:
    # ours live under lwpose_b200.*, so a plain sys.path entry is safe.
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import importlib
    ns = types.SimpleNamespace()
    ns.keypoints = importlib.import_module("modules.keypoints")
    ns.with_mobilenet = importlib.import_module("models.with_mobilenet")
    ns.conv = importlib.import_module("modules.conv")
    ns.val = importlib.import_module("val")
    ns.demo = importlib.import_module("demo")
    ns.load_state = importlib.import_module("modules.load_state")
    return ns
