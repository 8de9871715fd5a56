"""Import the UNMODIFIED reference from /root/reference on CPU.

Only used in the build container by ``make_golden.py`` (and optional cross-checks);
never on the GPU box (/root/reference does not exist there) and never by the product.
Modules that the reference imports but the hot path never calls are stubbed.
"""
import sys
import types

REF_ROOT = "/root/reference"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install_stubs():
    if "trimesh" not in sys.modules:
        tm = _stub("trimesh")
        tm.visual = _stub("trimesh.visual", texture=None, TextureVisuals=None)
    if "open3d" not in sys.modules:
        o3d = _stub("open3d")
        o3d.core = _stub("open3d.core")
    if "easydict" not in sys.modules:
        class EasyDict(dict):
            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError:
                    raise AttributeError(k)

            def __setattr__(self, k, v):
                self[k] = v
        _stub("easydict", EasyDict=EasyDict)
    if "termcolor" not in sys.modules:
        _stub("termcolor", colored=lambda s, *a, **k: s)
    if "imageio" not in sys.modules:
        _stub("imageio")
    try:
        import skimage  # noqa: F401
    except Exception:
        sk = _stub("skimage")
        sk.filters = _stub("skimage.filters", difference_of_gaussians=None, gabor_kernel=None)


def _pin_reference_packages():
    """The reference's `Utils` directory has no __init__.py (a namespace package), this repository's drop-in `Utils`
    does -- and a regular package beats a namespace package wherever it sits on sys.path.  Bind the name `Utils` (and
    the top-level module names both trees use) to the REFERENCE explicitly, and refuse to go on if anything that is
    already imported under those names comes from elsewhere."""
    for name in ("Utils", "PMVO", "HairGrow", "options", "log"):
        m = sys.modules.get(name)
        origin = getattr(m, "__file__", None) or (list(getattr(m, "__path__", [])) or [""])[0]
        if m is not None and not str(origin).startswith(REF_ROOT):
            for k in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
                del sys.modules[k]
    pkg = types.ModuleType("Utils")
    pkg.__path__ = [REF_ROOT + "/Utils"]
    pkg.__package__ = "Utils"
    sys.modules.setdefault("Utils", pkg)


def _assert_from_reference(mods):
    for name, m in list(mods.items()) + [(k, v) for k, v in sys.modules.items() if k.startswith("Utils.")]:
        f = getattr(m, "__file__", None)
        assert f is None or f.startswith(REF_ROOT), f"{name} was imported from {f}, not from the reference"


def import_reference():
    """returns dict of reference modules (PMVO, HairGrow, Camera_utils, PMVO_utils, Utils)."""
    install_stubs()
    for q in (REF_ROOT + "/preprocess_capture_data", REF_ROOT):
        if q in sys.path:
            sys.path.remove(q)
        sys.path.insert(0, q)
    _pin_reference_packages()
    import importlib
    mods = {}
    mods["PMVO"] = importlib.import_module("PMVO")
    mods["HairGrow"] = importlib.import_module("HairGrow")
    mods["Camera_utils"] = importlib.import_module("Utils.Camera_utils")
    mods["PMVO_utils"] = importlib.import_module("Utils.PMVO_utils")
    mods["Utils"] = importlib.import_module("Utils.Utils")
    _assert_from_reference(mods)
    return mods


def import_reference_gabor():
    """GaborFilter.py hard-codes .cuda(); on CPU make it the identity."""
    install_stubs()
    import torch
    if REF_ROOT + "/preprocess_capture_data" not in sys.path:
        sys.path.insert(0, REF_ROOT + "/preprocess_capture_data")
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    import importlib
    return importlib.import_module("GaborFilter")


def build_ref_pmvo(scene, patch_size=7, visible_threshold=1, conf_threshold=0.15):
    import numpy as np
    mods = import_reference()
    Camera = mods["Camera_utils"].Camera
    cams = {c["file"]: Camera(c["ndc_prj"], np.linalg.inv(np.array(c["pose"])), c["file"]) for c in scene.cams}
    Ori, Conf = scene.ref_ori_conf()
    pmvo = mods["PMVO"].PMVO(cams, scene.ref_depths(), Ori, Conf, scene.ref_masks(), device="cpu",
                             image_size=[scene.H, scene.W], patch_size=patch_size,
                             visible_threshold=visible_threshold, conf_threshold=conf_threshold)
    return pmvo, mods
