"""Drop-in module name of the reference (Utils/Camera_utils.py)."""
from monohair_b200.camera import Camera, load_cam, parsing_camera  # noqa: F401
