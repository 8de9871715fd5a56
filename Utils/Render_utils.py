"""Drop-in module name of the reference (Utils/Render_utils.py): the depth-map producer of the PMVO inputs."""
from monohair_b200.render import DepthRenderer, render_bust_hair_depth  # noqa: F401
