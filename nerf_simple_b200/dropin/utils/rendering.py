"""Drop-in for the reference's `utils.rendering`: same names, B200 engine underneath."""
from nerf_simple_b200.rendering import *  # noqa: F401,F403
from nerf_simple_b200.rendering import render_nerf, volume_render, render_image, render_poses  # noqa: F401
