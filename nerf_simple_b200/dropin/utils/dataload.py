"""Drop-in for the reference's `utils.dataload`: same names, B200 engine underneath."""
from nerf_simple_b200.dataload import *  # noqa: F401,F403
from nerf_simple_b200.dataload import load_data, rays_dataset, RayGenerator  # noqa: F401
