"""Drop-in for the reference's `utils.xyz`: same names, B200 engine underneath."""
from nerf_simple_b200.xyz import *  # noqa: F401,F403
from nerf_simple_b200.xyz import (gamma, positional_encoder, rays_single_cam, polar_to_mat,  # noqa: F401
                                  phi_to_mat, spherical_to_pose, poses_to_render)
import numpy as np  # noqa: F401,E402  (star-imports of the reference module leak these)
import torch  # noqa: F401,E402
