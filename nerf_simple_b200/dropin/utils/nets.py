"""Drop-in for the reference's `utils.nets`: same names, B200 engine underneath."""
from nerf_simple_b200.nets import *  # noqa: F401,F403
from nerf_simple_b200.nets import Nerf, CoarseNet, FineNet  # noqa: F401
from nerf_simple_b200.xyz import *  # noqa: F401,F403  (the reference re-exports utils.xyz here)
