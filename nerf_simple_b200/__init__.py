"""nerf_simple_b200 -- B200-native (sm_100a) engine behind the Nerf-Simple call surface.

    from nerf_simple_b200.nets import Nerf
    from nerf_simple_b200.rendering import render_nerf, volume_render, render_image, render_poses

or, to drive the reference's own train.py / test.py unchanged:

    python -m nerf_simple_b200.run /path/to/Nerf-Simple/train.py --config_path cfg.yaml

Importing the package does not load CUDA; the first compute call loads libnerf_b200.so and
raises if it is missing (there is no CPU or eager-PyTorch fallback).
"""
from . import config  # noqa: F401
from .config import set_precision, set_sampler  # noqa: F401

__version__ = "0.1.0"
