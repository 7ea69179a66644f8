"""Writes a tiny Blender-synthetic-format dataset (the layout utils/dataload.py:12-104 reads):
train/ val/ test/ PNGs (+ test depth/normal maps) and transforms_{train,val,test}.json with
camera_angle_x and dome poses.  Test infrastructure."""
import json
import os

import cv2
import numpy as np

from oracle import nerf_oracle as O

FOV = 0.6911112070083618


def write_dataset(root, H=16, W=16, n_train=3, n_val=2, n_test=2, seed=0):
    rng = np.random.default_rng(seed)
    poses = O.poses_to_render(4, -30, n_train + n_val + n_test + 1)
    k = 0
    for split, n in (("train", n_train), ("val", n_val), ("test", n_test)):
        os.makedirs(os.path.join(root, split), exist_ok=True)
        frames = []
        for i in range(n):
            img = (rng.random((H, W, 3)) * 255).astype(np.uint8)
            # smooth blobs so that a few training steps can actually fit something
            img = cv2.GaussianBlur(img, (0, 0), 3)
            cv2.imwrite(os.path.join(root, split, f"r_{i}.png"), img)
            if split == "test":
                cv2.imwrite(os.path.join(root, split, f"r_{i}_depth_0001.png"), img)
                cv2.imwrite(os.path.join(root, split, f"r_{i}_normal_0001.png"), img)
            frames.append({"file_path": f"./{split}/r_{i}", "rotation": 0.0,
                           "transform_matrix": poses[k].astype(np.float64).tolist()})
            k += 1
        with open(os.path.join(root, f"transforms_{split}.json"), "w") as fh:
            json.dump({"camera_angle_x": FOV, "frames": frames}, fh)
    return root
