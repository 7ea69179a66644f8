"""`utils.dataload` of the reference: Blender-synthetic loader + per-split ray tables.

Host-side data plumbing (PNG decode, JSON poses) -- out of the accelerated path but
part of the call surface train.py/test.py import.  The ray tables are produced by the device ray
kernel and returned as CPU tensors in the reference's layout, because train.py:49 indexes a CPU
image tensor with the returned ray ids.
"""
from __future__ import annotations

import glob
import json
import os
import re

import numpy as np
import torch

from . import _lib, config, ops


def _natural_key(path):
    """Case-insensitive natural sort key (the reference uses natsort, utils/dataload.py:35)."""
    return [int(tok) if tok.isdigit() else tok.lower() for tok in re.split(r"(\d+)", path)]


def _read_rgb(path, half_res):
    import cv2
    img = cv2.cvtColor(cv2.imread(path), cv2.COLOR_BGR2RGB) / 255.0
    if half_res:
        h, w = img.shape[:2]
        img = cv2.resize(img, (w // 2, h // 2), interpolation=cv2.INTER_AREA)
    return img


def load_data(path, half_res=True, num_imgs=-1):
    """Returns (samples, [H, W, f]) with samples[split] = list of {'img','transform','metadata'}
    (+ 'img_depth','img_normal' for test), as utils/dataload.py:12-112."""
    import cv2
    splits = {}
    files = {
        "train": sorted(glob.glob(os.path.join(path, "train", "*")), key=_natural_key),
        "val": sorted(glob.glob(os.path.join(path, "val", "*")), key=_natural_key),
        "test": sorted([os.path.join(path, "test", fn) for fn in os.listdir(os.path.join(path, "test"))
                        if re.match(r"r_[0-9]+.png", fn)], key=_natural_key),
    }
    depth_files = sorted(glob.glob(os.path.join(path, "test", "r_*_depth*")), key=_natural_key)
    normal_files = sorted(glob.glob(os.path.join(path, "test", "r_*_normal*")), key=_natural_key)
    meta = {}
    for split in ("train", "test", "val"):
        with open(os.path.join(path, f"transforms_{split}.json")) as fh:
            meta[split] = json.load(fh)
    img = None
    for split in ("train", "val", "test"):
        count = len(files[split]) if num_imgs < 0 else num_imgs
        items = []
        for i in range(count):
            frame = meta[split]["frames"][i]
            pose = torch.from_numpy(np.array(frame["transform_matrix"])).float()
            if split == "test":
                # depth / normal maps stay full-res in the reference (:93-94)
                full = cv2.cvtColor(cv2.imread(files[split][i]), cv2.COLOR_BGR2RGB) / 255.0
                img = full
                if half_res:
                    h, w = full.shape[:2]
                    img = cv2.resize(full, (w // 2, h // 2), interpolation=cv2.INTER_AREA)
                items.append({"img": img,
                              "img_depth": cv2.cvtColor(cv2.imread(depth_files[i]), cv2.COLOR_BGR2RGB) / 255.0,
                              "img_normal": cv2.cvtColor(cv2.imread(normal_files[i]), cv2.COLOR_BGR2RGB) / 255.0,
                              "transform": pose, "metadata": frame})
            else:
                items.append({"img": _read_rgb(files[split][i], half_res), "transform": pose,
                              "metadata": frame})
        splits[split] = items
    # H, W come from the last test image; f from the train fov (:102-105)
    fov = meta["train"]["camera_angle_x"]
    H, W = img.shape[:2]
    f = W / (2 * np.tan(fov / 2))
    return splits, [H, W, f]


def rays_dataset(samples, cam_params):
    """{'train','test','val'} -> [num_images*H*W, 6] CPU ray tables (utils/dataload.py:114-129)."""
    H, W, f = cam_params
    rays = {}
    for split in ("train", "test", "val"):
        poses = torch.stack([s["transform"] for s in samples[split]]).cuda()
        rays[split] = ops.generate_rays(poses, int(H), int(W), float(f)).cpu()
    return rays


class RayGenerator:
    def __init__(self, path, half_res=True, num_imgs=-1):
        samples, cam_params = load_data(path, half_res, num_imgs)
        self.samples = samples
        self.cam_params = cam_params
        self.H, self.W, self.f = cam_params
        self.rays_dataset = rays_dataset(self.samples, cam_params)

    def select(self, mode='train', N=4096):
        """N random rays of a split and their row ids (utils/dataload.py:141-153)."""
        table = self.rays_dataset[mode]
        if config.get_select() == "device":
            return self._select_device(mode, N)
        ray_ids = torch.randperm(table.size(0))[:N]
        return table[ray_ids, :], ray_ids

    def _select_device(self, mode, N):
        """the same call, without the CPU randperm over the whole table.  Indices are
        drawn on the device (uniform WITH replacement, Philox), the rays are gathered there and returned
        as a CUDA tensor (train.py:51's `.cuda()` is then a no-op); the ids come back as the CPU int64
        tensor train.py:49 indexes the image table with."""
        lib = _lib.load()
        if not hasattr(self, "_dev_tables"):
            self._dev_tables, self._sel_offset = {}, 0
        if mode not in self._dev_tables:
            self._dev_tables[mode] = self.rays_dataset[mode].cuda().float().contiguous()
        table = self._dev_tables[mode]
        rays = torch.empty((N, 6), dtype=torch.float32, device=table.device)
        ids = torch.empty((N,), dtype=torch.int64, device=table.device)
        _lib.check(lib.nb200_select_rays(_lib.ptr(table), None, table.shape[0], config._state["seed"] ^ 0x5E1EC7,
                                         self._sel_offset, N, _lib.ptr(rays), None, _lib.ptr(ids),
                                         _lib.stream_ptr(table.device)), "nb200_select_rays")
        self._sel_offset += N
        return rays, ids.cpu()

    def select_imgs(self, mode='train', N=4096, im_idxs=[0, 1, 2]):
        """N random rays restricted to the given images (utils/dataload.py:155-179)."""
        per = self.H * self.W
        ids = np.concatenate([np.arange(i * per, (i + 1) * per) for i in im_idxs])
        pick = np.random.choice(ids.shape[0], (N,), replace=False)
        ray_ids = ids[pick]
        return self.rays_dataset[mode][torch.from_numpy(ray_ids), :], ray_ids
