"""`utils.rendering` of the reference on the B200 engine: same functions, same arguments, same
return values; three kernel launches per render_nerf call instead of ~260 ATen launches.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from . import _lib, config, ops

__all__ = ["render_nerf", "volume_render", "render_image", "render_poses"]


def render_nerf(rays, net, N, tn=2, tf=6):
    """rays [B,6] (CUDA) -> (rgb [B,3], disparity [B], alpha [B,N], acc [B], weights [B,N]).
    utils/rendering.py:13-45.  In the default 'reference' sampler mode the jitter is one
    torch.rand(B,N) from the CPU global generator, exactly like :28, so a seeded run sees the
    same sample depths as the reference; config.set_sampler('philox') draws on the device."""
    _lib.require_cuda(rays, "rays")
    B = rays.size(0)
    rays = rays.float().contiguous()
    if config.get_sampler() == "reference":
        u = torch.rand(B, N)
        ts = ops.stratified_ts(B, N, tn, tf, u=u.to(rays.device, non_blocking=True))
    else:
        ts = ops.stratified_ts(B, N, tn, tf, device=rays.device)
    out = ops.mlp_apply(net, _lib.IN_RAYS, rays, ts, N).view(B, N, 4)
    return ops.composite_apply(out, ts, rays, dirs_mode=1)


def volume_render(nerf_outs, ts, dirs):
    """nerf_outs [B,N,4], ts [B,N], dirs [B,3] -> (rgb, disp, alpha, acc, weights).
    utils/rendering.py:47-85."""
    return ops.composite_apply(nerf_outs, ts, dirs, dirs_mode=0)


def _render_chunk(chunk, net, N):
    """(rgb, disparity) of one no-grad chunk: render_nerf's three kernels, or -- with
    config.set_fused_render(True) and a supported shape (bf16, N in {32,64,128}) -- the single fused
    kernel.  Same sampler semantics as render_nerf either way."""
    if not (config.get_fused_render() and ops.fused_render_supported(net, N)):
        rgb, depth, _, _, _ = render_nerf(chunk, net, N=N)
        return rgb, depth
    if config.get_sampler() == "reference":
        u = torch.rand(chunk.size(0), N)                       # utils/rendering.py:28, CPU global generator
        ts = ops.stratified_ts(chunk.size(0), N, 2, 6, u=u.to(chunk.device, non_blocking=True))
        rgb, depth, _ = ops.render_fused(net, N, rays=chunk, ts=ts)
    else:
        rgb, depth, _ = ops.render_fused(net, N, rays=chunk)
    return rgb, depth


def _render_chunks(net, rays, batch_size, N=128):
    """Chunked no-grad render of a ray table; unlike :100,:143 the remainder chunk is kept."""
    rgbs, depths = [], []
    dev = next(net.parameters()).device
    with torch.no_grad():
        for s in range(0, rays.size(0), batch_size):
            chunk = rays[s:s + batch_size].to(dev, non_blocking=True).float().contiguous()
            rgb, depth = _render_chunk(chunk, net, N)
            rgbs.append(rgb.clamp_(0.0, 1.0))       # :103,:146
            depths.append(depth)
    return torch.cat(rgbs), torch.cat(depths)


def render_image(net, rg, batch_size=64000, im_idx=0, im_set='val', N=128):
    """Render image `im_idx` of split `im_set` (N=128 samples/ray like the reference; `N` is an
    extension); returns CPU tensors (rgb [1,H,W,3], disparity [1,H,W,1], gt [1,H,W,3]).
    utils/rendering.py:88-113."""
    gt_img = rg.samples[im_set][im_idx]['img']
    H, W = gt_img.shape[0], gt_img.shape[1]
    n = H * W
    net = net.cuda()
    rays = rg.rays_dataset[im_set][im_idx * n:(im_idx + 1) * n, :]
    rgb, depth = _render_chunks(net, rays, batch_size, N=N)
    return rgb.cpu().reshape(1, H, W, 3), depth.cpu().reshape(1, H, W, 1), gt_img.reshape(1, H, W, 3)


def render_poses(net, poses, cam_params, batch_size, savepath=''):
    """Render every pose (list of 4x4 camera-to-world) at [H,W,f] and write an mp4, like
    utils/rendering.py:116-160 -- but rays are generated on the device per chunk (24 B/ray of
    H2D traffic and the 461 MB host ray table disappear) and the frames are clipped, BGR-swapped and
    quantised to uint8 on the device (:146,:158-159), so 3 B/pixel come back instead of 12.
    Returns the float RGB frames (the reference returns None; callers that ignore it are unaffected)."""
    import cv2
    H, W, f = int(cam_params[0]), int(cam_params[1]), float(cam_params[2])
    n = H * W
    net = net.cuda()
    dev = next(net.parameters()).device
    pose_t = torch.stack([torch.as_tensor(p).float() for p in poses]).to(dev)
    frames, frames_u8 = [], []
    with torch.no_grad():
        for idx in range(len(poses)):
            rgbs, depths = [], []
            for s in range(0, n, batch_size):
                cnt = min(batch_size, n - s)
                if config.get_fused_render() and ops.fused_render_supported(net, 128) and config.get_sampler() == "philox":
                    rgb, depth, _ = ops.render_fused(net, 128, poses=pose_t, H=H, W=W, f=f, ray_begin=idx * n + s, n_rays=cnt)
                else:
                    rays = ops.generate_rays(pose_t, H, W, f, ray_begin=idx * n + s, n_rays=cnt)
                    rgb, depth = _render_chunk(rays, net, 128)
                rgbs.append(rgb.clamp_(0.0, 1.0))
                depths.append(depth)
            frame = torch.cat(rgbs).reshape(H, W, 3)
            frames.append(frame)
            frames_u8.append(ops.frame_to_u8(frame, bgr=True).cpu().numpy())
    tstamp = str(time.time())
    out = cv2.VideoWriter(os.path.join(savepath, f'nerf_rgb{tstamp[-10:]}.mp4'),
                          cv2.VideoWriter_fourcc('m', 'p', '4', 'v'), 15, (H, W))   # (H,W) as in :156
    for frame in frames_u8:
        out.write(frame)
    out.release()
    return [fr.cpu().numpy() for fr in frames]
