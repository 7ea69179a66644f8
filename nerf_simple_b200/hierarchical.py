"""Hierarchical (coarse + fine) sampling -- EXTENSION beyond the reference.

The reference stops at "coarse and fine is not implemented yet" (configs/lego.yaml:7; CoarseNet /
FineNet are empty classes, utils/nets.py:45-49).  BASELINE config 4 asks for 64 coarse + 128 fine
samples, so this module adds the original NeRF paper's scheme (Mildenhall et al. 2020, sec. 5.2) on
top of the same kernels: coarse pass -> inverse-CDF importance sampler (`nb200_sample_pdf_merge`)
-> fine pass over the merged, sorted depths.  No gradient flows through the sampler (as in the
paper's code).  The tests check it against a CPU restatement of the paper's sample_pdf, not of the
reference (which has none): parity unpinned by the reference.
"""
from __future__ import annotations

import torch

from . import _lib, config, ops


def sample_pdf_merge(ts, weights, Nf, u=None, det=False, seed=None, offset=0):
    """ts, weights [B,Nc] (CUDA) -> merged ascending depths [B, Nc+Nf]."""
    lib = _lib.load()
    ts = ops._f32c(ts, "ts")
    weights = ops._f32c(weights.detach(), "weights")
    B, Nc = ts.shape
    z = torch.empty((B, Nc + Nf), dtype=torch.float32, device=ts.device)
    if u is not None:
        mode, u = 0, ops._f32c(u, "u")
    elif det:
        mode = 1
    else:
        mode = 2
        if seed is None:
            seed, offset = config.next_philox(B * Nf)
    rc = lib.nb200_sample_pdf_merge(_lib.ptr(ts), _lib.ptr(weights), _lib.ptr(u), mode, int(seed or 0), int(offset),
                                    B, Nc, Nf, _lib.ptr(z), _lib.stream_ptr(ts.device))
    _lib.check(rc, "nb200_sample_pdf_merge")
    return z


def render_nerf_hierarchical(rays, net_coarse, net_fine, Nc=64, Nf=128, tn=2, tf=6, u_coarse=None, u_fine=None,
                             det_fine=False):
    """Coarse pass with Nc stratified samples, importance-resample Nf depths, fine pass on Nc+Nf.
    Returns (fine, coarse) where each is the 5-tuple of `render_nerf`
    (rgb, disparity, alpha, acc, weights)."""
    _lib.require_cuda(rays, "rays")
    rays = rays.float().contiguous()
    B = rays.size(0)
    if u_coarse is None and config.get_sampler() == "reference":
        u_coarse = torch.rand(B, Nc).to(rays.device, non_blocking=True)
    ts = ops.stratified_ts(B, Nc, tn, tf, u=u_coarse, device=rays.device)
    out_c = ops.mlp_apply(net_coarse, _lib.IN_RAYS, rays, ts, Nc).view(B, Nc, 4)
    coarse = ops.composite_apply(out_c, ts, rays, dirs_mode=1)
    if u_fine is None and not det_fine and config.get_sampler() == "reference":
        u_fine = torch.rand(B, Nf).to(rays.device, non_blocking=True)
    z = sample_pdf_merge(ts, coarse[4], Nf, u=u_fine, det=det_fine)
    out_f = ops.mlp_apply(net_fine, _lib.IN_RAYS, rays, z, Nc + Nf).view(B, Nc + Nf, 4)
    fine = ops.composite_apply(out_f, z, rays, dirs_mode=1)
    return fine, coarse
