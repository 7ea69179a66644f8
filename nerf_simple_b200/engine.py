"""Device-resident drivers of the hot path: what bench.py and the multi-GPU entry points call.

`FrameRenderer` renders whole novel views (the loop of utils/rendering.py:139-151) with four kernels
per frame: ray generation, Philox sampler, fused posenc+MLP, compositing.  With `fused=True` (or
config.set_fused_render(True)) a coarse-only frame with N in {32, 64, 128} is ONE launch: rays from
the camera, jitter, the tcgen05 MLP chain and the compositing all happen inside
chain_kernel<FwdEpi<render>>, so neither rays, sample depths nor per-sample (r,g,b,sigma) touch
HBM -- measured 2 % slower on B200 (the chain kernel is bound by its epilogue warps, not by HBM),
hence not the default.  `shard_range` / `render_sharded` split the rays
of a frame across ranks with one final gather.
"""
from __future__ import annotations

import torch

from . import _lib, config, ops


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced [begin, end) range of `n_items` for `rank` of `world`."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class FrameRenderer:
    def __init__(self, net, H, W, f, N=64, tn=2.0, tf=6.0, seed=1, precision=None, net_fine=None, Nf=0, fused=None):
        self.net, self.H, self.W, self.f, self.N = net, int(H), int(W), float(f), int(N)
        self.net_fine, self.Nf = net_fine, int(Nf)   # hierarchical extension: N coarse + Nf fine samples
        self.tn, self.tf, self.seed = float(tn), float(tf), int(seed)
        self.precision = precision
        fused = config.get_fused_render() if fused is None else fused
        self.fused = bool(fused) and net_fine is None and ops.fused_render_supported(net, N, precision)
        self.device = next(net.parameters()).device
        self._offset = 0
        self.launches = 0          # kernels of libnerf_b200 launched so far
        self.mlp_events = []       # optional (start, stop) CUDA events around the MLP kernel
        self._copy_stream, self._last_copy = None, None

    def render_rays(self, poses_dev, ray_begin, n_rays, time_mlp=False, philox_base=None):
        """rgb [n,3] clipped to [0,1] and disparity [n] for rays [ray_begin, ray_begin+n) of the
        pose table `poses_dev` ([P,4,4] on the device).  philox_base: position of the first sample of this call
        in the jitter stream (default: continue where the previous call stopped)."""
        if philox_base is not None:
            self._offset = int(philox_base)
        with torch.no_grad():
            if self.fused:
                if time_mlp:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                rgb, disp, _ = ops.render_fused(self.net, self.N, poses=poses_dev, H=self.H, W=self.W, f=self.f,
                                                ray_begin=ray_begin, n_rays=n_rays, tn=self.tn, tf=self.tf,
                                                seed=self.seed, offset=self._offset, precision=self.precision)
                if time_mlp:
                    e1.record()
                    self.mlp_events.append((e0, e1))
                self._offset += (n_rays * self.N + 3) // 4
                self.launches += 1
                return rgb.clamp_(0.0, 1.0), disp
            rays = ops.generate_rays(poses_dev, self.H, self.W, self.f, ray_begin, n_rays)
            ts = ops.stratified_ts(n_rays, self.N, self.tn, self.tf, device=self.device, seed=self.seed,
                                   offset=self._offset)
            self._offset += (n_rays * self.N + 3) // 4
            if time_mlp:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            out = ops.mlp_apply(self.net, _lib.IN_RAYS, rays, ts, self.N, precision=self.precision)
            if time_mlp:
                e1.record()
                self.mlp_events.append((e0, e1))
            if self.net_fine is None:
                rgb, disp, _ = ops.composite_apply(out.view(n_rays, self.N, 4), ts, rays, dirs_mode=1,
                                                   want_alpha_weights=False)
                self.launches += 4
                return rgb.clamp_(0.0, 1.0), disp
            # hierarchical extension: importance-resample Nf depths from the coarse weights, fine pass
            from .hierarchical import sample_pdf_merge
            w = ops.composite_apply(out.view(n_rays, self.N, 4), ts, rays, dirs_mode=1)[4]
            z = sample_pdf_merge(ts, w, self.Nf, seed=self.seed + 7919, offset=self._offset)
            self._offset += (n_rays * self.Nf + 3) // 4
            NA = self.N + self.Nf
            out_f = ops.mlp_apply(self.net_fine, _lib.IN_RAYS, rays, z, NA, precision=self.precision)
            rgb, disp, _ = ops.composite_apply(out_f.view(n_rays, NA, 4), z, rays, dirs_mode=1, want_alpha_weights=False)
            self.launches += 7
            return rgb.clamp_(0.0, 1.0), disp

    def frame_quads(self):
        """Philox calls (4 draws each) one whole frame consumes."""
        n = self.H * self.W
        return (n * self.N + 3) // 4 + (n * self.Nf + 3) // 4

    def render_frame(self, poses_dev, idx=0, time_mlp=False):
        n = self.H * self.W
        rgb, disp = self.render_rays(poses_dev, idx * n, n, time_mlp)
        return rgb.view(self.H, self.W, 3), disp.view(self.H, self.W)

    def render_frame_host(self, pose_cpu_pinned, out_rgb_pinned, out_disp_pinned, wait=True):
        """End-to-end call with HOST buffers: pose (pinned, [4,4]) in, frame out (pinned).
        wait=False returns a CUDA event instead of blocking: the device->host copy of this frame then
        runs on a side stream under the next frame's kernels (the caller synchronises the event -- or
        calls `finish()` -- before reading the buffers and must not reuse them for another frame earlier)."""
        pose = pose_cpu_pinned.to(self.device, non_blocking=True).view(1, 4, 4)
        rgb, disp = self.render_frame(pose, 0)
        main = torch.cuda.current_stream(self.device)
        if wait:
            out_rgb_pinned.copy_(rgb, non_blocking=True)
            out_disp_pinned.copy_(disp, non_blocking=True)
            main.synchronize()
            return out_rgb_pinned, out_disp_pinned
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        rendered = torch.cuda.Event()
        rendered.record(main)
        self._copy_stream.wait_event(rendered)
        with torch.cuda.stream(self._copy_stream):
            out_rgb_pinned.copy_(rgb, non_blocking=True)
            out_disp_pinned.copy_(disp, non_blocking=True)
            rgb.record_stream(self._copy_stream)
            disp.record_stream(self._copy_stream)
            done = torch.cuda.Event()
            done.record(self._copy_stream)
        self._last_copy = done
        return done

    def render_frame_u8_host(self, pose_cpu_pinned, out_u8_pinned, bgr=True):
        """Video path (render_poses, utils/rendering.py:139-160): pose (pinned [4,4]) in, uint8 frame [H,W,3] (pinned)
        out, clipped / BGR-swapped / quantised on the device -- 3 B/pixel cross PCIe instead of 16."""
        pose = pose_cpu_pinned.to(self.device, non_blocking=True).view(1, 4, 4)
        rgb, _ = self.render_frame(pose, 0)
        u8 = ops.frame_to_u8(rgb, bgr=bgr)
        self.launches += 1
        out_u8_pinned.copy_(u8, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return out_u8_pinned

    def finish(self):
        """Block until every frame handed to render_frame_host(wait=False) is on the host."""
        if self._last_copy is not None:
            self._last_copy.synchronize()


def gather_shards(local, n_items, rank, world, group=None):
    """all_gather of per-rank row blocks whose sizes follow `shard_range` (they differ by at most one
    row, so every rank pads to the largest block; collectives need equal sizes)."""
    import torch.distributed as dist
    if n_items % world == 0:     # equal bands (every lego size: 1600^2 / 8, 800^2 / 8, ...): one collective into one buffer
        full = local.new_empty((n_items,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(full, local.contiguous(), group=group)
        return full
    spans = [shard_range(n_items, r, world) for r in range(world)]
    width = max(e - b for b, e in spans)
    padded = local
    if local.shape[0] < width:
        padded = torch.cat([local, local.new_zeros((width - local.shape[0],) + tuple(local.shape[1:]))])
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    return torch.cat([p[:e - b] for p, (b, e) in zip(parts, spans)])


def render_sharded(renderer: FrameRenderer, poses_dev, idx, rank, world, gather=True):
    """Ray-sharded render of frame `idx`: each rank renders a contiguous band of rays; one
    all_gather of (rgb, disp) = 16 B/ray assembles the frame on every rank.  Every sample keeps the place in the
    jitter stream it has in the unsharded frame (the Philox position is a function of the ray's index in the frame),
    so for N % 4 == 0 the assembled frame is bit-identical to the one a single GPU renders."""
    n = renderer.H * renderer.W
    b, e = shard_range(n, rank, world)
    base = renderer._offset
    exact = renderer.net_fine is None and (b * renderer.N) % 4 == 0
    rgb, disp = renderer.render_rays(poses_dev, idx * n + b, e - b, philox_base=base + (b * renderer.N) // 4 if exact else None)
    renderer._offset = base + renderer.frame_quads()          # every rank advances by one whole frame
    if world == 1:
        return rgb.view(renderer.H, renderer.W, 3), disp.view(renderer.H, renderer.W)
    if not gather:
        return rgb, disp
    full = gather_shards(torch.cat([rgb, disp[:, None]], dim=1), n, rank, world)     # [n, 4]
    return full[:, :3].reshape(renderer.H, renderer.W, 3), full[:, 3].reshape(renderer.H, renderer.W)
