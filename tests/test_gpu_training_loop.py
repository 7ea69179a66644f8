"""GPU test of the host drivers exactly as train.py / test.py use them (train.py:33-82,
test.py:25-45), on a synthetic Blender-format dataset: RayGenerator -> select -> render_nerf ->
MSELoss -> backward -> Adam -> render_image -> state_dict round trip."""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from synth_dataset import write_dataset

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    return write_dataset(str(tmp_path_factory.mktemp("blender")), H=16, W=16)


def test_ray_generator_matches_reference_layout(dataset):
    from nerf_simple_b200.dataload import RayGenerator, load_data
    samples, cam = load_data(dataset, half_res=False)
    assert [len(samples[k]) for k in ("train", "val", "test")] == [3, 2, 2]
    H, W, f = cam
    assert (H, W) == (16, 16) and abs(f - 16 / (2 * np.tan(0.6911112070083618 / 2))) < 1e-9
    assert samples["train"][0]["img"].shape == (16, 16, 3) and samples["train"][0]["img"].dtype == np.float64
    assert "img_depth" in samples["test"][0] and samples["train"][1]["metadata"]["file_path"] == "./train/r_1"
    rg = RayGenerator(dataset, half_res=True)                    # INTER_AREA half-res like the reference
    assert (rg.H, rg.W) == (8, 8) and rg.samples["val"][0]["img"].shape == (8, 8, 3)
    assert not rg.rays_dataset["train"].is_cuda and rg.rays_dataset["train"].shape == (3 * 64, 6)
    poses = np.stack([s["transform"].numpy() for s in rg.samples["train"]])
    ref = O.world_rays(poses, O.rays_single_cam(8, 8, rg.f))     # utils/dataload.py:114-129
    assert np.abs(rg.rays_dataset["train"].numpy() - ref).max() <= 1e-6
    rays, ids = rg.select("train", N=50)
    assert rays.shape == (50, 6) and torch.equal(rays, rg.rays_dataset["train"][ids])
    rays, ids = rg.select_imgs("train", N=40, im_idxs=[1])
    assert rays.shape == (40, 6) and ids.min() >= 64 and ids.max() < 128
    # opt-in device-side selection: same call, rays on the device, ids on the host
    from nerf_simple_b200 import config
    config.set_select("device")
    try:
        rays, ids = rg.select("train", N=50)
        assert rays.is_cuda and not ids.is_cuda and ids.dtype == torch.int64 and rays.shape == (50, 6)
        assert torch.equal(rays.cpu(), rg.rays_dataset["train"][ids]) and int(ids.min()) >= 0 and int(ids.max()) < 3 * 64
        _, ids2 = rg.select("train", N=50)
        assert not torch.equal(ids, ids2)                      # the Philox offset advances from call to call
    finally:
        config.set_select("reference")


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_train_loop_like_train_py(dataset, precision, tmp_path):
    from nerf_simple_b200 import config
    from nerf_simple_b200.dataload import RayGenerator
    from nerf_simple_b200.nets import Nerf
    from nerf_simple_b200.rendering import render_image, render_nerf
    config.set_precision(precision)
    config.set_sampler("reference")
    torch.manual_seed(0)
    rg = RayGenerator(dataset, half_res=False, num_imgs=2)
    train_imgs = torch.stack([torch.from_numpy(s["img"]) for s in rg.samples["train"]]).reshape(-1, 3)
    net = Nerf().cuda()
    criterion = torch.nn.MSELoss()
    optimizer = torch.optim.Adam(net.parameters(), lr=5e-4)
    losses = []
    for i in range(60):
        rays, ray_ids = rg.select(mode="train", N=256)
        gt = train_imgs[ray_ids, :].float().cuda()
        optimizer.zero_grad()
        rgb, depth, alpha, acc, w = render_nerf(rays.cuda(), net, 32)
        loss = criterion(rgb, gt)
        loss.backward()
        optimizer.step()
        losses.append(loss.item())
    assert np.mean(losses[-10:]) < 0.5 * np.mean(losses[:5]), losses[::10]
    rgb_img, depth_img, gt_img = render_image(net, rg, batch_size=100, im_idx=1, im_set="train")
    assert rgb_img.shape == (1, 16, 16, 3) and depth_img.shape == (1, 16, 16, 1) and gt_img.shape == (1, 16, 16, 3)
    mse = torch.mean((rgb_img - torch.from_numpy(gt_img).float()) ** 2)
    assert torch.isfinite(mse)
    # checkpoint round trip exactly like train.py:87 / test.py:28
    path = str(tmp_path / "ckpt.pth")
    torch.save(net.state_dict(), path)
    net2 = Nerf().cuda()
    net2.load_state_dict(torch.load(path), strict=True)
    torch.manual_seed(5)
    a = render_nerf(rg.rays_dataset["val"][:64].cuda(), net, 32)[0]
    torch.manual_seed(5)
    b = render_nerf(rg.rays_dataset["val"][:64].cuda(), net2, 32)[0]
    assert torch.equal(a, b)
    config.set_precision("bf16")


def test_device_trainer_converges():
    from nerf_simple_b200 import ops
    from nerf_simple_b200.nets import Nerf
    from nerf_simple_b200.trainer import Trainer
    from nerf_simple_b200.xyz import poses_to_render
    torch.manual_seed(0)
    net = Nerf().cuda()
    poses = torch.stack(poses_to_render(4, -30, 4)).cuda()
    rays = ops.generate_rays(poses, 32, 32, 44.4)
    gt = torch.sigmoid(rays[:, 3:6] * 3)                      # a smooth, learnable colour field
    tr = Trainer(net, rays, gt, N=32, batch_size=1024, precision="bf16")
    losses = [tr.step(sync_loss=True) for _ in range(80)]
    assert np.mean(losses[-10:]) < 0.3 * np.mean(losses[:5]), losses[::10]
    # the 24 nn.Parameters are views of the flat buffer the optimizer updates
    assert all(p.data_ptr() >= tr.flat_param.data_ptr() for p in net.parameters())


def test_flat_adam_matches_torch_adam():
    """nb200_adam_step == torch.optim.Adam(lr=5e-4) (train.py:43) over several steps."""
    from nerf_simple_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(0)
    n = 100003
    p0 = torch.randn(n, device="cuda")
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=5e-4)
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for t in range(1, 6):
        g = torch.randn(n, device="cuda") * (0.1 * t)
        ref.grad = g.clone()
        opt.step()
        rc = lib.nb200_adam_step(_lib.ptr(p), _lib.ptr(g), _lib.ptr(m), _lib.ptr(v), n, t, 5e-4, 0.9, 0.999, 1e-8,
                                 _lib.stream_ptr())
        assert rc == 0
    assert float((p - ref.detach()).abs().max()) <= 1e-6


def test_select_rays_and_mse_kernels():
    """Device ray selection (train.py:47-49) and MSELoss + gradient (train.py:42,52-54) through the C ABI."""
    from nerf_simple_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(0)
    n, B = 100003, 8192
    table = torch.randn(n, 6, device="cuda")
    gt_table = torch.rand(n, 3, device="cuda")
    rays, gt = torch.empty(B, 6, device="cuda"), torch.empty(B, 3, device="cuda")
    ids = torch.empty(B, dtype=torch.int64, device="cuda")

    def select(seed, off, ids_t=ids):
        rc = lib.nb200_select_rays(_lib.ptr(table), _lib.ptr(gt_table), n, seed, off, B, _lib.ptr(rays), _lib.ptr(gt),
                                   _lib.ptr(ids_t), _lib.stream_ptr())
        assert rc == 0
        return ids_t.clone()

    a = select(3, 0)
    assert int(a.min()) >= 0 and int(a.max()) < n
    assert torch.equal(rays, table[a]) and torch.equal(gt, gt_table[a])          # gather == table[ids]
    assert torch.equal(select(3, 0), a) and not torch.equal(select(4, 0), a)       # keyed by (seed, offset)
    assert torch.equal(select(3, 100)[:-100], a[100:])                             # offset = position in the stream
    # uniform over the table: mean / variance of ids/n, and every decile populated evenly
    u = a.double() / n
    assert abs(float(u.mean()) - 0.5) < 0.02 and abs(float(u.var()) - 1 / 12) < 0.01
    hist = torch.histc(u.float(), bins=10, min=0, max=1)
    assert float(hist.min()) > 0.8 * B / 10 and float(hist.max()) < 1.2 * B / 10
    assert lib.nb200_select_rays(_lib.ptr(table), None, n, 1, 0, 0, None, None, None, _lib.stream_ptr()) == 0    # empty batch
    # MSE: loss and gradient against torch autograd
    rgb = torch.rand(B, 3, device="cuda", requires_grad=True)
    loss_ref = torch.nn.MSELoss()(rgb, gt)
    loss_ref.backward()
    d_rgb, loss = torch.empty(B, 3, device="cuda"), torch.zeros((), device="cuda")
    rc = lib.nb200_mse_loss_grad(_lib.ptr(rgb.detach()), _lib.ptr(gt), B, _lib.ptr(d_rgb), _lib.ptr(loss), _lib.stream_ptr())
    assert rc == 0
    assert abs(float(loss) - loss_ref.item()) <= 1e-6 * max(1.0, loss_ref.item())
    assert float((d_rgb - rgb.grad).abs().max()) <= 1e-9 + 1e-6 * float(rgb.grad.abs().max())


def test_bf16_vs_fp32_trained_render_psnr():
    """North-star acceptance for the bf16 mode on a TRAINED net: the same weights rendered in bf16 and in fp32
    give images whose PSNR against the ground truth differs by <= 0.1 dB, and training in bf16 tracks training
    in fp32 (same initial weights, same ray batches, same jitter)."""
    from nerf_simple_b200 import config, ops
    from nerf_simple_b200.nets import Nerf
    from nerf_simple_b200.rendering import render_nerf
    from nerf_simple_b200.xyz import poses_to_render
    poses = torch.stack(poses_to_render(4, -30, 6)).cuda()
    H = W = 24
    rays = ops.generate_rays(poses, H, W, 33.3)
    gt = torch.sigmoid(rays[:, 3:6] * 3)                      # a smooth, learnable colour field
    finals, nets = {}, {}
    for prec in ("fp32", "bf16"):
        config.set_precision(prec)
        config.set_sampler("philox", seed=5)                   # same device jitter stream for both runs
        torch.manual_seed(0)
        net = Nerf().cuda()
        opt = torch.optim.Adam(net.parameters(), lr=5e-4)
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        losses = []
        for _ in range(150):
            ids = torch.randint(0, rays.shape[0], (1024,), device="cuda", generator=g)
            opt.zero_grad()
            rgb = render_nerf(rays[ids], net, 32)[0]
            loss = torch.nn.functional.mse_loss(rgb, gt[ids])
            loss.backward()
            opt.step()
            losses.append(loss.item())
        finals[prec], nets[prec] = float(np.mean(losses[-20:])), net
    assert finals["fp32"] < 0.02                                                    # it did train
    assert abs(finals["bf16"] - finals["fp32"]) <= 0.1 * finals["fp32"]            # bf16 training tracks fp32
    # one trained net, rendered in both precisions with the same jitter: PSNR delta <= 0.1 dB
    psnr = {}
    with torch.no_grad():
        for prec in ("fp32", "bf16"):
            config.set_precision(prec)
            config.set_sampler("philox", seed=9)
            img = render_nerf(rays[:H * W], nets["fp32"], 64)[0].clamp(0, 1)
            psnr[prec] = float(-10 * torch.log10(torch.mean((img - gt[:H * W]) ** 2)))
    config.set_precision("bf16")
    config.set_sampler("reference")
    assert abs(psnr["bf16"] - psnr["fp32"]) <= 0.1, psnr


def test_trainer_graph_replay_matches_eager():
    """The CUDA-graph replayed step (device-resident Philox positions, Adam step count and lr) against the same
    step launched eagerly: same batches, same jitter, same optimizer schedule."""
    from nerf_simple_b200 import ops
    from nerf_simple_b200.nets import Nerf
    from nerf_simple_b200.trainer import Trainer
    from nerf_simple_b200.xyz import poses_to_render
    poses = torch.stack(poses_to_render(4, -30, 4)).cuda()
    rays = ops.generate_rays(poses, 32, 32, 44.4)
    gt = torch.sigmoid(rays[:, 3:6] * 3)
    runs = {}
    for use_graph in (False, True):
        torch.manual_seed(0)
        net = Nerf().cuda()
        tr = Trainer(net, rays, gt, N=32, batch_size=1024, lr_decay=0.99, use_graph=use_graph)
        losses = [tr.step(sync_loss=True) for _ in range(25)]
        state = tr._state.cpu()
        runs[use_graph] = (losses, state, tr)
    assert runs[True][2]._graphs and runs[True][2].launch_mode.startswith("one CUDA-graph")    # the capture succeeded and is in use
    assert not runs[False][2]._graphs
    # device-resident state after 25 steps: select offset, sampler offset (quads), step count, lr
    for losses, state, tr in runs.values():
        q = state[:24].view(torch.int64)
        assert int(q[0]) == 25 * 1024 and int(q[1]) == 25 * (1024 * 32 // 4) and int(q[2]) == 25
        assert abs(float(state[24:28].view(torch.float32)) - 5e-4 * 0.99 ** 25) < 1e-9
        assert abs(tr.lr - 5e-4 * 0.99 ** 25) < 1e-12 and tr.t == 25
    a, b = np.array(runs[False][0]), np.array(runs[True][0])
    assert a[0] == b[0]                                    # identical first step (same kernels, nothing accumulated yet)
    assert np.abs(a - b).max() <= 2e-2 * a.max()           # wgrad accumulates with atomics: trajectories agree closely
    assert b[-5:].mean() < 0.5 * b[:3].mean()


def test_trainer_says_why_the_graph_path_is_off():
    """A step that cannot be one CUDA-graph replay (ragged N, fp32 mode) still trains, eagerly, and says so."""
    import warnings
    from nerf_simple_b200 import ops
    from nerf_simple_b200.nets import Nerf
    from nerf_simple_b200.trainer import Trainer
    from nerf_simple_b200.xyz import poses_to_render
    torch.manual_seed(0)
    poses = torch.stack(poses_to_render(4, -30, 2)).cuda()
    rays = ops.generate_rays(poses, 24, 24, 33.3)
    gt = torch.sigmoid(rays[:, 3:6] * 3)
    tr = Trainer(Nerf().cuda(), rays, gt, N=32, batch_size=256, precision="bf16")
    assert tr.use_graph and tr.graph_off_reason is None
    tr.close()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        tr = Trainer(Nerf().cuda(), rays, gt, N=30, batch_size=256, precision="bf16")       # N % 4 != 0
    assert not tr.use_graph and "N=30" in tr.graph_off_reason and any("graph" in str(x.message) for x in w)
    losses = [tr.step(sync_loss=True) for _ in range(6)]
    assert all(np.isfinite(losses)) and not tr._graphs
    tr.close()
