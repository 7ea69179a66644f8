"""torch-CPU restatement of the reference's forward render path -- CPU BASELINE LEG ONLY.

The reference's own implementation is PyTorch (utils/rendering.py:13-85, utils/nets.py:34-43,
utils/xyz.py:6-36); it cannot travel to the GPU box, so this is the same op sequence written
against plain weight tensors, timed by bench.py's `cpu_baseline` / `--impl reference` legs with all
intra-op threads.  tests/test_oracle.py pins it to the reference-generated golden vectors.
"""
import torch
import torch.nn.functional as F


def _gamma(x, L):
    return torch.cat([fn((2 ** i) * x) for i in range(L) for fn in (torch.sin, torch.cos)], dim=1)


def _encode(v, Lp=10, Ld=4):
    c = [v[:, i:i + 1] for i in range(6)]
    posx = torch.cat(c[:3] + [_gamma(c[0], Lp), _gamma(c[1], Lp), _gamma(c[2], Lp)], dim=1)
    posd = torch.cat(c[3:] + [_gamma(c[3], Ld), _gamma(c[4], Ld), _gamma(c[5], Ld)], dim=1)
    return posx, posd


def mlp(v, P):
    posx, posd = _encode(v)
    h = posx
    for i in (0, 2, 4, 6, 8):
        h = F.relu(F.linear(h, P[f"layers_0.{i}.weight"], P[f"layers_0.{i}.bias"]))
    h = F.relu(F.linear(torch.cat([h, posx], 1), P["skip_conn_layer.0.weight"], P["skip_conn_layer.0.bias"]))
    for i in (0, 2):
        h = F.relu(F.linear(h, P[f"layers_1.{i}.weight"], P[f"layers_1.{i}.bias"]))
    sigma = F.linear(h, P["sigma_fc.0.weight"], P["sigma_fc.0.bias"])
    g = F.linear(h, P["layers_2.weight"], P["layers_2.bias"])
    c1 = F.relu(F.linear(torch.cat([g, posd], 1), P["color_fc.0.weight"], P["color_fc.0.bias"]))
    return torch.cat([F.linear(c1, P["color_fc.2.weight"], P["color_fc.2.bias"]), sigma], 1)


def render_nerf(rays, P, N, u, tn=2.0, tf=6.0):
    B = rays.shape[0]
    bins = torch.linspace(tn, tf, N + 1)
    ts = (bins[1] - bins[0]) * u + bins[:-1]
    o, d = rays[:, :3], rays[:, 3:]
    locs = o.unsqueeze(-1) + d.unsqueeze(-1) * ts.unsqueeze(1)
    dn = d / torch.norm(d, dim=1, keepdim=True)
    q = torch.cat((locs, dn.unsqueeze(-1).expand(-1, -1, N)), dim=1).permute(0, 2, 1).reshape(-1, 6)
    out = mlp(q, P).reshape(B, N, 4)
    deltas = torch.cat((ts[:, 1:] - ts[:, :-1], 1e10 * torch.ones_like(ts[:, :1])), dim=1)
    deltas = deltas * torch.norm(dn[..., None, :], dim=-1)
    alpha = 1 - torch.exp(-F.softplus(out[..., 3]) * deltas)
    T = torch.cumprod(torch.cat([torch.ones((B, 1)), 1. - alpha + 1e-10], -1), -1)[:, :-1]
    w = alpha * T
    rgb = torch.sum(w.unsqueeze(-1) * out[..., :3], dim=1)
    depth = torch.sum(w * ts, dim=-1)
    acc = torch.sum(w, dim=-1)
    disp = 1. / torch.max(1e-10 * torch.ones_like(depth), depth / acc)
    return rgb, disp, alpha, acc, w
