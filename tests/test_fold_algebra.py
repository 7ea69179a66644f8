"""CPU check of the algebra behind the default bf16 chain (nerf_simple_b200/csrc/mlp_tc.cu, top):
layers_2 has no activation (utils/nets.py:41-42), so it folds into color_fc.0, and the gradients of both layers
un-fold exactly from Gm = delta_c1^T h7 and s = column sums of delta_c1.  Checked in float64 against the oracle's
layer-by-layer forward and backward (oracle/nerf_oracle.py: mlp_forward / mlp_backward), i.e. against the reference's
own order of operations.  The GPU tests then check the kernels that implement these formulas against the goldens."""
import numpy as np

from conftest import load_golden
from oracle import nerf_oracle as O


def test_fold_and_unfold_are_exact():
    P = {k: v.astype(np.float64) for k, v in load_golden("weights_seed0.npz").items()}
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.uniform(-1, 1, (512, 3)), rng.normal(size=(512, 3))], axis=1)
    v[:, 3:] /= np.linalg.norm(v[:, 3:], axis=1, keepdims=True)
    out, saved = O.mlp_forward(v, P, dtype=np.float64, keep=True)
    h7, posd, c1 = saved["acts"][7], saved["posd"], saved["c1"]
    Wc0, bc0, Wg, bg = P["color_fc.0.weight"], P["color_fc.0.bias"], P["layers_2.weight"], P["layers_2.bias"]

    # forward: fold_weights_kernel
    Wf = Wc0[:, :256] @ Wg
    bf = Wc0[:, :256] @ bg + bc0
    c1_fold = np.maximum(h7 @ Wf.T + posd @ Wc0[:, 256:].T + bf, 0)
    assert np.max(np.abs(c1_fold - c1)) <= 1e-12

    # backward: DgradEpi<true> (delta_h7 from delta_c1 in one layer) and fold_grads_kernel
    d_out = rng.normal(size=out.shape)
    G = O.mlp_backward(d_out, saved, P, dtype=np.float64)
    d_c1 = (d_out[:, :3] @ P["color_fc.2.weight"]) * (c1 > 0)
    Gm = d_c1.T @ h7                     # the one wgrad item that replaces (delta_c1, g) and (delta_g, h7)
    s = d_c1.sum(0)
    dWc0 = np.concatenate([Gm @ Wg.T + np.outer(s, bg), d_c1.T @ posd], axis=1)
    assert np.max(np.abs(dWc0 - G["color_fc.0.weight"])) <= 1e-10
    assert np.max(np.abs(s - G["color_fc.0.bias"])) <= 1e-12
    assert np.max(np.abs(Wc0[:, :256].T @ Gm - G["layers_2.weight"])) <= 1e-10
    assert np.max(np.abs(Wc0[:, :256].T @ s - G["layers_2.bias"])) <= 1e-10
    # the delta that continues down the chain
    d_h7_fold = (d_c1 @ Wf + d_out[:, 3:4] @ P["sigma_fc.0.weight"]) * (h7 > 0)
    d_g = d_c1 @ Wc0[:, :256]
    d_h7_ref = (d_g @ Wg + d_out[:, 3:4] @ P["sigma_fc.0.weight"]) * (h7 > 0)
    assert np.max(np.abs(d_h7_fold - d_h7_ref)) <= 1e-12


def test_both_bf16_modes_share_buffer_sizes():
    """NB200_BF16 (folded) and NB200_BF16_LAYERWISE use one packed image and the same saved / scratch layouts."""
    from nerf_simple_b200 import _lib
    lib = _lib.load()
    assert lib.nb200_packed_weights_bytes(_lib.BF16) == lib.nb200_packed_weights_bytes(_lib.BF16_LAYERWISE) > 0
    for M in (128, 4096 * 64):
        assert lib.nb200_mlp_saved_bytes(_lib.BF16, M) == lib.nb200_mlp_saved_bytes(_lib.BF16_LAYERWISE, M) > 0
        assert lib.nb200_mlp_scratch_bytes(_lib.BF16, M, 1) == lib.nb200_mlp_scratch_bytes(_lib.BF16_LAYERWISE, M, 1) > 0
    assert lib.nb200_packed_weights_bytes(7) == 0
