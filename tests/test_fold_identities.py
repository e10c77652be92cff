"""The algebraic identities behind the host-side folds of the tile kernel (multimodal-flows_b200/csrc/tftile_model.cu: fold_ln,
load_block, score_bound), checked in fp64 / fp32 torch on the CPU.  The kernels themselves are held to
the reference goldens by the GPU tests; this file pins the mathematics they rely on."""
import math

import torch


def _ln(x, eps=1e-5):
    m = x.mean(-1, keepdim=True)
    v = ((x - m) ** 2).mean(-1, keepdim=True)
    return (x - m) / torch.sqrt(v + eps)


def test_layernorm_affine_folds_into_the_next_linear():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(37, 256, generator=g, dtype=torch.float64) * 3 + 1
    gam, beta = torch.randn(256, generator=g, dtype=torch.float64), torch.randn(256, generator=g, dtype=torch.float64)
    W, b = torch.randn(768, 256, generator=g, dtype=torch.float64), torch.randn(768, generator=g, dtype=torch.float64)
    ref = (_ln(x) * gam + beta) @ W.T + b
    folded = _ln(x) @ (W * gam).T + (b + W @ beta)                      # W' = W diag(g), b' = b + W beta
    assert torch.allclose(ref, folded, rtol=1e-12, atol=1e-10)


def test_softmax_needs_no_maximum_inside_the_score_bound():
    """q, k = per-head LayerNorm outputs: |q.k| <= (max|g_q| sqrt(hs) + |b_q|)(max|g_k| sqrt(hs) + |b_k|); inside the bound the
    unshifted base-2 exponentials are finite in fp32 and the normalised probabilities equal the shifted ones."""
    g = torch.Generator().manual_seed(1)
    for hs in (32, 64):
        gq, bq = torch.randn(hs, generator=g) * 1.5, torch.randn(hs, generator=g) * 0.5
        gk, bk = torch.randn(hs, generator=g) * 1.5, torch.randn(hs, generator=g) * 0.5
        c = math.log2(math.e) / math.sqrt(hs)
        bound = (gq.abs().max() * math.sqrt(hs) + bq.norm()) * (gk.abs().max() * math.sqrt(hs) + bk.norm()) * c
        q = _ln(torch.randn(128, hs, generator=g) * 7) * gq + bq
        k = _ln(torch.randn(128, hs, generator=g) * 7) * gk + bk
        s2 = (q * c) @ k.T                                             # the scale folded into q: scores in log2 units
        assert float(s2.abs().max()) <= float(bound)
        if float(bound) <= 64:
            p_raw = torch.exp2(s2)
            assert torch.isfinite(p_raw).all() and float(p_raw.sum(-1).min()) > 0
            p_shift = torch.exp2(s2 - s2.max(-1, keepdim=True).values)
            a, b = p_raw / p_raw.sum(-1, keepdim=True), p_shift / p_shift.sum(-1, keepdim=True)
            assert torch.allclose(a, b, rtol=2e-5, atol=1e-7)
            assert torch.allclose(a, torch.softmax(q @ k.T / math.sqrt(hs), -1), rtol=1e-4, atol=1e-6)


def test_v_bias_travels_into_the_projection_bias():
    g = torch.Generator().manual_seed(2)
    P = torch.softmax(torch.randn(50, 50, generator=g, dtype=torch.float64), -1)
    V, bv = torch.randn(50, 64, generator=g, dtype=torch.float64), torch.randn(64, generator=g, dtype=torch.float64)
    Wp, bp = torch.randn(256, 64, generator=g, dtype=torch.float64), torch.randn(256, generator=g, dtype=torch.float64)
    ref = (P @ (V + bv)) @ Wp.T + bp
    folded = (P @ V) @ Wp.T + (bp + Wp @ bv)                           # rows of P sum to one
    assert torch.allclose(ref, folded, rtol=1e-12, atol=1e-10)


def test_half_of_gelu_moves_into_the_next_weights_exactly():
    g = torch.Generator().manual_seed(3)
    z = torch.randn(64, 512, generator=g)
    W = torch.randn(256, 512, generator=g).bfloat16()
    h = torch.nn.functional.gelu(z)
    assert torch.equal((W.float() * 0.5).bfloat16().float(), W.float() * 0.5)        # a power of two: no rounding in bf16
    ref = h.double() @ W.double().T
    folded = (2.0 * h).double() @ (W.float() * 0.5).double().T
    assert torch.allclose(ref, folded, rtol=1e-12, atol=1e-9)


def test_centred_head_rows_make_the_per_head_layernorm_mean_free():
    """LN over a head ignores a constant added to all of the head's outputs: with the head's rows of W (and biases) centred the
    pre-activations sum to zero, and LN = y * rsqrt(mean(y^2) + eps) * g + b equals the LN of the un-centred projection."""
    g = torch.Generator().manual_seed(4)
    hs, C = 64, 256
    x = torch.randn(50, C, generator=g, dtype=torch.float64)
    W, b = torch.randn(hs, C, generator=g, dtype=torch.float64) + 0.7, torch.randn(hs, generator=g, dtype=torch.float64) + 2.0
    gam, beta = torch.randn(hs, generator=g, dtype=torch.float64), torch.randn(hs, generator=g, dtype=torch.float64)
    ref = _ln(x @ W.T + b) * gam + beta
    Wc, bc = W - W.mean(0, keepdim=True), b - b.mean()
    y = x @ Wc.T + bc
    assert float(y.sum(-1).abs().max()) < 1e-9
    out = y * torch.rsqrt((y * y).mean(-1, keepdim=True) + 1e-5) * gam + beta
    assert torch.allclose(ref, out, rtol=1e-10, atol=1e-10)
