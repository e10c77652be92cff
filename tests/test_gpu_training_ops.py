"""Operators of the training step (include/mmf_b200_train.h) one by one against plain fp32 torch on the same inputs.
Tolerances: fp32 kernels 1e-5 relative; kernels with bf16 operands / outputs are compared with torch evaluated on the SAME
bf16-rounded operands, so what remains is accumulation order and the bf16 rounding of the result (2^-8 relative)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def ops():
    from mmf_b200._train_abi import Ops
    return Ops(torch.device(DEV))


def bf(x):
    return x.to(torch.bfloat16)


def jets(ns):
    import numpy as np
    n = np.array(ns)
    off = np.zeros(len(ns) + 1, np.int32)
    off[1:] = np.cumsum(n)
    poff = np.zeros(len(ns) + 1, np.int64)
    poff[1:] = np.cumsum(n * n)
    row_jet = np.repeat(np.arange(len(ns), dtype=np.int32), n)
    T = lambda a: torch.from_numpy(a).to(DEV)
    return T(off), T(poff), T(row_jet), int(n.sum()), int(n.max())


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 384, 128), (1000, 128, 512), (77, 256, 768), (13819, 512, 256)])
@pytest.mark.parametrize("mode", [0, 1])
def test_gemm_store_modes(ops, M, N, K, mode):
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    A, B = bf(torch.randn(M, K, device=DEV, generator=g)), bf(torch.randn(N, K, device=DEV, generator=g) * 0.1)
    bias = torch.randn(N, device=DEV, generator=g)
    C = torch.full((M, N), 7.0, device=DEV, dtype=torch.bfloat16 if mode == 0 else torch.float32)
    ops.gemm(A, B, C, bias, mode)
    want = A.float() @ B.float().T + bias
    assert rel(C.float(), want) < (4e-3 if mode == 0 else 2e-6), rel(C.float(), want)


@pytest.mark.parametrize("M,N,K", [(300, 512, 128), (13819, 512, 256), (1000, 256, 512)])
def test_gemm_gelu_epilogues(ops, M, N, K):
    """mode 3: z = x W^T + b and h = GELU(z) from one epilogue (h is the GELU of the bf16 z the backward pass reads);
    mode 4: dz = (dh W) * GELU'(z) with the z tile loaded next to the accumulator"""
    g = torch.Generator(device=DEV).manual_seed(M + K)
    A, B = bf(torch.randn(M, K, device=DEV, generator=g)), bf(torch.randn(N, K, device=DEV, generator=g) * 0.1)
    bias = torch.randn(N, device=DEV, generator=g)
    z, h = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16), torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, B, z, bias, 3, aux=h)
    want = A.float() @ B.float().T + bias
    assert rel(z.float(), want) < 4e-3
    assert rel(h.float(), torch.nn.functional.gelu(z.float())) < 3e-3
    zin = bf(torch.randn(M, N, device=DEV, generator=g) * 1.5)
    dz = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, B, dz, None, 4, aux=zin)
    zz = zin.float().clone().requires_grad_(True)
    torch.nn.functional.gelu(zz).backward(A.float() @ B.float().T)
    assert rel(dz.float(), zz.grad) < 4e-3, rel(dz.float(), zz.grad)


@pytest.mark.parametrize("C,H", [(128, 4), (256, 4)])
def test_gemm_qkv_with_head_layernorm_epilogue(ops, C, H):
    """c_attn + q / k LayerNorm in one kernel: qkv as the plain bf16 product, qkn = LayerNorm over each head of those bf16 values"""
    g = torch.Generator(device=DEV).manual_seed(C + 3)
    M, hs = 1000, C // H
    A, W, bias = bf(torch.randn(M, C, device=DEV, generator=g)), bf(torch.randn(3 * C, C, device=DEV, generator=g) * 0.2), torch.randn(3 * C, device=DEV, generator=g)
    qg, qb, kg, kb = (torch.rand(hs, device=DEV, generator=g) + 0.5 for _ in range(4))
    qkv, qkn = torch.zeros(M, 3 * C, device=DEV, dtype=torch.bfloat16), torch.zeros(M, 2 * C, device=DEV, dtype=torch.bfloat16)
    ops.gemm_qkv(A, W, bias, qkv, qkn, H, qg, qb, kg, kb)
    want = A.float() @ W.float().T + bias
    assert rel(qkv.float(), want) < 4e-3
    ln = lambda x, gam, bet: torch.nn.functional.layer_norm(x.float().view(M, H, hs), (hs,), gam, bet, 1e-5).reshape(M, C)
    assert rel(qkn[:, :C].float(), ln(qkv[:, :C], qg, qb)) < 4e-3 and rel(qkn[:, C:].float(), ln(qkv[:, C:2 * C], kg, kb)) < 4e-3
    # and it equals the two-kernel path bit for bit on the stored bf16 values
    qn2, kn2 = torch.empty(M, C, device=DEV, dtype=torch.bfloat16), torch.empty(M, C, device=DEV, dtype=torch.bfloat16)
    ops.qkln_fwd(qkv, C, H, qg, qb, kg, kb, qn2, kn2)
    assert rel(qkn[:, :C].float(), qn2.float()) < 1e-3 and rel(qkn[:, C:].float(), kn2.float()) < 1e-3


def test_gemm_residual_epilogue(ops):
    """mode 5: out = resid + x W^T + b + tadd[row_jet], out of place, on column slices of 256-wide fp32 buffers"""
    g = torch.Generator(device=DEV).manual_seed(11)
    off, _, row_jet, M, _ = jets([5, 1, 150, 44, 77])
    A, W, bias = bf(torch.randn(M, 512, device=DEV, generator=g)), bf(torch.randn(128, 512, device=DEV, generator=g) * 0.1), torch.randn(128, device=DEV, generator=g)
    R, tadd = torch.randn(M, 256, device=DEV, generator=g), torch.randn(5, 256, device=DEV, generator=g)
    out = torch.zeros(M, 256, device=DEV)
    ops.gemm(A, W, out[:, 128:], bias, 5, resid=R[:, 128:], tadd=tadd[:, 128:], row_jet=row_jet)
    want = R[:, 128:] + A.float() @ W.float().T + bias + tadd[row_jet.long()][:, 128:]
    assert rel(out[:, 128:], want) < 2e-6 and float(out[:, :128].abs().max()) == 0.0
    ops.gemm(A, W, out[:, :128], None, 5, resid=R[:, :128])
    assert rel(out[:, :128], R[:, :128] + A.float() @ W.float().T) < 2e-6


def test_gemm_strided_views_and_no_bias(ops):
    g = torch.Generator(device=DEV).manual_seed(5)
    big_a, big_c = bf(torch.randn(500, 256, device=DEV, generator=g)), torch.zeros(500, 1024, device=DEV, dtype=torch.bfloat16)
    W = bf(torch.randn(512, 128, device=DEV, generator=g) * 0.1)
    ops.gemm(big_a[:, 128:], W, big_c[:, 512:], None, 0)
    want = big_a[:, 128:].float() @ W.float().T
    assert rel(big_c[:, 512:].float(), want) < 4e-3
    assert float(big_c[:, :512].abs().max()) == 0.0


@pytest.mark.parametrize("M,N,K,ks", [(384, 128, 13819, 32), (128, 256, 1000, 4), (512, 128, 130, 1), (256, 256, 5000, 100)])
def test_gemm_reduce_add_split_k(ops, M, N, K, ks):
    """weight gradient: dW += dy^T x with tokens on the K axis, ragged K, operands as transposed [C, Mp] arrays"""
    g = torch.Generator(device=DEV).manual_seed(K)
    Kp = (K + 63) // 64 * 64
    A, B = bf(torch.randn(M, Kp, device=DEV, generator=g)), bf(torch.randn(N, Kp, device=DEV, generator=g))
    C0 = torch.randn(M, N, device=DEV, generator=g)
    C = C0.clone()
    ops.gemm(A[:, :K], B[:, :K], C, None, 2, ks)
    want = C0 + A[:, :K].float() @ B[:, :K].float().T
    assert rel(C, want) < 2e-5, rel(C, want)


@pytest.mark.parametrize("M,N,K,ks", [(384, 128, 13819, 32), (128, 256, 1000, 4), (512, 128, 130, 1), (256, 512, 5000, 100), (128, 128, 64, 1)])
def test_gemm_tn_weight_gradient_from_row_major_operands(ops, M, N, K, ks):
    """dW += dy^T x with dy [tokens, out], x [tokens, in] as they lie in memory (MN-major tcgen05 operands), column slices included"""
    g = torch.Generator(device=DEV).manual_seed(K + 1)
    dy, x = bf(torch.randn(K, M + 128, device=DEV, generator=g)), bf(torch.randn(K, N + 64, device=DEV, generator=g))
    dyv, xv = dy[:, 128:], x[:, :N]
    C0 = torch.randn(M, N, device=DEV, generator=g)
    C = C0.clone()
    ops.gemm_tn(dyv, xv, C, ks)
    want = C0 + dyv.float().T @ xv.float()
    assert rel(C, want) < 2e-5, rel(C, want)


@pytest.mark.parametrize("f32", [True, False])
def test_cast_transpose_colsum(ops, f32):
    g = torch.Generator(device=DEV).manual_seed(1)
    rows, cols, Mp = 1001, 384, 1024
    x = torch.randn(rows, 512, device=DEV, generator=g)
    x = x if f32 else bf(x)
    xv = x[:, 64:64 + cols]
    out, outT, cs = torch.zeros(rows, cols, device=DEV, dtype=torch.bfloat16), torch.zeros(cols, Mp, device=DEV, dtype=torch.bfloat16), torch.ones(cols, device=DEV)
    ops.cast_transpose(xv, out, outT, cs)
    assert torch.equal(out, bf(xv)) and torch.equal(outT[:, :rows], bf(xv).T) and float(outT[:, rows:].abs().max()) == 0
    assert rel(cs, 1.0 + xv.float().sum(0)) < 1e-5


def test_sgemm_general_strides(ops):
    g = torch.Generator(device=DEV).manual_seed(2)
    A, B, bias = torch.randn(37, 50, device=DEV, generator=g), torch.randn(50, 29, device=DEV, generator=g), torch.randn(29, device=DEV, generator=g)
    C = torch.zeros(37, 29, device=DEV)
    ops.sgemm(A, 50, 1, B, 29, 1, C, 37, 29, 50, bias=bias)
    assert rel(C, A @ B + bias) < 1e-5
    At = torch.zeros(50, 50, device=DEV)
    ops.sgemm(A, 1, 50, A, 50, 1, At, 50, 50, 37, accumulate=True)                       # A^T A through strides
    assert rel(At, A.T @ A) < 1e-5


@pytest.mark.parametrize("C", [128, 256])
def test_layernorm_forward_backward(ops, C):
    g = torch.Generator(device=DEV).manual_seed(C)
    off, _, row_jet, M, _ = jets([5, 1, 150, 44])
    x, add = torch.randn(M, 256, device=DEV, generator=g) * 2 + 0.5, torch.randn(M, 256, device=DEV, generator=g)
    gam, bet = torch.rand(C, device=DEV, generator=g) + 0.5, torch.randn(C, device=DEV, generator=g)
    tadd = torch.randn(4, 256, device=DEV, generator=g)
    xv, av, tv = x[:, 256 - C:], add[:, 256 - C:], tadd[:, 256 - C:]
    out16, out32 = torch.empty(M, C, device=DEV, dtype=torch.bfloat16), torch.empty(M, C, device=DEV)
    mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    ops.ln_fwd(xv, gam, bet, mean, rstd, add=av, tadd=tv, row_jet=row_jet, out16=out16, out32=out32)
    xs = (xv + av).clone().requires_grad_(True)
    gp, bp = gam.clone().requires_grad_(True), bet.clone().requires_grad_(True)
    y = torch.nn.functional.layer_norm(xs, (C,), gp, bp, 1e-5)
    want = y + tv[row_jet.long()]
    assert rel(out32, want) < 2e-6 and torch.equal(out16, bf(out32))
    dy = torch.randn(M, C, device=DEV, generator=g)
    y.backward(dy)
    dx0 = torch.randn(M, 256, device=DEV, generator=g)
    dx, dg, db = dx0.clone(), torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    dx16, dxsum = torch.zeros(M, 256, device=DEV, dtype=torch.bfloat16), torch.ones(C, device=DEV)
    ops.ln_bwd(dy, xv, mean, rstd, gam, dx[:, :C], dg, db, add=av, accumulate=True, dx16=dx16[:, 256 - C:], dxsum=dxsum)
    assert rel(dx[:, :C] - dx0[:, :C], xs.grad) < 1e-5 and torch.equal(dx[:, C:], dx0[:, C:])
    assert rel(dg, gp.grad) < 1e-5 and rel(db, bp.grad) < 1e-5
    # the hand-off to the next linear: bf16 copy of the accumulated gradient and its column sums
    assert torch.equal(dx16[:, 256 - C:], bf(dx[:, :C])) and rel(dxsum, 1.0 + dx[:, :C].sum(0)) < 1e-5


@pytest.mark.parametrize("C,H", [(128, 4), (256, 4)])
def test_qk_layernorm_forward_backward(ops, C, H):
    g = torch.Generator(device=DEV).manual_seed(C)
    M, hs = 333, C // H
    qkv = bf(torch.randn(M, 3 * C, device=DEV, generator=g) * 1.5)
    qg, qb, kg, kb = (torch.rand(hs, device=DEV, generator=g) + 0.5 for _ in range(4))
    qn, kn = torch.empty(M, C, device=DEV, dtype=torch.bfloat16), torch.empty(M, C, device=DEV, dtype=torch.bfloat16)
    ops.qkln_fwd(qkv, C, H, qg, qb, kg, kb, qn, kn)
    src = qkv.float().clone().requires_grad_(True)
    ps = [p.clone().requires_grad_(True) for p in (qg, qb, kg, kb)]
    q = torch.nn.functional.layer_norm(src[:, :C].view(M, H, hs), (hs,), ps[0], ps[1], 1e-5).reshape(M, C)
    k = torch.nn.functional.layer_norm(src[:, C:2 * C].view(M, H, hs), (hs,), ps[2], ps[3], 1e-5).reshape(M, C)
    assert rel(qn.float(), q) < 4e-3 and rel(kn.float(), k) < 4e-3
    d = bf(torch.randn(M, 3 * C, device=DEV, generator=g))
    (q * d[:, :C].float()).sum().backward(retain_graph=True)
    (k * d[:, C:2 * C].float()).sum().backward()
    dq = d.clone()
    grads = [torch.zeros(hs, device=DEV) for _ in range(4)]
    ops.qkln_bwd(dq, qkv, C, H, qg, kg, grads[0], grads[1], grads[2], grads[3])
    assert rel(dq[:, :2 * C].float(), src.grad[:, :2 * C]) < 5e-3 and torch.equal(dq[:, 2 * C:], d[:, 2 * C:])
    for a, p in zip(grads, ps):
        assert rel(a, p.grad) < 1e-4


@pytest.mark.parametrize("hs,H", [(32, 4), (64, 4)])
def test_attention_forward_backward(ops, hs, H):
    g = torch.Generator(device=DEV).manual_seed(hs)
    ns = [3, 150, 1, 64, 129, 17]
    off, poff, row_jet, M, nmax = jets(ns)
    C = hs * H
    qn, kn = bf(torch.randn(M, C, device=DEV, generator=g)), bf(torch.randn(M, C, device=DEV, generator=g))
    qkv = bf(torch.randn(M, 3 * C, device=DEV, generator=g))
    v = qkv[:, 2 * C:]
    o, P = torch.empty(M, C, device=DEV, dtype=torch.bfloat16), torch.empty(int(poff[-1]) * H, device=DEV, dtype=torch.bfloat16)
    ops.attn_fwd(qn, kn, v, off, poff, len(ns), H, hs, nmax, o, P)
    qf, kf, vf = (t.float().clone().requires_grad_(True) for t in (qn, kn, v))
    outs = []
    for b, n in enumerate(ns):
        r = slice(int(off[b]), int(off[b + 1]))
        qq, kk, vv = (t[r].view(n, H, hs).transpose(0, 1) for t in (qf, kf, vf))
        outs.append(torch.nn.functional.scaled_dot_product_attention(qq, kk, vv).transpose(0, 1).reshape(n, C))
    want = torch.cat(outs)
    assert rel(o.float(), want) < 4e-3, rel(o.float(), want)
    dO = bf(torch.randn(M, C, device=DEV, generator=g))
    want.backward(dO.float())
    dqkv = torch.zeros(M, 3 * C, device=DEV, dtype=torch.bfloat16)
    ops.attn_bwd(dO, o, P, qn, kn, v, off, poff, len(ns), H, hs, nmax, dqkv, C)
    # P and O are kept in bf16 between the passes: 2^-8-sized relative errors on the gradients
    assert rel(dqkv[:, 2 * C:].float(), vf.grad) < 8e-3, rel(dqkv[:, 2 * C:].float(), vf.grad)
    assert rel(dqkv[:, :C].float(), qf.grad) < 2e-2, rel(dqkv[:, :C].float(), qf.grad)
    assert rel(dqkv[:, C:2 * C].float(), kf.grad) < 2e-2, rel(dqkv[:, C:2 * C].float(), kf.grad)


def test_gelu_add_jetsum_time_embed(ops):
    from oracle import mmf_oracle as orc
    g = torch.Generator(device=DEV).manual_seed(3)
    off, _, row_jet, M, _ = jets([5, 1, 150, 44])
    z = bf(torch.randn(M, 512, device=DEV, generator=g) * 2)
    h = torch.empty_like(z)
    ops.gelu_fwd(z, h)
    assert rel(h.float(), torch.nn.functional.gelu(z.float())) < 3e-3
    dh = bf(torch.randn(M, 512, device=DEV, generator=g))
    zz = z.float().clone().requires_grad_(True)
    torch.nn.functional.gelu(zz).backward(dh.float())
    dz = torch.empty_like(z)
    ops.gelu_bwd(dh, z, dz)
    assert rel(dz.float(), zz.grad) < 3e-3
    zf = torch.randn(7, 33, device=DEV, generator=g)
    hf, dzf = torch.empty_like(zf), torch.empty_like(zf)
    ops.gelu_fwd(zf, hf)
    assert rel(hf, torch.nn.functional.gelu(zf)) < 1e-6
    zz = zf.clone().requires_grad_(True)
    torch.nn.functional.gelu(zz).backward(torch.ones_like(zf))
    ops.gelu_bwd(torch.ones_like(zf), zf, dzf)
    assert rel(dzf, zz.grad) < 1e-5
    a, y, tadd = torch.randn(M, 256, device=DEV, generator=g), torch.randn(M, 128, device=DEV, generator=g), torch.randn(4, 256, device=DEV, generator=g)
    out = torch.zeros(M, 256, device=DEV)
    ops.add(out[:, 128:], a[:, 128:], y, tadd[:, 128:], row_jet)
    assert torch.allclose(out[:, 128:], a[:, 128:] + y + tadd[row_jet.long()][:, 128:], atol=1e-6) and float(out[:, :128].abs().max()) == 0
    js = torch.ones(4, 256, device=DEV)
    ops.jet_sum(a, off, 4, js, accumulate=True)
    want = torch.stack([a[int(off[b]):int(off[b + 1])].sum(0) for b in range(4)]) + 1
    assert rel(js, want) < 1e-5
    t = torch.rand(9, device=DEV, generator=g)
    emb = torch.zeros(9, 256, device=DEV)
    ops.time_embed(t, 128, True, emb)
    ref = orc.timestep_embedding(t.cpu(), 128).to(DEV)
    assert torch.allclose(emb[:, :128], ref, atol=2e-6) and torch.equal(emb[:, 128:], emb[:, :128])
    perm = torch.tensor([3, 0, 8, 1, 2, 7, 6, 5, 4], device=DEV, dtype=torch.int32)
    ops.time_embed(t, 128, False, emb, perm=perm)
    assert torch.allclose(emb[:, :128], ref[perm.long()], atol=2e-6)


def test_embeddings_forward_backward(ops):
    g = torch.Generator(device=DEV).manual_seed(4)
    M, E, V = 777, 256, 9
    xs, ks = torch.randn(M, 3, device=DEV, generator=g), torch.randint(0, V, (M,), device=DEV, generator=g, dtype=torch.int32)
    w0, b0, emb = (torch.randn(E, 3, device=DEV, generator=g), torch.randn(E, device=DEV, generator=g), torch.randn(V, E, device=DEV, generator=g))
    h, gg = torch.empty(M, E, device=DEV, dtype=torch.bfloat16), torch.empty(M, E, device=DEV, dtype=torch.bfloat16)
    ops.embed_x_fwd(xs, w0, b0, h)
    ops.embed_y_fwd(ks, emb, gg)
    wp, bp, ep = (t.clone().requires_grad_(True) for t in (w0, b0, emb))
    hx = torch.nn.functional.gelu(xs @ wp.T + bp)
    hy = torch.nn.functional.gelu(ep[ks.long()])
    assert rel(h.float(), hx) < 3e-3 and rel(gg.float(), hy) < 3e-3
    d = bf(torch.randn(M, E, device=DEV, generator=g))
    (hx * d.float()).sum().backward()
    (hy * d.float()).sum().backward()
    dw, db, de = torch.zeros_like(w0), torch.zeros_like(b0), torch.zeros_like(emb)
    ops.embed_x_bwd(d, xs, w0, b0, dw, db)
    ops.embed_y_bwd(d, ks, emb, de)
    assert rel(dw, wp.grad) < 1e-4 and rel(db, bp.grad) < 1e-4 and rel(de, ep.grad) < 1e-4


def test_heads_and_loss_forward_backward(ops):
    g = torch.Generator(device=DEV).manual_seed(6)
    ns = [5, 1, 150, 44, 9]
    off, _, row_jet, M, _ = jets(ns)
    B, I, V = len(ns), 512, 9
    z = bf(torch.randn(M, 2 * I, device=DEV, generator=g))
    h = torch.empty_like(z)
    ops.gelu_fwd(z, h)
    wx, bx, wy, by = (torch.randn(3, I, device=DEV, generator=g) * 0.1, torch.randn(3, device=DEV, generator=g),
                      torch.randn(V, I, device=DEV, generator=g) * 0.1, torch.randn(V, device=DEV, generator=g))
    vt, logits = torch.empty(M, 3, device=DEV), torch.empty(M, V, device=DEV)
    ops.head_fwd(h, I, wx, bx, wy, by, vt, logits)
    zp = z.float().clone().requires_grad_(True)
    ps = [t.clone().requires_grad_(True) for t in (wx, bx, wy, by)]
    hp = torch.nn.functional.gelu(zp)
    hq = hp + (h.float() - hp).detach()                # forward value = the bf16 h the kernel saw; gradient through GELU(z)
    vt_w, lg_w = hq[:, :I] @ ps[0].T + ps[1], hq[:, I:] @ ps[2].T + ps[3]
    assert rel(vt, vt_w) < 1e-5 and rel(logits, lg_w) < 1e-5
    # loss: per-jet masked MSE + CE(ignore_index 0), time-weighted combination
    tgt, k1 = torch.randn(M, 3, device=DEV, generator=g), torch.randint(0, V, (M,), device=DEV, generator=g, dtype=torch.int32)
    u = torch.randn(B, 2, device=DEV, generator=g) * 0.3
    l1, l2 = torch.empty(B, device=DEV), torch.empty(B, device=DEV)
    ops.loss_fwd(vt, logits, tgt, k1, off, B, V, l1, l2)
    up = u.clone().requires_grad_(True)
    n = torch.tensor(ns, device=DEV).float().clamp_min(1)
    rj = row_jet.long()
    mse = torch.zeros(B, device=DEV).index_add(0, rj, ((vt_w - tgt) ** 2).sum(1)) / n
    ce = torch.zeros(B, device=DEV).index_add(0, rj, torch.nn.functional.cross_entropy(lg_w, k1.long(), ignore_index=0, reduction="none")) / n
    assert rel(l1, mse) < 1e-5 and rel(l2, ce) < 1e-5
    w1, w2 = torch.exp(-up[:, 0]), torch.exp(-up[:, 1])
    loss = (0.5 * (up[:, 0] + w1 * mse) + 0.5 * (up[:, 1] + w2 * ce)).mean()
    out5, gl1, gl2, du = torch.empty(5, device=DEV), torch.empty(B, device=DEV), torch.empty(B, device=DEV), torch.empty(B, 2, device=DEV)
    ops.loss_combine(l1, l2, u, out5, gl1, gl2, du)
    assert abs(float(out5[0]) - float(loss)) < 1e-5 * abs(float(loss)) and abs(float(out5[3]) - float(w1.mean())) < 1e-5
    loss.backward()
    assert rel(du, up.grad) < 1e-5
    dvt, dlog = torch.empty(M, 3, device=DEV), torch.empty(M, V, device=DEV)
    ops.loss_bwd(vt, logits, tgt, k1, row_jet, off, gl1, gl2, V, dvt, dlog)
    dz, grads = torch.empty_like(z), [torch.zeros_like(t) for t in (wx, bx, wy, by)]
    ops.head_bwd(dvt, dlog, h, z, I, wx, wy, dz, grads[0], grads[1], grads[2], grads[3])
    assert rel(dz.float(), zp.grad) < 5e-3, rel(dz.float(), zp.grad)
    for a, p in zip(grads, ps):
        assert rel(a, p.grad) < 1e-4, rel(a, p.grad)
    # "sum" mode
    ops.loss_combine(l1, l2, None, out5, gl1, gl2, None)
    assert abs(float(out5[0]) - float((mse + ce).mean())) < 1e-5 * abs(float(out5[0])) and torch.allclose(gl1, torch.full_like(gl1, 1.0 / B))


def test_adam_with_norm_clipping_matches_torch(ops):
    g = torch.Generator(device=DEV).manual_seed(7)
    n = 100_003
    p0, grads = torch.randn(n, device=DEV, generator=g), [torch.randn(n, device=DEV, generator=g) * s for s in (0.001, 3.0, 0.5)]
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3)
    p, m, v = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    p16, ss = torch.empty(n, device=DEV, dtype=torch.bfloat16), torch.zeros(2048, device=DEV)
    for step, gr in enumerate(grads, 1):
        ref.grad = gr.clone()
        torch.nn.utils.clip_grad_norm_([ref], 1.0)
        opt.step()
        ops.sumsq(gr, ss)
        assert abs(float(ss[0]) - float((gr.double() ** 2).sum())) < 1e-4 * float(ss[0])
        first = float(ss[0])
        ops.sumsq(gr, ss)
        assert float(ss[0]) == first                                  # fixed-order reduction: bit-reproducible
        ops.adam(p, gr, m, v, 1e-3, 0.9, 0.999, 1e-8, step, sumsq=ss, max_norm=1.0, grad_scale=1.0, p16=p16)
        assert float((p - ref.data).abs().max()) < 2e-6, (step, float((p - ref.data).abs().max()))
    assert torch.equal(p16, bf(p))


def test_weights_transpose_jobs(ops):
    import numpy as np
    g = torch.Generator(device=DEV).manual_seed(8)
    shapes, offs, total = [(384, 128), (128, 512), (512, 128)], [], 0
    for s in shapes:
        offs.append(total)
        total += s[0] * s[1] + 64
    P = torch.randn(total, device=DEV, generator=g)
    PT = torch.zeros(total, device=DEV, dtype=torch.bfloat16)
    rec = np.zeros(len(shapes), dtype=[("src", "<i8"), ("dst", "<i8"), ("rows", "<i4"), ("cols", "<i4"), ("tile0", "<i4"), ("pad", "<i4")])
    t0 = 0
    for i, (s, o) in enumerate(zip(shapes, offs)):
        rec[i] = (o, o, s[0], s[1], t0, 0)
        t0 += (s[0] // 32) * (s[1] // 32)
    jobs = torch.from_numpy(rec.view(np.uint8).copy()).to(DEV)
    ops.weights_transpose(P, PT, jobs, len(shapes), t0)
    for s, o in zip(shapes, offs):
        assert torch.equal(PT[o:o + s[0] * s[1]].view(s[1], s[0]), bf(P[o:o + s[0] * s[1]].view(s).T))


@pytest.mark.parametrize("hs,H", [(32, 4), (64, 4)])
def test_tensor_core_attention_forward_backward(ops, hs, H):
    """tcgen05 attention over items of whole jets (block-diagonal inside a 128-row tile) against torch SDPA and its autograd,
    jet by jet; a jet of more than 128 particles in the batch is left to the CUDA-core kernels (its rows are not touched)."""
    from mmf_b200 import synthetic
    from mmf_b200.training import _Plan
    g = torch.Generator(device=DEV).manual_seed(hs + 1)
    ns = [3, 100, 1, 64, 128, 17, 60, 50, 140, 2, 127, 5, 121]
    plan = _Plan(synthetic.prefix_masks(torch.tensor(ns)), torch.device(DEV))
    # best-fit decreasing: 128 | 127 + 1 | 121 + 5 + 2 | 100 + 17 | 64 + 60 + 3 | 50, then the 140-particle jet
    assert plan.has_big and plan.h_items[:, 1].tolist() == [128, 128, 128, 117, 127, 50] and plan.n_packed[-1] == 140
    assert sorted(plan.h_perm.tolist()) == list(range(len(ns)))
    ns = [int(v) for v in plan.n_packed]                          # from here on: the packed order
    M, C = plan.M, hs * H
    qn, kn = bf(torch.randn(M, C, device=DEV, generator=g)), bf(torch.randn(M, C, device=DEV, generator=g))
    qkv = bf(torch.randn(M, 3 * C, device=DEV, generator=g))
    v = qkv[:, 2 * C:]
    o = torch.full((M, C), 7.0, device=DEV, dtype=torch.bfloat16)
    stats = torch.zeros(M, H, 2, device=DEV)
    ops.attn_tc_fwd(qn, kn, v, hs, plan.items, plan.n_items, plan.grid_items + 3, plan.row_jet, plan.jet_off, stats, o)
    qf, kf, vf = (t.float().clone().requires_grad_(True) for t in (qn, kn, v))
    outs, off = [], plan.h_jet_off
    for b, n in enumerate(ns):
        r = slice(int(off[b]), int(off[b + 1]))
        qq, kk, vv = (t[r].view(n, H, hs).transpose(0, 1) for t in (qf, kf, vf))
        outs.append(torch.nn.functional.scaled_dot_product_attention(qq, kk, vv).transpose(0, 1).reshape(n, C))
    want = torch.cat(outs)
    small = torch.tensor(np.repeat(np.array(ns) <= 128, ns), device=DEV)
    assert rel(o[small].float(), want[small]) < 4e-3, rel(o[small].float(), want[small])
    assert bool((o[~small] == 7.0).all())
    dO = bf(torch.randn(M, C, device=DEV, generator=g))
    want.backward(dO.float())
    dqkv = torch.full((M, 3 * C), 7.0, device=DEV, dtype=torch.bfloat16)
    ops.attn_tc_bwd(dO, qn, kn, v, hs, plan.items, plan.n_items, plan.grid_items + 3, plan.row_jet, plan.jet_off, stats, dqkv)
    assert bool((dqkv[~small] == 7.0).all())
    for w, ref in enumerate((qf.grad, kf.grad, vf.grad)):
        got = dqkv[:, w * C:(w + 1) * C][small].float()
        assert rel(got, ref[small]) < 1.2e-2, (w, rel(got, ref[small]))
