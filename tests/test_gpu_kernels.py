"""Building-block kernels through the C ABI against plain fp32 torch on the same inputs."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _abi():
    from mmf_b200 import _abi
    return _abi


def _rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20))


def _bf16(t):
    return t.to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K,mode,act", [(256, 128, 128, 0, 0), (384, 512, 256, 0, 1), (128, 256, 512, 1, 0),
                                            (256, 128, 256, 1, 0), (1280, 384, 128, 0, 1)])
def test_gemm_store(M, N, K, mode, act):
    abi = _abi()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M + N + K)
    A = _bf16(torch.randn(M, K, generator=g)).to(dev)
    W = _bf16(torch.randn(N, K, generator=g) * 0.1).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    out = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16 if mode == 0 else torch.float32)
    abi.check(abi.lib().mmf_dbg_gemm(A.data_ptr(), W.data_ptr(), bias.data_ptr(), M, N, K, mode, act, out.data_ptr(), 0, None))
    torch.cuda.synchronize()
    ref = A.float() @ W.float().T + bias
    if act:
        ref = torch.nn.functional.gelu(ref)
    tol = 6e-3 if mode == 0 else 2e-5
    assert torch.isfinite(out.float()).all()
    assert _rel(out, ref) < tol, (_rel(out, ref), (out.float() - ref).abs().max().item())


@pytest.mark.parametrize("M,C,K", [(256, 128, 128), (256, 128, 512), (384, 256, 256), (256, 256, 512)])
def test_gemm_residual_layernorm(M, C, K):
    abi = _abi()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M + C + K)
    A = _bf16(torch.randn(M, K, generator=g)).to(dev)
    W = _bf16(torch.randn(C, K, generator=g) * 0.1).to(dev)
    bias = torch.randn(C, generator=g).to(dev)
    temb = torch.randn(C, generator=g).to(dev)
    lg = (1 + 0.2 * torch.randn(C, generator=g)).to(dev)
    lb = (0.1 * torch.randn(C, generator=g)).to(dev)
    resid0 = torch.randn(M, C, generator=g).to(dev)
    resid = resid0.clone()
    act = torch.full((M, C), float("nan"), device=dev, dtype=torch.bfloat16)
    abi.check(abi.lib().mmf_dbg_gemm_resln(A.data_ptr(), W.data_ptr(), bias.data_ptr(), temb.data_ptr(), lg.data_ptr(),
                                           lb.data_ptr(), M, C, K, resid.data_ptr(), act.data_ptr(), 0, None))
    torch.cuda.synchronize()
    ref = resid0 + A.float() @ W.float().T + bias + temb
    ref_act = torch.nn.functional.layer_norm(ref, (C,), lg, lb, 1e-5)
    assert _rel(resid, ref) < 2e-5, _rel(resid, ref)
    assert _rel(act, ref_act) < 6e-3, _rel(act, ref_act)


@pytest.mark.parametrize("M,C,hs", [(256, 128, 32), (256, 256, 64), (128, 256, 32)])
def test_gemm_qkv(M, C, hs):
    abi = _abi()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M + C + hs)
    A = _bf16(torch.randn(M, C, generator=g)).to(dev)
    W = _bf16(torch.randn(3 * C, C, generator=g) * 0.1).to(dev)
    bias = torch.randn(3 * C, generator=g).to(dev)
    qg, qb, kg, kb = [(1 + 0.2 * torch.randn(hs, generator=g)).to(dev), (0.1 * torch.randn(hs, generator=g)).to(dev),
                      (1 + 0.2 * torch.randn(hs, generator=g)).to(dev), (0.1 * torch.randn(hs, generator=g)).to(dev)]
    q = torch.full((M, C), float("nan"), device=dev, dtype=torch.bfloat16)
    k = torch.full_like(q, float("nan"))
    vT = torch.full((C, M), float("nan"), device=dev, dtype=torch.bfloat16)
    abi.check(abi.lib().mmf_dbg_gemm_qkv(A.data_ptr(), W.data_ptr(), bias.data_ptr(), qg.data_ptr(), qb.data_ptr(),
                                         kg.data_ptr(), kb.data_ptr(), M, C, hs, q.data_ptr(), k.data_ptr(),
                                         vT.data_ptr(), 0, None))
    torch.cuda.synchronize()
    ref = A.float() @ W.float().T + bias
    rq, rk, rv = ref.split(C, dim=1)
    ln = lambda x, w, b: torch.nn.functional.layer_norm(x.view(M, C // hs, hs), (hs,), w, b, 1e-5).view(M, C)
    assert _rel(q, ln(rq, qg, qb)) < 6e-3, _rel(q, ln(rq, qg, qb))
    assert _rel(k, ln(rk, kg, kb)) < 6e-3, _rel(k, ln(rk, kg, kb))
    assert _rel(vT.T, rv) < 6e-3, _rel(vT.T, rv)


@pytest.mark.parametrize("C,hs", [(256, 64), (256, 32), (128, 32)])
def test_attention(C, hs):
    abi = _abi()
    dev = torch.device("cuda:0")
    jets = [1, 7, 33, 64, 129, 150, 20, 20, 20, 20, 20, 20, 20, 128, 5, 90, 90, 3]
    rows = sum(jets)
    M = (rows + 127) // 128 * 128
    g = torch.Generator().manual_seed(C + hs)
    q = _bf16(torch.randn(M, C, generator=g)).to(dev)
    k = _bf16(torch.randn(M, C, generator=g)).to(dev)
    v = _bf16(torch.randn(M, C, generator=g)).to(dev)
    vT = v.T.contiguous()
    out = torch.zeros(M, C, device=dev, dtype=torch.bfloat16)
    jn = (ctypes.c_int32 * len(jets))(*jets)
    abi.check(abi.lib().mmf_dbg_attention(q.data_ptr(), k.data_ptr(), vT.data_ptr(), jn, len(jets), M, C, hs,
                                          out.data_ptr(), 0, None))
    torch.cuda.synchronize()
    H = C // hs
    ref = torch.zeros(M, C, device=dev)
    s = 0
    for n in jets:
        qq = q[s:s + n].float().view(n, H, hs).transpose(0, 1)
        kk = k[s:s + n].float().view(n, H, hs).transpose(0, 1)
        vv = v[s:s + n].float().view(n, H, hs).transpose(0, 1)
        att = torch.softmax(qq @ kk.transpose(1, 2) / hs ** 0.5, dim=-1)
        ref[s:s + n] = (att @ vv).transpose(0, 1).reshape(n, C)
        s += n
    err = _rel(out[:rows], ref[:rows])
    worst = []
    s = 0
    for n in jets:
        worst.append(round(_rel(out[s:s + n], ref[s:s + n]), 4))
        s += n
    assert err < 1.5e-2, (err, worst)
