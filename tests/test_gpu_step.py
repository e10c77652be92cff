"""The fused hybrid step kernel through the C ABI.

Bit-exact targets: the C oracle (same deterministic arithmetic) on any input, and the reference's own
outputs on the committed tie-free golden vectors.  Rates vs the reference: 1e-6 relative.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run_cuda(vt, logits, x, k, t, dt, u, T, top_k, top_p, beta=0.075, want_rates=True):
    from mmf_b200 import _abi
    dev = torch.device("cuda:0")
    opts = _abi.MmfStepOptions(temperature=T, beta=beta, top_k=int(top_k or 0), top_p=float(top_p or 0.0),
                               use_final_max_rates=0, seed=0, first_global_jet=0)
    xd = x.clone().to(dev).contiguous()
    kd = k.reshape(x.shape[0], x.shape[1]).long().clone().to(dev).contiguous()
    rates = _abi.hybrid_step(vt.to(dev), logits.to(dev), xd, kd, t.to(dev), dt, opts, u=u.to(dev) if u is not None else None,
                             want_rates=want_rates)
    torch.cuda.synchronize()
    return xd.cpu(), kd.cpu(), rates.cpu() if rates is not None else None


def test_step_matches_reference_goldens(golden_dir):
    g = np.load(os.path.join(golden_dir, "step_cases.npz"))
    for ci in range(int(g["num_cases"])):
        p = f"c{ci}_"
        T = lambda n: torch.from_numpy(g[p + n])
        top_k = int(g[p + "top_k"]) or None
        top_p = float(g[p + "top_p"]) or None
        x, k, rates = _run_cuda(T("vt"), T("logits"), T("x"), T("k"), T("t"), float(g[p + "dt"]), T("u"),
                                float(g[p + "T"]), top_k, top_p)
        k_ref = T("k_out").long().reshape(k.shape)
        assert int((k != k_ref).sum()) == 0, f"case {ci}: {int((k != k_ref).sum())} token mismatches"
        assert torch.equal(x, T("x_out")), f"case {ci}: continuous update not bit-exact"
        rel = ((rates - T("rates")).abs() / T("rates").abs()).max().item()
        assert rel < 1e-6, f"case {ci}: rates rel err {rel}"


@pytest.mark.parametrize("T,top_k,top_p", [(1.0, None, None), (0.8, None, None), (1.2, 5, None), (1.0, None, 0.9), (0.9, 3, 0.7)])
def test_step_bit_exact_vs_c_oracle(T, top_k, top_p):
    from oracle import step_oracle
    B, D, V = 64, 150, 9
    g = torch.Generator().manual_seed(7)
    vt = torch.randn(B, D, 3, generator=g) * 2
    logits = torch.randn(B, D, V, generator=g) * 2.5
    x = torch.randn(B, D, 3, generator=g)
    k = torch.randint(0, V, (B, D, 1), generator=g)
    t = torch.linspace(1e-5, 1 - 1e-5, B)
    u = torch.rand(B, D, V, generator=g)          # NOT tie-filtered: same arithmetic must give same decisions
    dt = 0.010100808
    xo, ko, ro = step_oracle.hybrid_step(vt, logits, x, k, t, dt, u, temperature=T, top_k=top_k, top_p=top_p)
    xc, kc, rc = _run_cuda(vt, logits, x, k, t, dt, u, T, top_k, top_p)
    assert torch.equal(kc, ko.reshape(kc.shape)), int((kc != ko.reshape(kc.shape)).sum())
    assert torch.equal(xc, xo)
    assert torch.equal(rc, ro), (rc - ro).abs().max().item()


def test_step_philox_statistics_and_invariance():
    """In-kernel draws: jump statistics agree with supplied uniforms, and draws depend on the global slot only."""
    from mmf_b200 import _abi
    dev = torch.device("cuda:0")
    B, D, V = 512, 150, 9
    g = torch.Generator().manual_seed(11)
    vt = torch.randn(B, D, 3, generator=g).to(dev)
    logits = (torch.randn(B, D, V, generator=g) * 2).to(dev)
    x0 = torch.randn(B, D, 3, generator=g).to(dev)
    k0 = torch.randint(0, V, (B, D), generator=g).to(dev)
    t = torch.full((B,), 0.5).to(dev)
    opts = _abi.MmfStepOptions(1.0, 0.075, 0, 0.0, 0, 1234, 0)
    ka = k0.clone(); xa = x0.clone()
    _abi.hybrid_step(vt, logits, xa, ka, t, 0.0101, opts, u=None, step_index=3, want_rates=False)
    # same jets presented as the second half of a larger launch starting at global jet 0
    opts2 = _abi.MmfStepOptions(1.0, 0.075, 0, 0.0, 0, 1234, B // 2)
    kb = k0[B // 2:].clone().contiguous(); xb = x0[B // 2:].clone().contiguous()
    _abi.hybrid_step(vt[B // 2:].contiguous(), logits[B // 2:].contiguous(), xb, kb, t[B // 2:].contiguous(), 0.0101, opts2,
                     u=None, step_index=3, want_rates=False)
    torch.cuda.synchronize()
    assert torch.equal(ka[B // 2:], kb), "draws must depend on (seed, global slot, step) only"
    u = torch.rand(B, D, V, generator=g).to(dev)
    kc = k0.clone(); xc = x0.clone()
    _abi.hybrid_step(vt, logits, xc, kc, t, 0.0101, opts, u=u, want_rates=False)
    fa = (ka != k0).float().mean().item(); fc = (kc != k0).float().mean().item()
    assert abs(fa - fc) < 0.01, (fa, fc)
    ha = torch.bincount(ka.flatten(), minlength=V).float() / ka.numel()
    hc = torch.bincount(kc.flatten(), minlength=V).float() / kc.numel()
    assert (ha - hc).abs().max().item() < 0.01


def test_production_step_matches_reproducible_step_in_distribution(monkeypatch):
    """Production mode (Philox draws, no rates) uses MUFU arithmetic and the two-uniform form of the jump law (total count
    Poisson(L): change iff exactly one event, target ~ lam_j / L).  It must agree in distribution with the
    reproducible per-channel arithmetic (MMF_STEP_EXACT=1): jump fraction, destination histogram and the joint (from, to)
    table, also with temperature and top-k / top-p filters; the Euler part is identical."""
    from mmf_b200 import _abi
    dev = torch.device("cuda:0")
    B, D, V = 4096, 150, 9
    g = torch.Generator().manual_seed(21)
    vt = torch.randn(B, D, 3, generator=g).to(dev)
    logits = (torch.randn(1, 1, V, generator=g) * 1.5 + torch.randn(B, D, V, generator=g) * 0.5).to(dev)
    x0 = torch.randn(B, D, 3, generator=g).to(dev)
    k0 = torch.randint(0, V, (B, D), generator=g).to(dev)
    for tval, opts in ((0.3, _abi.MmfStepOptions(1.0, 0.075, 0, 0.0, 0, 77, 0)), (0.9, _abi.MmfStepOptions(0.8, 0.075, 5, 0.9, 0, 78, 0))):
        t = torch.full((B,), tval).to(dev)
        out = {}
        for exact in ("1", "0"):
            monkeypatch.setenv("MMF_STEP_EXACT", exact)
            k = k0.clone(); x = x0.clone()
            _abi.hybrid_step(vt, logits, x, k, t, 0.0101, opts, u=None, step_index=5, want_rates=False)
            torch.cuda.synchronize()
            out[exact] = (k, x)
        ke, kf = out["1"][0], out["0"][0]
        n = float(k0.numel())
        fe, ff = (ke != k0).float().mean().item(), (kf != k0).float().mean().item()
        assert fe > 0.01 and abs(fe - ff) < 4 * (fe / n) ** 0.5 + 1e-4, (fe, ff)
        je = torch.bincount((k0 * V + ke).flatten(), minlength=V * V).float() / n
        jf = torch.bincount((k0 * V + kf).flatten(), minlength=V * V).float() / n
        # two independent samples of n particles: sd of a difference of frequencies p is sqrt(2 p / n); 6 sd over 81 cells
        tol = 6 * (2 * torch.maximum(je, jf).clamp_min(1.0 / n) / n).sqrt() + 3e-5
        assert ((je - jf).abs() < tol).all(), ((je - jf).abs() / tol).max().item()
        assert torch.equal(out["1"][1], out["0"][1])


@pytest.mark.parametrize("D", [7, 150])
def test_production_step_sharding_tail_and_alignment(D):
    """Production kernel (bulk-copy chunks of 512 particles, one Philox block per pair of adjacent global slots): the result
    for a particle depends on (seed, global slot, step) only - not on where the launch starts (odd first slot: pairs straddle
    Philox blocks), on the ragged last chunk, or on 16-byte alignment (unaligned views take the plain-load path)."""
    from mmf_b200 import _abi
    dev = torch.device("cuda:0")
    B, V = 1543, 9
    g = torch.Generator().manual_seed(31)
    vt = torch.randn(B, D, 3, generator=g).to(dev)
    logits = (torch.randn(B, D, V, generator=g) * 2).to(dev)
    x0 = torch.randn(B, D, 3, generator=g).to(dev)
    k0 = torch.randint(0, V, (B, D), generator=g).to(dev)
    t = torch.rand(B, generator=g).mul(0.9).to(dev)            # per-jet times
    def run(lo, hi, misalign):
        def view(a):
            a = a[lo:hi].contiguous()
            if not misalign:
                return a.clone()
            buf = torch.empty(a.numel() + 1, dtype=a.dtype, device=dev)
            v = buf[1:].view(a.shape)                          # 4 / 8 bytes past an aligned allocation
            v.copy_(a)
            return v
        x, k = view(x0), view(k0)
        opts = _abi.MmfStepOptions(1.0, 0.075, 0, 0.0, 0, 99, lo)
        _abi.hybrid_step(view(vt), view(logits), x, k, t[lo:hi].contiguous(), 0.0101, opts, u=None, step_index=7, want_rates=False)
        torch.cuda.synchronize()
        return x, k
    xa, ka = run(0, B, False)
    assert (ka != k0).float().mean().item() > 0.01
    assert torch.equal(xa, (x0 + vt * 0.0101)) or torch.allclose(xa, x0 + vt * 0.0101, rtol=0, atol=1e-6)
    for lo, hi, mis in ((1, B, False), (0, B - 3, False), (777, 1290, False), (0, B, True), (5, 1031, True)):
        xb, kb = run(lo, hi, mis)
        assert torch.equal(kb, ka[lo:hi]), (D, lo, hi, mis)
        assert torch.equal(xb, xa[lo:hi]), (D, lo, hi, mis)


def test_step_rejects_out_of_range_tokens_only_in_flag():
    """Tokens outside [0,V) are clamped and flagged (the reference asserts, MJB.py:177-182); no crash."""
    from mmf_b200 import _abi
    dev = torch.device("cuda:0")
    B, D, V = 2, 8, 9
    vt = torch.zeros(B, D, 3, device=dev); logits = torch.zeros(B, D, V, device=dev)
    x = torch.zeros(B, D, 3, device=dev); k = torch.full((B, D), 11, device=dev, dtype=torch.int64)
    t = torch.full((B,), 0.3, device=dev)
    opts = _abi.MmfStepOptions(1.0, 0.075, 0, 0.0, 0, 0, 0)
    _abi.hybrid_step(vt, logits, x, k, t, 0.01, opts, u=torch.rand(B, D, V, device=dev), want_rates=False)
    torch.cuda.synchronize()
    assert int(k.max()) < V
