"""The N-step sampler through the drop-in API on the GPU: tile kernel vs layered path, sharding invariance, and
jet-observable histograms vs the CPU oracle (SURVEY.md 8(c) level L2).

Jets of <= 128 particles run in the persistent tile kernel, larger ones in the layered kernels; both must agree with
each other (same supplied uniforms) and with the reference within the bf16 tolerance.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model(name, flavor="wide", seed=0, **over):
    from mmf_b200 import _abi, synthetic
    from mmf_b200.param_spec import make_config
    cfg = make_config(name, **over)
    sd = synthetic.make_state_dict(cfg, flavor=flavor, seed=seed)
    return cfg, sd, _abi.NativeModel(cfg, sd, torch.device(DEV))


def _rel(a, b, real):
    a, b = a[real].float(), b[real].float()
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize("name", ["FusedParticleFormer", "ParticleFormer"])
def test_tile_kernel_agrees_with_layered_path(name, monkeypatch):
    """Same weights, inputs and supplied uniforms through both CUDA paths (the layered one is forced by the env switch)."""
    from mmf_b200 import _abi, synthetic
    from oracle import mmf_oracle as orc
    cfg, sd, nm_tile = _model(name, num_timesteps=12)
    monkeypatch.setenv("MMF_NO_TILE_KERNEL", "1")
    nm_lay = _abi.NativeModel(cfg, sd, torch.device(DEV))
    monkeypatch.delenv("MMF_NO_TILE_KERNEL")
    src = synthetic.source_state(24, seed=101).to(DEV)
    u = synthetic.uniform_draws(cfg.num_timesteps, 24, seed=102).to(DEV)
    ts, dt = orc.time_grid(cfg)
    opts = _abi.step_options(cfg)
    real = src.mask.bool().squeeze(-1)
    # teacher-forced tokens so the continuous trajectories are comparable step for step
    forced = torch.randint(1, 9, (cfg.num_timesteps, 24, 150), device=DEV, dtype=torch.uint8) * src.mask.squeeze(-1).to(torch.uint8)
    xa, ka, ra = nm_tile.generate(src.continuous, src.discrete, src.mask, ts, float(dt), opts, u=u, forced_k=forced, want_rates=True)
    xb, kb, rb = nm_lay.generate(src.continuous, src.discrete, src.mask, ts, float(dt), opts, u=u, forced_k=forced, want_rates=True)
    torch.cuda.synchronize()
    assert nm_tile.launches <= 4 and nm_lay.launches > 100           # one persistent launch vs one launch per layer
    assert _rel(xa, xb, real) < 1e-2
    assert _rel(ra, rb, real) < 2e-2                                   # (tokens are forced: comparing them would be vacuous)
    # un-forced: the two paths take the same decisions except where a draw sits on a threshold
    ts4 = ts[:4]
    _, kfa, _ = nm_tile.generate(src.continuous, src.discrete, src.mask, ts4, float(dt), opts, u=u[:4])
    _, kfb, _ = nm_lay.generate(src.continuous, src.discrete, src.mask, ts4, float(dt), opts, u=u[:4])
    torch.cuda.synchronize()
    assert (kfa[real] == kfb[real]).float().mean() > 0.985
    assert (xa[~real] == 0).all() and (ka[~real] == 0).all()
    # forward API (per-jet times) through both
    t = torch.rand(24, device=DEV)
    va, la = nm_tile.forward(src.continuous, src.discrete, src.mask, t)
    vb, lb = nm_lay.forward(src.continuous, src.discrete, src.mask, t)
    assert _rel(va, vb, real) < 1e-2 and _rel(la, lb, real) < 1e-2


@pytest.mark.parametrize("name", ["FusedParticleFormer", "ParticleFormer"])
def test_ragged_edge_cases_vs_oracle(name):
    """Multiplicity edge cases against the fp32 CPU oracle (forward API, per-jet times): tiles packed with dozens of 1..3
    particle jets (every 16-key softmax group cut by jet boundaries), jets of exactly 128 particles (one full tile), 127 + 1,
    empty jets (nothing to generate, outputs stay zero), and 129..150-particle jets (layered path) in the same batch; then a
    short teacher-forced sampler run over the same batch.  Tolerances: SURVEY 8(c) level L1 (rel-L2 2e-2, max-abs 3e-2 max|ref|)."""
    from mmf_b200 import _abi, synthetic
    from oracle import mmf_oracle as orc
    cfg, sd, nm = _model(name, num_timesteps=4)
    g = torch.Generator().manual_seed(77)
    n = torch.cat([torch.randint(1, 4, (70,), generator=g), torch.tensor([128, 128, 127, 1, 0, 0, 64, 64, 129, 150, 140, 2, 126]),
                   torch.randint(30, 100, (10,), generator=g)])
    n = n[torch.randperm(len(n), generator=g)]
    B = len(n)
    mask = synthetic.prefix_masks(n, 150)
    x0 = torch.randn(B, 150, 3, generator=g) * mask
    k0 = torch.randint(1, 9, (B, 150, 1), generator=g) * mask
    t = torch.rand(B, generator=g)
    real = mask.bool().squeeze(-1)
    va, la = nm.forward(x0.to(DEV), k0.to(DEV), mask.to(DEV), t.to(DEV))
    vr, lr = orc.encoder_forward(sd, cfg, t, x0, k0, mask)
    for got, ref in ((va.cpu(), vr), (la.cpu(), lr)):
        d = (got[real] - ref[real]).float()
        assert float(d.norm() / ref[real].norm()) < 2e-2
        assert float(d.abs().max()) < 3e-2 * float(ref[real].abs().max())
        assert torch.isfinite(got).all()
    # per-jet check as well: a tiny jet must not be polluted by its tile neighbours
    for b in range(B):
        if 0 < int(n[b]) <= 3:
            d = (va.cpu()[b, :int(n[b])] - vr[b, :int(n[b])]).abs().max()
            assert float(d) < 5e-2 * float(vr[real].abs().max()), (b, int(n[b]), float(d))
    # free-running sampler over the same batch with the same supplied uniforms (4 timesteps: jumps rarely diverge)
    u = synthetic.uniform_draws(cfg.num_timesteps, B, seed=5)
    xo, ko, _ = orc.simulate_dynamics(sd, cfg, x0, k0, mask, u=u)
    ts, dt = orc.time_grid(cfg)
    xg, kg, _ = nm.generate(x0.to(DEV), k0.to(DEV), mask.to(DEV), ts, float(dt), _abi.step_options(cfg), u=u.to(DEV))
    torch.cuda.synchronize()
    assert torch.isfinite(xg).all() and (xg.cpu()[~real] == 0).all() and (kg.cpu()[~real] == 0).all()
    assert _rel(xg.cpu(), xo, real) < 2e-2
    assert (kg.cpu()[real].flatten() == ko[real].flatten()).float().mean() > 0.97


def test_philox_draws_are_keyed_on_the_global_jet_index():
    """Generating a batch in one call or as two shards (first_global_jet offsets) gives the same sample.

    The draws are identical by construction.  The encoder output of a jet may differ in the last bits with its position
    inside a 128-row tile (summation order of the masked softmax / tensor-core accumulation over keys), and a flipped
    knife-edge jump changes the rest of that jet's trajectory: two timesteps are compared tightly, eight loosely
    (measured: x rel-L2 5e-4 / 1e-2, token agreement 1.0 / 0.994)."""
    from mmf_b200 import _abi, synthetic
    from oracle import mmf_oracle as orc
    for nt, k_min, x_tol in ((2, 0.998, 2e-3), (8, 0.97, 5e-2)):      # (N = 1 has dt = 0/0 in the reference grid, MMF.py:183-185)
        cfg, sd, nm = _model("FusedParticleFormer", num_timesteps=nt)
        src = synthetic.source_state(16, seed=55).to(DEV)
        ts, dt = orc.time_grid(cfg)
        x, k, _ = nm.generate(src.continuous, src.discrete, src.mask, ts, float(dt), _abi.step_options(cfg, seed=5, first_global_jet=32))
        parts = []
        for lo, hi in ((0, 6), (6, 16)):
            s = src[lo:hi]
            parts.append(nm.generate(s.continuous, s.discrete, s.mask, ts, float(dt), _abi.step_options(cfg, seed=5, first_global_jet=32 + lo)))
        torch.cuda.synchronize()
        kc, xc = torch.cat([p[1] for p in parts]), torch.cat([p[0] for p in parts])
        real = src.mask.bool().squeeze(-1)
        assert (kc[real] == k[real]).float().mean() >= k_min
        assert _rel(xc, x, real) < x_tol
    # a different seed changes the jumps
    _, k2, _ = nm.generate(src.continuous, src.discrete, src.mask, ts, float(dt), _abi.step_options(cfg, seed=6, first_global_jet=32))
    assert not torch.equal(k2, k)


@pytest.mark.parametrize("temperature,big", [(1.0, False), (0.8, True)])
def test_free_running_sampler_histograms_match_oracle(temperature, big):
    """L2: free-running N=100 generation (in-kernel Philox) vs the fp32 CPU oracle with its own draws: jet mass,
    multiplicity-weighted token fractions and pT sums agree within the spread between two oracle seeds.  Second case:
    BASELINE config #3's temperature 0.8, with two jets of 140 / 150 particles in the batch (CTA-pair tiles for all 100 steps)."""
    from mmf_b200 import _abi, synthetic
    from oracle import mmf_oracle as orc
    cfg, sd, nm = _model("FusedParticleFormer", num_timesteps=100, temperature=temperature)
    B = 48
    src = synthetic.source_state(B, seed=900)
    if big:
        g = torch.Generator().manual_seed(901)
        for b, n in ((3, 140), (17, 150)):
            src.mask[b] = 0
            src.mask[b, :n] = 1
            src.continuous[b] = torch.randn(150, 3, generator=g) * src.mask[b]
            src.discrete[b] = torch.randint(1, 9, (150, 1), generator=g) * src.mask[b]
    ts, dt = orc.time_grid(cfg)
    x, k, _ = nm.generate(src.continuous.to(DEV), src.discrete.to(DEV), src.mask.to(DEV), ts, float(dt), _abi.step_options(cfg, seed=1))
    torch.cuda.synchronize()
    assert nm.launches <= 3
    g1, g2 = torch.Generator().manual_seed(11), torch.Generator().manual_seed(12)
    xo1, ko1, _ = orc.simulate_dynamics(sd, cfg, src.continuous, src.discrete, src.mask, generator=g1)
    xo2, ko2, _ = orc.simulate_dynamics(sd, cfg, src.continuous, src.discrete, src.mask, generator=g2)
    real = src.mask.bool().squeeze(-1)
    # the continuous ODE does not depend on the draws beyond the token feedback: trajectories stay close
    assert _rel(x.cpu(), xo1, real) < 0.15
    obs = orc.jet_observables(x.cpu(), k.cpu().unsqueeze(-1), src.mask)
    o1 = orc.jet_observables(xo1, ko1, src.mask)
    o2 = orc.jet_observables(xo2, ko2, src.mask)
    frac = lambda o: o["token_counts"].sum(0).double() / o["token_counts"].sum().double()
    spread = (frac(o1) - frac(o2)).abs().max().item()
    assert (frac(obs) - frac(o1)).abs().max().item() < max(3 * spread, 0.03)
    assert torch.equal(obs["multiplicity"], o1["multiplicity"])
    m_rel = ((obs["mass"] - o1["mass"]).abs() / (o1["mass"].abs() + 1e-3)).median().item()
    m_ref = ((o2["mass"] - o1["mass"]).abs() / (o1["mass"].abs() + 1e-3)).median().item()
    assert m_rel < max(3 * m_ref, 0.05), (m_rel, m_ref)
    if big:                                                  # the two large jets on their own: same closeness as the batch
        for b in (3, 17):
            rb = real[b:b + 1]
            assert _rel(x.cpu()[b:b + 1], xo1[b:b + 1], rb) < 0.2, b


def test_dropin_predict_step_host_roundtrip():
    """predict_step with a HOST batch returns a HOST TensorMultiModal with the reference field layout."""
    from mmf_b200 import synthetic
    from mmf_b200.mmf import MultiModalFlowBridge
    from mmf_b200.param_spec import make_config
    from mmf_b200.tensorclass import DataCoupling, TensorMultiModal
    cfg = make_config("ParticleFormer", num_timesteps=5)
    bridge = MultiModalFlowBridge(cfg)
    bridge.model.load_state_dict(synthetic.make_state_dict(cfg, "wide", seed=2))
    bridge = bridge.to(DEV)
    src = synthetic.source_state(9, seed=7)
    out = bridge.predict_step(DataCoupling(source=src, target=TensorMultiModal()), 0)
    assert out.continuous.device.type == "cpu" and out.continuous.shape == (9, 150, 3) and out.continuous.dtype == torch.float32
    assert out.discrete.shape == (9, 150, 1) and out.discrete.dtype == torch.int64
    assert out.time.shape == (9,) and abs(float(out.time[0]) - (1 - 1e-5)) < 1e-6
    real = src.mask.bool().squeeze(-1)
    assert (out.continuous[~real] == 0).all() and (out.discrete.squeeze(-1)[~real] == 0).all()
    assert out.discrete.min() >= 0 and out.discrete.max() < 9


@pytest.mark.parametrize("name", ["FusedParticleFormer", "ParticleFormer"])
def test_jets_above_128_particles_run_on_the_tile_path(name):
    """Jets of 129...150 particles are split over a 2-CTA cluster (pair tiles: K / V rows exchanged through DSMEM); they no
    longer fall back to the layered kernels.  Forward (per-jet times) and a 4-step sampler against the fp32 oracle, per jet;
    mixed with small jets (both launches in flight) and alone (only the pair launch)."""
    from mmf_b200 import _abi, synthetic
    from oracle import mmf_oracle as orc
    cfg, sd, nm = _model(name, num_timesteps=4)
    g = torch.Generator().manual_seed(91)
    for ns in ([129, 150, 140, 133, 149, 130, 17, 64, 128, 1, 150], [150, 129], [145]):
        n = torch.tensor(ns)
        B = len(n)
        mask = synthetic.prefix_masks(n, 150)
        x0 = torch.randn(B, 150, 3, generator=g) * mask
        k0 = torch.randint(1, 9, (B, 150, 1), generator=g) * mask
        t = torch.rand(B, generator=g)
        real = mask.bool().squeeze(-1)
        l0 = nm.launches
        va, la = nm.forward(x0.to(DEV), k0.to(DEV), mask.to(DEV), t.to(DEV))
        torch.cuda.synchronize()
        assert nm.launches - l0 <= 3, "pack + at most two tile launches: no layered kernels"
        vr, lr = orc.encoder_forward(sd, cfg, t, x0, k0, mask)
        for b in range(B):
            rb = real[b:b + 1]
            for got, ref, full in ((va.cpu()[b:b + 1], vr[b:b + 1], vr), (la.cpu()[b:b + 1], lr[b:b + 1], lr)):
                d = (got[rb] - ref[rb]).float()
                assert float(d.norm() / ref[rb].norm()) < 2e-2, (ns, b, int(n[b]))
                assert float(d.abs().max()) < 3e-2 * float(full[real].abs().max()), (ns, b, int(n[b]))
        assert torch.isfinite(va).all() and (va.cpu()[~real] == 0).all() and (la.cpu()[~real] == 0).all()
        u = synthetic.uniform_draws(cfg.num_timesteps, B, seed=92)
        xo, ko, ro = orc.simulate_dynamics(sd, cfg, x0, k0, mask, u=u)
        ts, dt = orc.time_grid(cfg)
        l0 = nm.launches
        xg, kg, rg = nm.generate(x0.to(DEV), k0.to(DEV), mask.to(DEV), ts, float(dt), _abi.step_options(cfg), u=u.to(DEV), want_rates=True)
        torch.cuda.synchronize()
        assert nm.launches - l0 <= 3
        assert _rel(xg.cpu(), xo, real) < 2e-2
        assert (kg.cpu()[real] == ko.squeeze(-1)[real]).float().mean() > 0.96
        assert _rel(rg.cpu(), ro, real) < 5e-2
        assert (xg.cpu()[~real] == 0).all() and (kg.cpu()[~real] == 0).all()


def test_dense_batch_of_150_particle_jets_matches_oracle():
    """The dense worst case (every jet 150 particles, `bench.py --dense`): 24 jets = 24 CTA pairs, 3 timesteps."""
    from mmf_b200 import _abi, synthetic
    from oracle import mmf_oracle as orc
    cfg, sd, nm = _model("ParticleFormer", num_timesteps=3)
    src = synthetic.source_state(24, dense=True, seed=93)
    u = synthetic.uniform_draws(3, 24, seed=94)
    xo, ko, _ = orc.simulate_dynamics(sd, cfg, src.continuous, src.discrete, src.mask, u=u)
    ts, dt = orc.time_grid(cfg)
    xg, kg, _ = nm.generate(src.continuous.to(DEV), src.discrete.to(DEV), src.mask.to(DEV), ts, float(dt), _abi.step_options(cfg), u=u.to(DEV))
    torch.cuda.synchronize()
    real = src.mask.bool().squeeze(-1)
    assert nm.launches <= 2
    assert _rel(xg.cpu(), xo, real) < 2e-2
    assert (kg.cpu()[real] == ko.squeeze(-1)[real]).float().mean() > 0.96


def test_other_vocabulary_runs_on_the_layered_kernels():
    """Outside the tile kernels' envelope (vocab_size != 9) the model still runs natively - on the layered tcgen05 kernels, one
    launch per layer - and matches the oracle: forward with per-jet times and a 3-step sampler, incl. a 140-particle jet."""
    from mmf_b200 import _abi, synthetic
    from oracle import mmf_oracle as orc
    cfg, sd, nm = _model("FusedParticleFormer", num_timesteps=3, vocab_size=8)
    g = torch.Generator().manual_seed(61)
    n = torch.tensor([140, 12, 64, 128, 1, 77])
    B = len(n)
    mask = synthetic.prefix_masks(n, 150)
    x0 = torch.randn(B, 150, 3, generator=g) * mask
    k0 = torch.randint(1, 8, (B, 150, 1), generator=g) * mask
    t = torch.rand(B, generator=g)
    real = mask.bool().squeeze(-1)
    l0 = nm.launches
    va, la = nm.forward(x0.to(DEV), k0.to(DEV), mask.to(DEV), t.to(DEV))
    torch.cuda.synchronize()
    assert nm.launches - l0 > 20 and la.shape[-1] == 8          # one launch per layer (the tile path takes 2)
    vr, lr = orc.encoder_forward(sd, cfg, t, x0, k0, mask)
    assert _rel(va.cpu(), vr, real) < 2e-2 and _rel(la.cpu(), lr, real) < 2e-2
    u = synthetic.uniform_draws(3, B, 150, 8, seed=62)
    xo, ko, _ = orc.simulate_dynamics(sd, cfg, x0, k0, mask, u=u)
    ts, dt = orc.time_grid(cfg)
    xg, kg, _ = nm.generate(x0.to(DEV), k0.to(DEV), mask.to(DEV), ts, float(dt), _abi.step_options(cfg), u=u.to(DEV))
    torch.cuda.synchronize()
    assert _rel(xg.cpu(), xo, real) < 2e-2
    assert (kg.cpu()[real] == ko.squeeze(-1)[real]).float().mean() > 0.96
    assert int(kg.max()) < 8
