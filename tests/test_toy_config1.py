"""BASELINE config #1 (the reference's tutorial toy: coloured 8 Gaussians -> 2 moons, CPU) as a parity case of the hot path.

The toy's network and bridges are notebook cells, none of the package's encoders; what it shares with the hot path is the
sampler step.  oracle/toy_oracle.py restates the notebook; here a briefly trained toy is sampled for 100 grid points
(linspace(0, 1): the last point has w = 1 and an infinite rate coefficient) and at EVERY grid point the library's fused hybrid step
(vocab_size 8, beta 0.25, through HybridSolver.fwd_step) must reproduce the notebook's step bit for bit on the same (ut, ht, u)."""
import pytest
import torch

from oracle import toy_oracle as toy

S, E, BETA, SIGMA, N = 8, 128, 0.25, 0.1, 100


def _toy(train_steps=150):
    g = torch.Generator().manual_seed(5)
    x0, k0 = toy.eight_gaussians(400, g)
    x1, k1 = toy.two_moons(1600, seed=6)
    p = toy.init_params(E, S, g)
    last = toy.train(p, x0, k0, x1, k1, E, S, SIGMA, BETA, train_steps, g)
    return {k: v.detach() for k, v in p.items()}, x0, k0, last


def test_toy_oracle_trains_and_samples_on_cpu():
    """The CPU sanity run of config #1 in miniature: the loss falls and the 100-step sampler moves the 8 coloured blobs onto 2 labels."""
    g = torch.Generator().manual_seed(5)
    x0, k0 = toy.eight_gaussians(400, g)
    x1, k1 = toy.two_moons(1600, seed=6)
    p = toy.init_params(E, S, g)
    first = float(toy.loss(p, x0[:256], k0[:256], x1[:256], k1[:256], E, S, SIGMA, BETA, g))
    last = toy.train(p, x0, k0, x1, k1, E, S, SIGMA, BETA, 300, g)
    assert last < 0.75 * first
    p = {k: v.detach() for k, v in p.items()}
    ts = torch.linspace(0.0, 1.0, N)
    dt = (ts[-1] - ts[0]) / (N - 1)
    x, k = x0[:512].clone(), k0[:512].clone()
    for i in range(N):
        t = torch.full((len(x),), ts[i].item())
        ut, ht = toy.forward(p, t, x, k, E)
        x, k, _ = toy.sampler_step(ut, ht, x, k, t, dt, torch.rand(len(x), S, generator=g), BETA, S)
    assert torch.isfinite(x).all()
    assert ((k == 1) | (k == 2)).float().mean() > 0.8        # the moons carry labels 1 and 2
    assert float(x.norm(dim=1).mean()) < 4.0                 # from the ring of radius 5 towards the moons (300 Adam steps only)


@pytest.mark.gpu
def test_toy_sampler_step_is_the_library_step_bit_for_bit():
    from mmf_b200.param_spec import make_config
    from mmf_b200.solvers import HybridSolver
    from mmf_b200.tensorclass import TensorMultiModal
    dev = "cuda:0"
    p, x0, k0, _ = _toy()
    cfg = make_config("ParticleFormer", vocab_size=S, beta=BETA)
    g = torch.Generator().manual_seed(9)
    B = 1200                                                 # laid out as (8 jets, 150 slots): the step has no notion of jets beyond the time
    x, k = x0[:B].clone(), k0[:B].clone()
    ts = torch.linspace(0.0, 1.0, N)
    dt = (ts[-1] - ts[0]) / (N - 1)
    changed = 0

    class Stub:
        def __call__(self, state):
            return self.vt, self.lg

    stub = Stub()
    solver = HybridSolver(stub, cfg)
    try:                                                     # the out-of-range flag is per device: start from a clean one
        solver.check(dev)
    except RuntimeError:
        pass
    for i in range(N):
        t = torch.full((B,), ts[i].item())
        ut, ht = toy.forward(p, t, x, k, E)
        u = torch.rand(B, S, generator=g)
        xr, kr, rr = toy.sampler_step(ut, ht, x, k, t, dt, u, BETA, S)
        # the library works on (B', D, 3) kinematics and tokens 0..V-1: pad the third component, shift the tokens
        stub.vt = torch.cat([ut, torch.zeros(B, 1)], -1).view(8, 150, 3).to(dev)
        stub.lg = ht.view(8, 150, S).to(dev)
        state = TensorMultiModal(time=torch.full((8,), ts[i].item(), device=dev), continuous=torch.cat([x, torch.zeros(B, 1)], -1).view(8, 150, 3).to(dev),
                                 discrete=(k - 1).view(8, 150, 1).to(dev), mask=torch.ones(8, 150, 1, dtype=torch.int64, device=dev))
        state, rates = solver.fwd_step(state, dt, u=u.view(8, 150, S).to(dev))
        kg = state.discrete.cpu().view(B) + 1
        xg = state.continuous.cpu().view(B, 3)
        assert torch.equal(kg, kr), (i, int((kg != kr).sum()))
        assert torch.equal(xg[:, :2], xr) and (xg[:, 2] == 0).all(), i
        if i + 1 < N:                                        # (the last grid point has w = 1: rates are inf / nan in both, not compared)
            rel = ((rates.cpu().view(B, S) - rr).abs() / rr.abs()).max()
            assert rel < 2e-6, (i, float(rel))
        changed += int((kr != k).sum())
        x, k = xr, kr
    solver.check(dev)
    assert changed > 200                                     # the run did jump: the equality above is not vacuous
