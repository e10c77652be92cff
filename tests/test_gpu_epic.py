"""EPiC through the C ABI: one forward vs the reference goldens, the N-step Euler sampler vs the reference trajectory,
and a ragged batch (several jets per tile, > 8 tiny jets, jets split over a CTA pair) vs the oracle.

Tolerance: bf16 tensor-core operands, fp32 accumulation vs the fp32 reference on real particles:
rel-L2 <= 2e-2 and max-abs <= 3e-2 * max|ref|.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL_L2 = 2e-2
MAX_ABS = 3e-2


def _setup(flavor, seed, **over):
    from mmf_b200 import _abi, synthetic
    from mmf_b200.param_spec import make_config
    cfg = make_config("EPiC", **over)
    sd = synthetic.make_state_dict(cfg, flavor=flavor, seed=seed)
    return cfg, sd, _abi.NativeModel(cfg, sd, torch.device("cuda:0")), synthetic


def _errs(out, ref, real):
    o, r = out[real].float(), ref[real].float()
    return float((o - r).norm() / r.norm()), float((o - r).abs().max() / r.abs().max())


@pytest.mark.parametrize("flavor", ["default", "wide"])
def test_epic_forward_matches_reference_golden(flavor, golden_dir):
    g = np.load(os.path.join(golden_dir, f"encoder_EPiC_{flavor}.npz"))
    cfg, sd, nm, synthetic = _setup(flavor, int(g["weight_seed"]))
    assert abs(synthetic.state_dict_checksum(sd) - float(g["weight_checksum"])) < 1e-6 * abs(float(g["weight_checksum"])) + 1e-9
    dev = torch.device("cuda:0")
    T = lambda n: torch.from_numpy(g[n]).to(dev)
    vt, logits = nm.forward(T("continuous"), None, T("mask"), T("time"))
    torch.cuda.synchronize()
    assert logits is None
    real = T("mask").bool().squeeze(-1)
    assert torch.isfinite(vt).all() and (vt[~real] == 0).all()
    rel, mx = _errs(vt, T("vt"), real)
    # per-jet errors too: the golden batch holds n = 1, 7, 33, 64, 129, 150 (the last two run as CTA pairs)
    for b in range(vt.shape[0]):
        rb = _errs(vt[b:b + 1], T("vt")[b:b + 1], real[b:b + 1])
        assert rb[0] < REL_L2 and rb[1] < MAX_ABS, (b, rb)
    assert rel < REL_L2 and mx < MAX_ABS, (rel, mx)
    # SURVEY 8(c) L1, second clause: within 1.5x the reference's own bf16-autocast error on this fixture
    print(f"EPiC {flavor}: vt rel {rel:.2e} (autocast {float(g['autocast_vt_rel']):.2e}) max {mx:.2e} ({float(g['autocast_vt_maxabs']):.2e})")
    assert rel <= 1.5 * float(g["autocast_vt_rel"]) and mx <= 1.5 * float(g["autocast_vt_maxabs"]), (rel, mx)


def test_epic_sampler_matches_reference_trajectory(golden_dir):
    from oracle import mmf_oracle as orc
    g = np.load(os.path.join(golden_dir, "traj_EPiC.npz"))
    cfg, sd, nm, synthetic = _setup("wide", int(g["weight_seed"]), num_timesteps=int(g["num_timesteps"]))
    dev = torch.device("cuda:0")
    x0, mask = torch.from_numpy(g["x0"]).to(dev), torch.from_numpy(g["mask"]).to(dev)
    ts, dt = orc.time_grid(cfg)
    x, k, _ = nm.generate(x0, None, mask, ts, float(dt), None)
    torch.cuda.synchronize()
    assert k is None
    real = mask.bool().squeeze(-1)
    rel, mx = _errs(x, torch.from_numpy(g["x_out"]).to(dev), real)
    assert rel < REL_L2, (rel, mx)
    assert (x[~real] == 0).all()
    # in-place use: output aliasing the input
    x1 = x0.clone()
    from mmf_b200 import _abi
    import ctypes
    tg = ts.float().contiguous()
    m64 = mask.reshape(mask.shape[0], -1).contiguous()
    _abi.check(_abi.lib().mmf_generate(nm.handle, x1.data_ptr(), None, m64.data_ptr(), x1.shape[0], x1.shape[1], tg.data_ptr(),
                                       len(tg), ctypes.c_float(float(dt)), None, None, None, x1.data_ptr(), None, None,
                                       _abi.stream_handle(dev)))
    torch.cuda.synchronize()
    assert torch.equal(x1, x)


def test_epic_ragged_batch_vs_oracle():
    """96 jets with n in [1,150]: many tiny jets (the 8-jets-per-tile cap), full tiles and CTA-pair jets."""
    from oracle import mmf_oracle as orc
    cfg, sd, nm, synthetic = _setup("wide", 5, num_timesteps=6)
    g = torch.Generator().manual_seed(31)
    n = torch.cat([torch.randint(1, 6, (40,), generator=g), torch.randint(20, 129, (40,), generator=g),
                   torch.randint(129, 151, (16,), generator=g)])
    n = n[torch.randperm(len(n), generator=g)]
    mask = synthetic.prefix_masks(n, 150)
    x0 = torch.randn(len(n), 150, 3, generator=g) * mask
    t = torch.rand(len(n), generator=g)
    dev = torch.device("cuda:0")
    vt, _ = nm.forward(x0.to(dev), None, mask.to(dev), t.to(dev))
    ref = orc.epic_forward(sd, cfg, t, x0, mask)
    real = mask.bool().squeeze(-1)
    rel, mx = _errs(vt.cpu(), ref, real)
    assert rel < REL_L2 and mx < MAX_ABS, (rel, mx)
    ts, dt = orc.time_grid(cfg)
    x, _, _ = nm.generate(x0.to(dev), None, mask.to(dev), ts, float(dt), None)
    xr = orc.simulate_dynamics_cfm(sd, cfg, x0, mask)
    rel, mx = _errs(x.cpu(), xr, real)
    assert rel < REL_L2, (rel, mx)


def test_epic_dropin_module():
    """ConditionalFlowMatching(config) with MODEL_REGISTRY['EPiC']: simulate_dynamics / predict_step keep the reference contract."""
    from mmf_b200 import synthetic
    from mmf_b200.mmf import ConditionalFlowMatching
    from mmf_b200.param_spec import make_config
    from mmf_b200.tensorclass import DataCoupling, TensorMultiModal
    from oracle import mmf_oracle as orc
    cfg = make_config("EPiC", num_timesteps=10)
    sd = synthetic.make_state_dict(cfg, "wide", seed=4)
    cfm = ConditionalFlowMatching(cfg)
    cfm.model.load_state_dict(sd)
    cfm = cfm.to("cuda:0")
    src = synthetic.source_state(12, with_discrete=False, seed=8)
    out = cfm.predict_step(DataCoupling(source=src.to("cuda:0"), target=TensorMultiModal()), 0)
    assert out.continuous.device.type == "cpu" and out.discrete is None
    xr = orc.simulate_dynamics_cfm(sd, cfg, src.continuous, src.mask)
    real = src.mask.bool().squeeze(-1)
    assert _errs(out.continuous, xr, real)[0] < REL_L2
    # forward(state) -> vt
    state = TensorMultiModal(time=torch.rand(12), continuous=src.continuous, mask=src.mask).to("cuda:0")
    vt = cfm(state)
    ref = orc.epic_forward(sd, cfg, state.time.cpu(), src.continuous, src.mask)
    assert _errs(vt.cpu(), ref, real)[0] < REL_L2
