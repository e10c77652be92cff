"""Forward half of the MMF training step: bridge sampling + masked MSE / CE + MultiTaskLoss.

Goldens (tests/golden/loss_*.npz) come from the reference's own ``MultiModalFlowBridge.loss`` (model/MMF.py:138-170) with its three
random draws supplied (tests/golden/make_golden.py: gen_loss).  CPU: the oracle restatement against them.  GPU: the kernels
through the C ABI - bridge states bit-exact, the loss kernels against torch on identical (vt, logits) to 1e-5, and the whole
``loss()`` (bf16 tensor-core encoder in the middle) within the L1 tolerance.
"""
import os

import numpy as np
import pytest
import torch

CASES = [("FusedParticleFormer", "time-weighted"), ("ParticleFormer", "time-weighted"), ("FusedParticleFormer", "sum")]


def _load(golden_dir, model, mode):
    from mmf_b200 import synthetic
    from mmf_b200.param_spec import make_config
    g = np.load(os.path.join(golden_dir, f"loss_{model}_{mode}.npz"))
    cfg = make_config(model, multitask_loss=mode, sigma=float(g["sigma"]))
    sd = synthetic.make_state_dict(cfg, flavor="wide", seed=int(g["weight_seed"]))
    assert abs(synthetic.state_dict_checksum(sd) - float(g["weight_checksum"])) < 1e-6 * abs(float(g["weight_checksum"])) + 1e-9
    sd_loss = {k[4:].replace("uncertainty_net_", "uncertainty_net.").replace("c_fc_", "c_fc.").replace("c_proj_", "c_proj."): torch.from_numpy(g[k])
               for k in g.files if k.startswith("net_")}
    T = lambda n: torch.from_numpy(g[n])
    return g, cfg, sd, sd_loss, T


@pytest.mark.parametrize("model,mode", CASES)
def test_oracle_training_loss_matches_reference_golden(model, mode, golden_dir):
    from oracle import mmf_oracle as orc
    g, cfg, sd, sd_loss, T = _load(golden_dir, model, mode)
    out = orc.training_loss(sd, sd_loss, cfg, T("x0"), T("k0").long(), T("x1"), T("k1").long(), T("mask"), T("time"), T("z"), T("u"))
    ref = g["out"]
    for i in range(5):
        if np.isnan(ref[i]):
            assert out[i] is None
        else:
            assert abs(float(out[i]) - float(ref[i])) <= 1e-6 * abs(float(ref[i])), (i, float(out[i]), float(ref[i]))
    assert torch.equal(out[5], T("xt")) and torch.equal(out[6], T("kt").long())
    # time = eps + (1 - eps) * rand, as the reference draws it
    assert torch.equal(T("time"), cfg.time_eps + (1.0 - cfg.time_eps) * T("u01"))


@pytest.mark.gpu
@pytest.mark.parametrize("model,mode", CASES)
def test_training_loss_kernels_match_reference_golden(model, mode, golden_dir):
    from mmf_b200 import _abi
    from mmf_b200.mmf import MultiModalFlowBridge
    from mmf_b200.tensorclass import DataCoupling, TensorMultiModal
    from oracle import mmf_oracle as orc
    g, cfg, sd, sd_loss, T = _load(golden_dir, model, mode)
    dev = "cuda:0"
    x0, k0, x1, k1, mask, t = T("x0"), T("k0").long(), T("x1"), T("k1").long(), T("mask"), T("time")
    # (1) bridge states: bit-exact (individually rounded fp32 ops; tie-free uniforms)
    xt, kt = _abi.bridge_sample(x0.to(dev), x1.to(dev), k0.to(dev), k1.to(dev), t.to(dev), cfg.sigma, cfg.beta, cfg.vocab_size,
                                z=T("z").to(dev), u=T("u").to(dev))
    assert torch.equal(xt.cpu(), T("xt")), float((xt.cpu() - T("xt")).abs().max())
    assert torch.equal(kt.cpu(), T("kt").long())
    # (2) loss kernels on the oracle's own fp32 (vt, logits): 1e-5
    vt, logits = orc.encoder_forward(sd, cfg, t, T("xt"), T("kt").long(), mask)
    want = orc.multitask_loss(sd_loss, cfg, vt, logits, x0, x1, k1, mask, t)
    net = None if mode == "sum" else tuple(sd_loss[k].to(dev) for k in ("uncertainty_net.c_fc.weight", "uncertainty_net.c_fc.bias",
                                                                         "uncertainty_net.c_proj.weight", "uncertainty_net.c_proj.bias"))
    out, per_jet = _abi.multitask_loss(vt.to(dev), logits.to(dev), x0.to(dev), x1.to(dev), k1.to(dev), mask.to(dev), t.to(dev), mode, cfg.n_embd, net)
    for i in range(5):
        if want[i] is not None:
            assert abs(float(out[i]) - float(want[i])) <= 1e-5 * abs(float(want[i])) + 1e-7, (i, float(out[i]), float(want[i]))
    # (3) the drop-in loss(): reference signature, bf16 encoder in the middle -> L1 tolerance on the encoder-dependent terms
    bridge = MultiModalFlowBridge(cfg)
    bridge.model.load_state_dict(sd)
    if sd_loss:
        bridge.loss_combine.load_state_dict(sd_loss, strict=True)
    bridge = bridge.to(dev)
    batch = DataCoupling(source=TensorMultiModal(continuous=x0, discrete=k0, mask=mask), target=TensorMultiModal(continuous=x1, discrete=k1, mask=mask))
    res = bridge.loss(batch, time=t, z=T("z"), u=T("u"))
    ref = g["out"]
    print(model, mode, [None if r is None else float(r) for r in res], list(ref))
    for i, tol in enumerate((3e-2, 3e-2, 3e-2, 1e-5, 1e-5)):
        if np.isnan(ref[i]):
            assert res[i] is None
        else:
            assert abs(float(res[i]) - float(ref[i])) <= tol * abs(float(ref[i])), (i, float(res[i]), float(ref[i]))
    assert set(bridge.training_step(batch)) == {"loss"} and set(bridge.validation_step(batch)) == {"val_loss"}


@pytest.mark.gpu
def test_bridge_sample_philox_draws_have_the_right_law():
    """Without supplied draws: z ~ N(0,1) (moments) and kt follows the bridge probabilities (chi-square-sized tolerance)."""
    from mmf_b200 import _abi
    from oracle import mmf_oracle as orc
    dev = "cuda:0"
    B, D, V = 4096, 150, 9
    x0 = torch.zeros(B, D, 3, device=dev); x1 = torch.zeros(B, D, 3, device=dev)
    k0 = torch.full((B, D, 1), 3, device=dev); k1 = torch.full((B, D, 1), 5, device=dev)
    t = torch.full((B,), 0.4, device=dev)
    xt, kt = _abi.bridge_sample(x0, x1, k0, k1, t, 1.0, 0.075, V, seed=9, first_global_jet=17)
    z = xt.flatten().double()
    assert abs(float(z.mean())) < 3e-3 and abs(float(z.var()) - 1.0) < 5e-3 and abs(float((z ** 4).mean()) - 3.0) < 5e-2
    kk = torch.arange(V).view(1, 1, -1).float()
    p = (orc.telegraph_conditional_probability(torch.tensor([0.4]), 1.0, kk, torch.tensor([[[5]]]), 0.075, V)
         * orc.telegraph_conditional_probability(0.0, torch.tensor([0.4]), torch.tensor([[[3]]]), kk, 0.075, V)).flatten()
    p = p / p.sum()
    freq = torch.bincount(kt.flatten().cpu(), minlength=V).double() / kt.numel()
    assert float((freq - p.double()).abs().max()) < 4e-3
    xt2, kt2 = _abi.bridge_sample(x0, x1, k0, k1, t, 1.0, 0.075, V, seed=9, first_global_jet=17)
    assert torch.equal(xt2, xt) and torch.equal(kt2, kt)            # counter-based: reproducible


@pytest.mark.parametrize("model,mode", [("FusedParticleFormer", "time-weighted"), ("FusedParticleFormer", "sum"), ("ParticleFormer", "time-weighted")])
def test_oracle_autograd_matches_reference_gradient_golden(model, mode, golden_dir):
    """tests/golden/grad_*.npz: norms and sampled entries of the gradients of the reference's own loss(batch)[0].backward()
    (tests/golden/make_golden_grads.py).  Autograd over the oracle restatement must reproduce them - that autograd is what the GPU
    tests hold the backward kernels to."""
    from oracle import mmf_oracle as orc
    g, cfg, sd, sd_loss, T = _load(golden_dir, model, mode)
    gg = np.load(os.path.join(golden_dir, f"grad_{model}_{mode}.npz"))
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    slg = {k: v.clone().requires_grad_(True) for k, v in sd_loss.items()}
    out = orc.training_loss(sdg, slg, cfg, T("x0"), T("k0").long(), T("x1"), T("k1").long(), T("mask"), T("time"), T("z"), T("u"))
    out[0].backward()
    grads = {"model." + k: v.grad for k, v in sdg.items()}
    grads.update({"loss_combine." + k: v.grad for k, v in slg.items()})
    names = [str(n) for n in gg["names"]]
    assert sorted(names) == sorted(grads)
    gnorm = sum(float(gg["norm:" + n.replace(".", "/")]) ** 2 for n in names) ** 0.5

    def positions(name, numel):                      # the generator's seeded sample positions (FNV-1a of the name)
        h = 2166136261
        for ch in name.encode():
            h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
        return np.random.default_rng(h % (2 ** 32)).integers(0, numel, size=min(48, numel))

    for n in names:
        a = grads[n].detach().double().flatten()
        want_norm, want_val = float(gg["norm:" + n.replace(".", "/")]), torch.from_numpy(gg["val:" + n.replace(".", "/")]).double()
        got_val = a[torch.from_numpy(positions(n, a.numel()))]
        # fp32 autograd both sides; reductions in another order -> 1e-4 of the tensor's own scale (or of the whole gradient for
        # the tensors whose gradient vanishes identically, e.g. the key-LayerNorm bias)
        scale = max(want_norm, 1e-6 * gnorm)
        assert abs(float(a.norm()) - want_norm) <= 2e-4 * scale, (n, float(a.norm()), want_norm)
        assert float((got_val - want_val).abs().max()) <= 2e-4 * scale, (n, float((got_val - want_val).abs().max()), scale)
