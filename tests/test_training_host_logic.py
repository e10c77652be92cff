"""Host logic of the training step on the CPU: TrainEngine's forward / backward PROGRAM (mmf_b200/training.py) driven through a
plain-torch stand-in for the library operators (tests/mock_train_ops.py) must reproduce torch autograd over the oracle
restatement of MultiModalFlowBridge.loss; flat-buffer bookkeeping; the DDP gradient average over gloo (world size 2)."""
import os

import numpy as np
import pytest
import torch

from grad_check import compare_gradients
from mock_train_ops import MockOps

CASES = [("FusedParticleFormer", "time-weighted"), ("ParticleFormer", "time-weighted"), ("FusedParticleFormer", "sum")]


def _setup(golden_dir, model, mode, jets=None):
    from mmf_b200 import synthetic
    from mmf_b200.mmf import MultiModalFlowBridge
    from mmf_b200.param_spec import make_config
    from mmf_b200.training import TrainEngine
    g = np.load(os.path.join(golden_dir, f"loss_{model}_{mode}.npz"))
    cfg = make_config(model, multitask_loss=mode, sigma=float(g["sigma"]), lr=1e-3)
    sd = synthetic.make_state_dict(cfg, flavor="wide", seed=int(g["weight_seed"]))
    sd_loss = {k[4:].replace("uncertainty_net_", "uncertainty_net.").replace("c_fc_", "c_fc.").replace("c_proj_", "c_proj."): torch.from_numpy(g[k])
               for k in g.files if k.startswith("net_")}
    sel = slice(None) if jets is None else jets
    T = lambda n, long=False: (torch.from_numpy(g[n]).long() if long else torch.from_numpy(g[n]))[sel]
    bridge = MultiModalFlowBridge(cfg)
    bridge.model.load_state_dict(sd)
    if sd_loss:
        bridge.loss_combine.load_state_dict(sd_loss, strict=True)
    eng = TrainEngine(bridge, lr=1e-3, _ops=MockOps())
    return cfg, sd, sd_loss, T, bridge, eng


def _run(eng, cfg, T):
    from oracle import mmf_oracle as orc
    from mmf_b200.training import _Plan
    xt, kt = orc.bridge_sample(T("x0"), T("x1"), T("k0", True), T("k1", True), T("time"), T("z"), T("u"), sigma=cfg.sigma, beta=cfg.beta,
                               vocab_size=cfg.vocab_size)
    plan = _Plan(T("mask"), torch.device("cpu"))
    rs = plan.row_slot.long()
    xs, ks = xt.reshape(-1, 3)[rs].contiguous(), kt.reshape(-1)[rs].int()
    tgt, k1p = (T("x1") - T("x0")).reshape(-1, 3)[rs].contiguous(), T("k1", True).reshape(-1)[rs].int()
    eng.G.zero_()
    c = eng._forward(plan, xs, ks, T("time"))
    out5 = eng._loss(plan, c, tgt, k1p, T("time"), True)
    eng._backward(plan, c, xs, ks)
    return out5


@pytest.mark.parametrize("model,mode", CASES)
def test_forward_backward_program_matches_autograd(model, mode, golden_dir):
    from oracle import mmf_oracle as orc
    jets = [0, 1, 2, 3] if model == "ParticleFormer" else None          # keeps the CPU suite short
    cfg, sd, sd_loss, T, bridge, eng = _setup(golden_dir, model, mode, jets)
    out5 = _run(eng, cfg, T)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    slg = {k: v.clone().requires_grad_(True) for k, v in sd_loss.items()}
    want = orc.training_loss(sdg, slg, cfg, T("x0"), T("k0", True), T("x1"), T("k1", True), T("mask"), T("time"), T("z"), T("u"))
    want[0].backward()
    assert abs(float(out5[0]) - float(want[0].detach())) < 3e-2 * abs(float(want[0].detach()))
    grads = {"model." + k: v.grad for k, v in sdg.items()}
    grads.update({"loss_combine." + k: v.grad for k, v in slg.items()})
    gcos, grel, _ = compare_gradients(eng, grads, verbose=f"{model} {mode} (mock operators)")
    assert gcos > 0.995 and grel < 5e-2


def test_flat_buffers_own_the_parameters(golden_dir):
    cfg, sd, sd_loss, T, bridge, eng = _setup(golden_dir, "FusedParticleFormer", "time-weighted", [0, 1])
    named = dict(bridge.model.named_parameters())
    for k, v in sd.items():
        assert torch.equal(named[k].detach(), v)                                   # values survived the re-homing
        assert named[k].data_ptr() == eng.p("model." + k).data_ptr()              # and live in the flat buffer
        assert named[k].grad.data_ptr() == eng.g("model." + k).data_ptr()
    assert all(eng.off[n] % 64 == 0 for n in eng.names)
    w = "model.transformer.blocks.0.ffw.c_fc.weight"
    assert torch.equal(eng.w16(w), eng.p(w).to(torch.bfloat16)) and torch.equal(eng.wT16(w), eng.p(w).to(torch.bfloat16).T)
    # an optimiser step moves the module's own parameters and refreshes both operand copies
    _run(eng, cfg, T)
    before = named["transformer.blocks.0.ffw.c_fc.weight"].detach().clone()
    eng.optimizer_step()
    after = named["transformer.blocks.0.ffw.c_fc.weight"].detach()
    assert not torch.equal(before, after)
    assert torch.equal(eng.w16(w), after.to(torch.bfloat16)) and torch.equal(eng.wT16(w), after.to(torch.bfloat16).T)
    assert set(bridge.model.state_dict()) == set(sd)                               # checkpoint layout unchanged


def _ddp_worker(rank, world, port, golden_dir, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        cfg, sd, sd_loss, T, bridge, eng = _setup(golden_dir, "FusedParticleFormer", "sum", [rank])      # one jet per rank
        _run(eng, cfg, T)
        local = eng.G.clone()
        eng.max_norm = 0.0
        eng.optimizer_step()
        q.put((rank, local.numpy(), eng.G.clone().numpy(), eng.P.clone().numpy()))
    finally:
        dist.destroy_process_group()


def test_ddp_gradient_average_over_gloo(golden_dir):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, golden_dir, q)) for r in range(2)]
    [p.start() for p in procs]
    got = sorted((q.get(timeout=600) for _ in range(2)), key=lambda t: t[0])
    [p.join(60) for p in procs]
    (_, l0, s0, p0), (_, l1, s1, p1) = got
    assert np.allclose(s0, l0 + l1) and np.array_equal(s0, s1)       # one all-reduce of the flat buffer
    assert np.array_equal(p0, p1)                                       # identical replicas after the step (grad_scale = 1 / world)


@pytest.mark.parametrize("overrides", [dict(bias=False), dict(qk_layernorm=False), dict(bias=False, qk_layernorm=False, multitask_loss="sum"),
                                       dict(multitask_loss="weighted")])
def test_program_handles_optional_parameters(overrides, golden_dir):
    """config.bias = False (no Linear / block-LayerNorm biases) and config.qk_layernorm = False (reference attention.py:32-51): the
    program skips exactly the operators whose parameters do not exist."""
    from mmf_b200 import synthetic
    from mmf_b200.mmf import MultiModalFlowBridge
    from mmf_b200.param_spec import make_config
    from mmf_b200.training import TrainEngine
    from oracle import mmf_oracle as orc
    g = np.load(os.path.join(golden_dir, "loss_FusedParticleFormer_time-weighted.npz"))
    cfg = make_config("FusedParticleFormer", sigma=float(g["sigma"]), lr=1e-3, n_layer=2, **overrides)
    sd = synthetic.make_state_dict(cfg, flavor="wide", seed=3)
    bridge = MultiModalFlowBridge(cfg)
    bridge.model.load_state_dict(sd)
    if cfg.multitask_loss == "weighted":
        bridge.loss_combine.load_state_dict({"loss_weights": torch.tensor([0.3, -0.2])})
    sd_loss = {k: v.detach().clone() for k, v in bridge.loss_combine.state_dict().items()}
    eng = TrainEngine(bridge, lr=1e-3, _ops=MockOps())
    sel = [1, 2]
    T = lambda n, long=False: (torch.from_numpy(g[n]).long() if long else torch.from_numpy(g[n]))[sel]
    out5 = _run(eng, cfg, T)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    slg = {k: v.clone().requires_grad_(True) for k, v in sd_loss.items()}
    want = orc.training_loss(sdg, slg, cfg, T("x0"), T("k0", True), T("x1"), T("k1", True), T("mask"), T("time"), T("z"), T("u"))
    want[0].backward()
    grads = {"model." + k: v.grad for k, v in sdg.items()}
    grads.update({"loss_combine." + k: v.grad for k, v in slg.items()})
    assert abs(float(out5[0]) - float(want[0].detach())) < 3e-2 * abs(float(want[0].detach()))
    gcos, grel, _ = compare_gradients(eng, grads)
    assert gcos > 0.995 and grel < 5e-2
