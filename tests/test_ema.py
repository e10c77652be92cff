"""EMA weight path: reference utils/callbacks.py:152-226 + timm ModelEmaV2.update."""
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_ema_update_is_bit_identical_to_the_torch_expression():
    from mmf_b200 import _abi
    g = torch.Generator().manual_seed(3)
    for n, decay in ((1, 0.5), (1000003, 0.9999), (4096, 0.0), (777, 1.0)):
        ema, p = torch.randn(n, generator=g), torch.randn(n, generator=g) * 3
        want = decay * ema + (1. - decay) * p                 # timm: ema_v.copy_(decay * ema_v + (1. - decay) * model_v)
        e = ema.to(DEV)
        _abi.ema_update(e, p.to(DEV), decay)
        assert torch.equal(e.cpu(), want), (n, decay, float((e.cpu() - want).abs().max()))


def test_ema_callback_follows_the_reference_hooks():
    from mmf_b200 import synthetic
    from mmf_b200.callbacks import EMACallback
    from mmf_b200.mmf import MultiModalFlowBridge
    from mmf_b200.param_spec import make_config
    from mmf_b200.tensorclass import DataCoupling, TensorMultiModal
    cfg = make_config("FusedParticleFormer", num_timesteps=3, ema_decay=0.9, use_ema_weights=True)
    bridge = MultiModalFlowBridge(cfg)
    sd0 = synthetic.make_state_dict(cfg, "wide", seed=1)
    bridge.model.load_state_dict(sd0)
    bridge = bridge.to(DEV)
    cb = EMACallback(SimpleNamespace(ema_decay=0.9, use_ema_weights=True))
    cb.on_fit_start(None, bridge)
    # two "training steps": the trained weights move, the EMA follows with decay 0.9
    sd1 = synthetic.make_state_dict(cfg, "wide", seed=2)
    want = {k: v.clone() for k, v in sd0.items()}
    for _ in range(2):
        bridge.model.load_state_dict(sd1)
        cb.on_train_batch_end(None, bridge)
        want = {k: 0.9 * want[k] + (1. - 0.9) * sd1[k] for k in want}
    got = cb.ema_model.module.state_dict()
    assert all(torch.equal(got[k].cpu(), want[k]) for k in want)
    # checkpoint hand-over and prediction with the EMA weights
    ckpt = synthetic.to_checkpoint(cfg, {k: v.cpu() for k, v in bridge.model.state_dict().items()}, ema=cb.state_dict()["ema_state_dict"])
    bridge.on_load_checkpoint(ckpt)
    src = synthetic.source_state(5, seed=4)
    batch = lambda: DataCoupling(source=src.to(DEV), target=TensorMultiModal())
    trained = bridge.simulate_dynamics(batch(), first_global_jet=0).target.continuous.clone()
    cb.on_predict_start(None, bridge)
    assert bridge.model is cb.ema_model.module
    ema_out = bridge.simulate_dynamics(batch(), first_global_jet=0).target.continuous.clone()
    cb.on_predict_end(None, bridge)
    assert bridge.model is not cb.ema_model.module
    again = bridge.simulate_dynamics(batch(), first_global_jet=0).target.continuous
    assert torch.equal(again, trained) and not torch.equal(ema_out, trained)
    # the EMA sample equals a fresh model loaded with the EMA dictionary
    ref = MultiModalFlowBridge(cfg)
    ref.model.load_state_dict(want)
    ref = ref.to(DEV)
    assert torch.equal(ref.simulate_dynamics(batch(), first_global_jet=0).target.continuous, ema_out)
