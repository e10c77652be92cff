"""Host-side mirror of the reference interface: plugin table, parameter layouts, containers, checkpoints."""
import os

import pytest
import torch

from mmf_b200 import synthetic
from mmf_b200.distributed import shard_bounds
from mmf_b200.mmf import ConditionalFlowMatching, MultiModalFlowBridge, time_grid
from mmf_b200.networks import MODEL_REGISTRY
from mmf_b200.param_spec import count_params, make_config, spec_for
from mmf_b200.tensorclass import DataCoupling, TensorMultiModal


def test_registry_keys_match_reference():
    # reference networks/registry.py:4-9
    assert set(MODEL_REGISTRY) == {"ParticleFormer", "KinFormer", "FlavorFormer", "FusedParticleFormer", "EPiC"}
    with pytest.raises(NotImplementedError):
        MODEL_REGISTRY["KinFormer"](make_config("KinFormer"))


@pytest.mark.parametrize("model,n_params", [("ParticleFormer", 5_390_092), ("FusedParticleFormer", 2_845_196),
                                            ("EPiC", 2_109_171)])
def test_parameter_counts_match_reference_probe(model, n_params):
    """SURVEY.md section 6 [probe]: parameter counts of the reference classes with train_mmf.py defaults."""
    assert count_params(spec_for(make_config(model))) == n_params


@pytest.mark.parametrize("model", ["ParticleFormer", "FusedParticleFormer", "EPiC"])
def test_shell_owns_reference_state_dict_layout(model):
    cfg = make_config(model)
    shell = MODEL_REGISTRY[model](cfg)
    sd = synthetic.make_state_dict(cfg, "wide", seed=1)
    assert list(shell.state_dict().keys()) == list(sd.keys())
    shell.load_state_dict(sd, strict=True)
    for k, v in shell.state_dict().items():
        assert torch.equal(v, sd[k])
    if model == "EPiC":          # weight-norm initialisation: g = ||v||_row
        fresh = MODEL_REGISTRY[model](cfg).state_dict()
        v, g = fresh["epic.layers.0.fc_loc1.weight_v"], fresh["epic.layers.0.fc_loc1.weight_g"]
        assert torch.allclose(g, v.norm(dim=1, keepdim=True))


def test_reference_state_dict_keys_live():
    """When the reference is mounted, its own modules must accept our state_dict unchanged."""
    import ref_harness
    if not ref_harness.available():
        pytest.skip("reference not mounted")
    ref = ref_harness.modules()
    for model in ("ParticleFormer", "FusedParticleFormer", "EPiC"):
        cfg = make_config(model)
        sd = synthetic.make_state_dict(cfg, "default", seed=0)
        m = ref.MODEL_REGISTRY[model](cfg)
        m.load_state_dict(sd, strict=True)


def test_time_grid_is_the_reference_grid():
    cfg = make_config("ParticleFormer", num_timesteps=100)
    ts, dt = time_grid(cfg)
    assert ts.dtype == torch.float32 and len(ts) == 100
    assert abs(dt - 0.010100808) < 1e-9


def test_tensor_multimodal_container():
    s = synthetic.source_state(5)
    assert len(s) == 5 and s.ndim == 3 and tuple(s.shape) == (5, 150)
    assert s.available_modes() == ["continuous", "discrete"]
    assert s.discrete.dtype == torch.int64 and s.mask.dtype == torch.int64 and s.continuous.dtype == torch.float32
    # prefix masks, source zero on pads, tokens 1..V-1 on real particles (reference sample_mmf.py:82-84)
    n = s.mask.squeeze(-1).sum(1)
    for i in range(5):
        assert s.mask[i, : n[i]].all() and not s.mask[i, n[i]:].any()
    real = s.mask.bool().squeeze(-1)
    assert (s.discrete.squeeze(-1)[real] >= 1).all() and (s.discrete.squeeze(-1)[~real] == 0).all()
    assert (s.continuous[~real] == 0).all()
    c = TensorMultiModal.cat([s[:2], s[2:]])
    assert torch.equal(c.continuous, s.continuous) and torch.equal(c.mask, s.mask)
    t = s.clone()
    t.continuous += 1
    t.apply_mask()
    assert (t.continuous[~real] == 0).all()
    assert len(DataCoupling(source=s, target=s)) == 5


def test_save_and_load_roundtrip(tmp_path):
    s = synthetic.source_state(3)
    s.time = torch.rand(3)
    out = s.save_to(str(tmp_path / "generated_sample.h5"))
    back = TensorMultiModal.load_from(out)
    for name in ("time", "continuous", "discrete", "mask"):
        assert torch.equal(getattr(back, name), getattr(s, name))


def test_checkpoint_layout_and_ema(tmp_path):
    """Lightning .ckpt layout: state_dict with 'model.' prefix, hyper_parameters, EMA dict without the prefix
    (reference model/MMF.py:112-134, scripts/sample_mmf.py:58-67)."""
    cfg = make_config("FusedParticleFormer")
    sd = synthetic.make_state_dict(cfg, "wide", seed=2)
    ema = {k: v * 0.5 for k, v in sd.items()}
    path = os.path.join(tmp_path, "best.ckpt")
    torch.save(synthetic.to_checkpoint(cfg, sd, ema), path)
    m = MultiModalFlowBridge.load_from_checkpoint(path, map_location="cpu", config=cfg)
    assert torch.equal(m.model.state_dict()["transformer.wxe.0.weight"], sd["transformer.wxe.0.weight"])
    assert m.use_ema_weights()
    assert torch.equal(m.model.state_dict()["transformer.wxe.0.weight"], ema["transformer.wxe.0.weight"])
    m2 = MultiModalFlowBridge.load_from_checkpoint(path)          # config from hyper_parameters
    assert m2.config.model == "FusedParticleFormer"
    assert isinstance(ConditionalFlowMatching(make_config("EPiC")).model, MODEL_REGISTRY["EPiC"])


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 256, 1000003):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _check_ring_plan(sizes_kb):
    """Replays a plan the way the device does and checks that no tile is written over one that may still be unread."""
    import ctypes
    import numpy as np
    from mmf_b200 import _abi
    L = _abi.lib()
    n = len(sizes_kb)
    kb = np.asarray(sizes_kb, dtype=np.int32)
    dst = np.zeros(n, dtype=np.int32)
    dep = np.zeros(n, dtype=np.int32)
    as_p = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
    _abi.check(L.mmf_dbg_ring_plan(as_p(kb), n, as_p(dst), as_p(dep)))
    assert ((dst >= 0) & (dst + kb <= 64)).all() and (dep >= 1).all() and (dep <= 12).all()
    # three timesteps: tile g may overwrite tile h (h < g) only if h <= g - dep[g % n], i.e. h is known to be consumed
    live = []                                         # (global index, begin, end) of the tiles written so far
    in_flight = []
    for g in range(3 * n):
        i = g % n
        beg, end = int(dst[i]), int(dst[i] + kb[i])
        consumed_upto = g - int(dep[i])               # the producer waits for this tile before writing
        for (h, b, e) in live:
            if b < end and beg < e:
                assert h <= consumed_upto, (g, h, dep[i])
        live = [(h, b, e) for (h, b, e) in live if not (b < end and beg < e)] + [(g, beg, end)]
        if g >= n:
            in_flight.append(g - max(consumed_upto, -1))
    return dst, dep, in_flight


def test_weight_ring_plan_uniform_tiles_is_a_plain_ring():
    dst, dep, _ = _check_ring_plan([16] * 12)
    assert sorted(set(dst.tolist())) == [0, 16, 32, 48] and (dep == 4).all()


def test_weight_ring_plan_mixed_tiles_keeps_two_big_tiles_in_flight():
    # one main block of the transformer: 4 x (QKV 4 x 24 KB, proj 32 KB), MLP 4 x 32 | 2 x 32 | 4 x 32 | 6 x 32 KB
    block = ([24] * 4 + [32]) * 4 + [32] * 16
    dst, dep, in_flight = _check_ring_plan(block * 2)
    big = [d for d, k in zip(dst.tolist(), block * 2) if k == 32]
    assert set(big) <= {0, 32}                        # 32 KB tiles never straddle the middle of the ring
    assert min(in_flight) >= 2                        # at least two tiles may always be in flight


def test_weight_ring_plan_rejects_oversized_tiles():
    import ctypes
    import numpy as np
    from mmf_b200 import _abi
    L = _abi.lib()
    kb = np.asarray([16, 80, 16], dtype=np.int32)
    out = np.zeros(3, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
    assert L.mmf_dbg_ring_plan(p(kb), 3, p(out), p(out.copy())) != 0


def test_global_jet_offset_rules(monkeypatch):
    """draws are keyed on the global jet index, so two calls must never share a range.  Single process: a cursor;
    under a predict loop with a configured batch size: (batch_idx * world + rank) * batch_size; a multi-process run without
    either information refuses instead of silently re-using draws."""
    import torch
    from mmf_b200.mmf import MultiModalFlowBridge
    from mmf_b200.param_spec import make_config
    b = MultiModalFlowBridge(make_config("FusedParticleFormer"))
    assert [b._next_jet_offset(5), b._next_jet_offset(7), b._next_jet_offset(3)] == [0, 5, 12]
    assert b._next_jet_offset(4, first_global_jet=99) == 99
    c = MultiModalFlowBridge(make_config("FusedParticleFormer", batch_size=16))
    assert c._next_jet_offset(16, batch_idx=3) == 48 and c._next_jet_offset(9, batch_idx=4) == 64      # short last batch
    monkeypatch.setattr(torch.distributed, "is_initialized", lambda: True)
    monkeypatch.setattr(torch.distributed, "get_rank", lambda: 1)
    monkeypatch.setattr(torch.distributed, "get_world_size", lambda: 4)
    assert c._next_jet_offset(16, batch_idx=2) == (2 * 4 + 1) * 16
    import pytest
    with pytest.raises(RuntimeError, match="global jet index"):
        b._next_jet_offset(5)
    with pytest.raises(RuntimeError, match="global jet index"):
        c._next_jet_offset(32, batch_idx=0)                  # batch longer than the configured stride
