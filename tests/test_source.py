"""Source construction on the device: oracle/source_oracle.py against published known answers and the
reference's own sample_from_empirical_masks (law), the CUDA kernel against the oracle (masks, multiplicities, tokens bit-exact;
normals 1e-5), invariance to batching / sharding, and the distribution of the draws.
"""
import math
import os

import numpy as np
import pytest
import torch

from oracle import source_oracle as so


def _law(golden_dir):
    g = np.load(os.path.join(golden_dir, "source_law.npz"))
    D = 150
    emp_n = torch.from_numpy(g["empirical_n"].astype(np.int64))
    emp = (torch.arange(D)[None, :] < emp_n[:, None]).long().unsqueeze(-1)
    return D, emp, g["reference_sampled_n"].astype(np.int64)


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
           ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
           ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))]
    for ctr, key, out in kat:
        assert tuple(int(v) for v in so.philox4x32_10(*ctr, *key)) == out


def test_oracle_follows_the_reference_law(golden_dir):
    D, emp, ref_n = _law(golden_dir)
    probs = so.empirical_multiplicity_probs(emp.numpy(), D)
    assert probs.shape == (D + 1,) and abs(float(probs.sum()) - 1.0) < 1e-6
    B = 50000
    x, k, mask, n = so.make_source(probs, B, D, 9, seed=5)
    assert np.array_equal(mask, (np.arange(D)[None, :] < n[:, None]).astype(np.int64))            # prefix masks (aoj.py:883)
    assert set(np.unique(n)).issubset(set(np.nonzero(probs)[0]))                                  # only multiplicities with weight
    # multiplicity histogram: the oracle's draws and the reference's own draws are two samples of the same law
    ho, hr = np.bincount(n, minlength=D + 1) / B, np.bincount(ref_n, minlength=D + 1) / len(ref_n)
    assert np.abs(ho - probs).sum() < 0.05 and np.abs(hr - probs).sum() < 0.05 and np.abs(ho - hr).sum() < 0.06
    assert abs(n.mean() - ref_n.mean()) < 0.3
    # noise (sample_mmf.py:82-83): N(0,1) on real slots, zero on pads; tokens uniform on 1..8, zero on pads
    real = mask.astype(bool)
    assert (x[~real] == 0).all() and (k[~real] == 0).all()
    xr = x[real].astype(np.float64)
    assert abs(xr.mean()) < 3e-3 and abs(xr.var() - 1.0) < 5e-3 and abs((xr ** 4).mean() - 3.0) < 3e-2
    hk = np.bincount(k[real], minlength=9) / real.sum()
    assert hk[0] == 0 and np.abs(hk[1:] - 1 / 8).max() < 2e-3
    # a slot's draws depend on (seed, global jet, slot) only
    x2, k2, mask2, n2 = so.make_source(probs, 100, D, 9, seed=5, first_global_jet=777)
    assert np.array_equal(x2, x[777:877]) and np.array_equal(k2, k[777:877]) and np.array_equal(n2, n[777:877])


@pytest.mark.gpu
def test_kernel_matches_oracle_and_is_sharding_invariant(golden_dir):
    from mmf_b200 import _abi
    from mmf_b200.source import empirical_multiplicity_probs, make_source, sample_from_empirical_masks
    dev = torch.device("cuda:0")
    D, emp, ref_n = _law(golden_dir)
    probs = empirical_multiplicity_probs(emp, D)
    assert np.array_equal(probs, so.empirical_multiplicity_probs(emp.numpy(), D))
    for B, Dd, V, seed, first in ((3000, D, 9, 5, 0), (257, D, 9, (1 << 40) + 3, 123456789), (64, 7, 4, 9, 1), (5, 1, 2, 0, 0)):
        p = probs if Dd == D else np.linspace(1.0, 2.0, Dd + 1).astype(np.float32)
        x, k, mask, n = _abi.make_source(p, B, Dd, V, seed, first, dev)
        xo, ko, mo, no = so.make_source(p, B, Dd, V, seed, first)
        assert np.array_equal(n.cpu().numpy(), no) and np.array_equal(mask.cpu().numpy(), mo), (B, Dd)
        assert np.array_equal(k.cpu().numpy(), ko), (B, Dd)
        assert np.abs(x.cpu().numpy() - xo).max() <= 1e-5, (B, Dd)
    # two shards of one global sample
    xa, ka, ma, na = _abi.make_source(probs, 1000, D, 9, 11, 0, dev)
    xb, kb, mb, nb = _abi.make_source(probs, 400, D, 9, 11, 600, dev)
    assert torch.equal(xb, xa[600:]) and torch.equal(kb, ka[600:]) and torch.equal(mb, ma[600:]) and torch.equal(nb, na[600:])
    # EPiC: no discrete modality; the drop-in containers
    x, k, mask, n = _abi.make_source(probs, 10, D, 9, 11, 0, dev, discrete=False)
    assert k is None and torch.equal(x, xa[:10])
    src = make_source(probs, 100, D, 9, 1e-5, seed=11, device=dev)
    assert src.continuous.shape == (100, D, 3) and src.discrete.shape == (100, D, 1) and src.mask.shape == (100, D, 1)
    assert torch.equal(src.continuous, xa[:100]) and float(src.time[0]) == pytest.approx(1e-5)
    m = sample_from_empirical_masks(emp, 50000, D, device=dev, seed=3)
    hn = torch.bincount(m.squeeze(-1).sum(1), minlength=D + 1).float().cpu().numpy() / 50000
    hr = np.bincount(ref_n, minlength=D + 1) / len(ref_n)
    assert np.abs(hn - hr).sum() < 0.06                                       # same law as the reference's own sample


@pytest.mark.gpu
def test_kernel_draw_statistics_at_scale():
    from mmf_b200 import _abi
    dev = torch.device("cuda:0")
    D = 150
    probs = np.zeros(D + 1, dtype=np.float32); probs[30:91] = 1.0
    x, k, mask, n = _abi.make_source(probs, 200000, D, 9, 21, 0, dev)
    real = mask.bool()
    assert int(n.min()) >= 30 and int(n.max()) <= 90 and torch.equal(mask.sum(1).int(), n)
    assert float(x[~real].abs().max()) == 0.0 and int(k[~real].abs().max()) == 0
    xr = x[real].double().flatten()
    N = xr.numel()
    assert abs(float(xr.mean())) < 5 / math.sqrt(N) and abs(float(xr.var()) - 1.0) < 8 / math.sqrt(N)
    assert abs(float((xr ** 4).mean()) - 3.0) < 60 / math.sqrt(N)
    c = x[real]                                                                # the three coordinates are uncorrelated
    cc = torch.corrcoef(c.T.double())
    assert float((cc - torch.eye(3, device=dev, dtype=torch.float64)).abs().max()) < 5 / math.sqrt(c.shape[0])
    # Kolmogorov-Smirnov against the normal CDF on a 1e6 subsample
    sub = xr[:: max(1, N // 1000000)][:1000000].sort().values
    cdf = 0.5 * (1 + torch.erf(sub / math.sqrt(2)))
    ks = float((cdf - torch.arange(1, sub.numel() + 1, device=dev) / sub.numel()).abs().max())
    assert ks < 2.0 / math.sqrt(sub.numel())
    hk = torch.bincount(k[real], minlength=9).double() / int(real.sum())
    assert float(hk[0]) == 0.0 and float((hk[1:] - 1 / 8).abs().max()) < 5 * math.sqrt(1 / 8 / int(real.sum()))


@pytest.mark.gpu
def test_generation_from_a_device_source_is_batch_size_invariant(golden_dir):
    """The whole run without a host-side source (mmf_b200.distributed.generate_from_device_source): source built per batch on
    the device, sampler, gather.  Source and Philox jump draws are keyed on the global jet index, so tokens and masks do not
    depend on the batch size (bit-exact); the kinematics agree to bf16 tile-packing tolerance (a jet's tile neighbours change)."""
    from mmf_b200 import synthetic
    from mmf_b200.distributed import generate_from_device_source
    from mmf_b200.mmf import MultiModalFlowBridge
    from mmf_b200.param_spec import make_config
    from mmf_b200.source import empirical_multiplicity_probs
    from mmf_b200.tensorclass import DataCoupling
    dev = torch.device("cuda:0")
    D, emp, _ = _law(golden_dir)
    probs = empirical_multiplicity_probs(emp, D)
    cfg = make_config("FusedParticleFormer", num_timesteps=6)
    bridge = MultiModalFlowBridge(cfg)
    bridge.model.load_state_dict(synthetic.make_state_dict(cfg, "wide", 0), strict=True)
    bridge = bridge.to(dev)
    run = lambda s, g0: bridge.simulate_dynamics(DataCoupling(source=s), first_global_jet=g0).target
    a = generate_from_device_source(run, probs, 96, 96, D, cfg.vocab_size, cfg.time_eps, seed=4, device=dev)
    b = generate_from_device_source(run, probs, 96, 40, D, cfg.vocab_size, cfg.time_eps, seed=4, device=dev)
    assert len(a) == 96 and a.continuous.shape == (96, D, 3) and a.discrete.shape == (96, D, 1)
    assert torch.equal(a.mask, b.mask)
    real = a.mask.bool().squeeze(-1)
    assert (a.discrete.squeeze(-1)[real] == b.discrete.squeeze(-1)[real]).float().mean().item() > 0.97
    rel = (a.continuous - b.continuous)[real].norm() / a.continuous[real].norm()
    assert rel.item() < 2e-2, rel.item()
    assert float(a.continuous[~real].abs().max()) == 0.0
