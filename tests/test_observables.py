"""Jet observables: the oracle against the golden written by the reference's own ParticleClouds /
JetFeatures / flavor_mutliplicities (tests/golden/make_golden_observables.py), and the fused CUDA kernel through the C ABI
against both.

Tolerances (fp32 sums of up to 150 terms, trigonometric / hyperbolic functions from different libraries):
  px py pz E pt    |d| <= 2e-6 * sum_i |term_i|-scale (taken as E, the largest sum) + 1e-6
  m2               |d| <= 4e-6 * E^2          (difference of squares of the sums: the reference's own fp32 rounding)
  m                compared where m2 > 1e-4 * E^2 (well-conditioned), relative 1e-3
  eta, jet_charge  relative 1e-4 where |pt - |pz|| > 1e-3 * pt, NaN pattern equal (empty jets)
  phi              1e-5 absolute (mod 2 pi)
  charge, multiplicity, token counts: exact
The kernel carries the sums in fp64, so it is also checked against the fp64 reading of the formulas (the yardstick).
"""
import math
import os

import numpy as np
import pytest
import torch

from oracle import observables_oracle as obs_orc


def _golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "observables.npz"))
    x, k, mask = torch.from_numpy(g["x"]), torch.from_numpy(g["k"]), torch.from_numpy(g["mask"])
    return g, x, k.squeeze(-1), mask.squeeze(-1), [float(v) for v in g["mean"]], [float(v) for v in g["std"]]


def _compare(got, ref, E, label):
    """got / ref: dicts of (B,) tensors (float64 for the comparison)."""
    E = E.double().abs()
    for name in ("px", "py", "pz", "E", "pt"):
        d = (got[name].double() - ref[name].double()).abs()
        assert (d <= 2e-6 * E + 1e-6).all(), (label, name, d.max().item())
    pt, pz = ref["pt"].double(), ref["pz"].double()
    if "m2" in got and "m2" in ref:
        d = (got["m2"].double() - ref["m2"].double()).abs()
        assert (d <= 4e-6 * E * E + 1e-9).all(), (label, "m2", (d / (E * E + 1e-30)).max().item())
    m2 = (ref["E"].double() ** 2 - pt ** 2 - pz ** 2)
    well = m2 > 1e-4 * E * E
    assert well.any()
    assert torch.allclose(got["m"].double()[well], ref["m"].double()[well], rtol=1e-3, atol=0), (label, "m")
    cond = (pt - pz.abs()).abs() > 1e-3 * pt
    nan_ref = torch.isnan(ref["eta"].double())
    assert torch.equal(torch.isnan(got["eta"].double()), nan_ref), (label, "eta NaN pattern")
    sel = cond & ~nan_ref
    assert torch.allclose(got["eta"].double()[sel], ref["eta"].double()[sel], rtol=1e-4, atol=1e-5), (label, "eta")
    dphi = (got["phi"].double() - ref["phi"].double() + math.pi) % (2 * math.pi) - math.pi
    assert (dphi.abs()[~nan_ref] <= 1e-5).all(), (label, "phi", dphi.abs().max().item())
    if "charge" in ref:
        assert torch.equal(got["charge"].double(), ref["charge"].double()), (label, "charge")
        nan_q = torch.isnan(ref["jet_charge"].double())
        assert torch.equal(torch.isnan(got["jet_charge"].double()), nan_q), (label, "jet_charge NaN pattern")
        assert torch.allclose(got["jet_charge"].double()[~nan_q], ref["jet_charge"].double()[~nan_q], rtol=1e-4, atol=1e-5), (label, "jet_charge")


def test_oracle_reproduces_reference_observables(golden_dir):
    g, x, k, mask, mean, std = _golden(golden_dir)
    out = obs_orc.jet_observables(x, k, mask, mean, std)
    ref = {n: torch.from_numpy(g["ref_" + n]) for n in ("px", "py", "pz", "E", "pt", "m", "eta", "phi", "charge", "jet_charge")}
    for name in ("px", "py", "pz", "E", "pt", "eta", "phi", "charge", "jet_charge"):        # same torch build, same formulas: to rounding
        assert torch.allclose(out[name], ref[name], rtol=1e-6, atol=1e-6, equal_nan=True), name
    assert torch.allclose(out["m"], ref["m"], rtol=1e-4, atol=1e-4, equal_nan=True)
    assert torch.equal(out["multiplicity"], torch.from_numpy(g["ref_numParticles"]).reshape(-1))
    tokens = k * (mask > 0)
    fm = obs_orc.flavor_mutliplicities(tokens)
    for name, v in fm.items():
        assert torch.equal(v, torch.from_numpy(g["flavor_" + name.replace(" ", "_")])), name
    # the fp64 reading of the same formulas is the yardstick the kernel is held to: the fp32 reference sits within the stated tolerances of it
    out64 = obs_orc.jet_observables(x, k, mask, mean, std, dtype=torch.float64)
    _compare(ref, out64, out64["E"], "reference vs fp64 oracle")


def test_host_mirror_flavor_dictionary_matches_oracle():
    from mmf_b200.observables import flavor_mutliplicities
    g = torch.Generator().manual_seed(3)
    tok = torch.randint(0, 9, (64, 150), generator=g)
    counts = torch.stack([(tok == v).sum(dim=1) for v in range(9)], dim=1).int()
    a, b = flavor_mutliplicities(counts), obs_orc.flavor_mutliplicities(tok)
    assert a.keys() == b.keys()
    for name in a:
        assert torch.equal(a[name], b[name]), name


@pytest.mark.gpu
def test_kernel_matches_reference_golden_and_oracle(golden_dir):
    from mmf_b200 import _abi
    from mmf_b200.observables import JetFeatures
    from mmf_b200.tensorclass import TensorMultiModal
    dev = torch.device("cuda:0")
    g, x, k, mask, mean, std = _golden(golden_dir)
    kin, counts = _abi.jet_observables(x.to(dev), k.to(dev), mask.to(dev), mean, std, 9)
    torch.cuda.synchronize()
    got = {n: kin[:, i].cpu() for i, n in enumerate(_abi.OBS_COLUMNS)}
    ref = {n: torch.from_numpy(g["ref_" + n]) for n in ("px", "py", "pz", "E", "pt", "m", "eta", "phi", "charge", "jet_charge")}
    _compare(got, ref, ref["E"], "kernel vs reference golden")
    out64 = obs_orc.jet_observables(x, k, mask, mean, std, dtype=torch.float64)
    _compare(got, out64, out64["E"], "kernel vs fp64 oracle")
    assert torch.equal(got["multiplicity"].long(), torch.from_numpy(g["ref_numParticles"]).reshape(-1))
    tokens = k * (mask > 0)
    for v in range(9):
        assert torch.equal(counts[:, v].cpu().long(), ((tokens == v) & (mask > 0)).sum(dim=1)), v
    # the drop-in container: same attribute names as the reference's JetFeatures
    jf = JetFeatures(TensorMultiModal(None, x.to(dev), k.unsqueeze(-1).to(dev), mask.unsqueeze(-1).to(dev)), mean, std)
    assert torch.equal(jf.m.cpu(), got["m"]) and torch.equal(jf.numParticles.cpu(), got["multiplicity"].long())
    fm = jf.flavor_mutliplicities()
    for name, v in obs_orc.flavor_mutliplicities(tokens).items():
        assert torch.equal(fm[name].cpu(), v), name


@pytest.mark.gpu
def test_kernel_edge_cases_and_properties():
    """Empty jets, D = 1, arbitrary (non-prefix) masks, no discrete modality, identity standardisation; at a large size the
    size-independent properties: counts sum to the multiplicity, permutation of the slots inside a jet changes nothing
    beyond fp64 summation order, masked slots do not leak (garbage / NaN at pads)."""
    from mmf_b200 import _abi
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(9)
    for B, D in ((5, 1), (7, 33), (4, 150), (3, 257)):
        x = torch.randn(B, D, 3, generator=g)
        k = torch.randint(0, 9, (B, D), generator=g)
        mask = (torch.rand(B, D, generator=g) < 0.6).long()
        mask[0] = 0                                                      # an empty jet
        x[mask == 0] = float("nan")                                     # pads must never be touched
        kin, counts = _abi.jet_observables(x.to(dev), k.to(dev), mask.to(dev))
        ref = obs_orc.jet_observables(torch.nan_to_num(x), k, mask, dtype=torch.float64)
        got = {n: kin[:, i].cpu() for i, n in enumerate(_abi.OBS_COLUMNS)}
        for name in ("px", "py", "pz", "E", "pt"):
            assert torch.allclose(got[name].double(), ref[name], rtol=1e-5, atol=1e-5), (B, D, name)
        assert torch.equal(got["multiplicity"].long(), mask.sum(1))
        assert torch.equal(got["charge"].double(), ref["charge"])
        assert math.isnan(got["eta"][0].item()) and got["px"][0].item() == 0.0
        kin2, counts2 = _abi.jet_observables(x.to(dev), None, mask.to(dev))
        assert counts2 is None and torch.equal(kin2[:, :8].cpu().nan_to_num(), kin[:, :8].cpu().nan_to_num())
    B, D = 20000, 150
    gd = torch.Generator(device=dev).manual_seed(1)
    n = torch.clamp(torch.round(55 + 18 * torch.randn(B, device=dev, generator=gd)), 1, D).long()
    mask = (torch.arange(D, device=dev)[None, :] < n[:, None]).long()
    x = torch.randn(B, D, 3, device=dev, generator=gd)
    k = torch.randint(1, 9, (B, D), device=dev, generator=gd)
    kin, counts = _abi.jet_observables(x, k, mask, [1.9, 0.0, 0.0], [0.8, 0.11, 0.1])
    assert torch.equal(counts.sum(dim=1).long(), n) and torch.equal(kin[:, 10].long(), n)
    perm = torch.argsort(torch.rand(B, D, device=dev, generator=gd), dim=1)
    kin_p, counts_p = _abi.jet_observables(torch.gather(x, 1, perm.unsqueeze(-1).expand(-1, -1, 3)).contiguous(), torch.gather(k, 1, perm),
                                           torch.gather(mask, 1, perm), [1.9, 0.0, 0.0], [0.8, 0.11, 0.1])
    assert torch.equal(counts, counts_p)
    assert torch.allclose(kin[:, :5], kin_p[:, :5], rtol=1e-6, atol=1e-6)
    assert torch.equal(kin[:, 8], kin_p[:, 8])
