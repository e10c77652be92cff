"""TEST INFRASTRUCTURE - CPU restatement of the reference's jet observables (never imported by the product package).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this file.

Follows, line for line in meaning:
  utils/callbacks.py:52-56   sample.continuous = sample.continuous * sig + mu ; apply_mask()
  utils/aoj.py:333-346       ParticleClouds: px = pt cos(phi), py = pt sin(phi), pz = pt sinh(eta), E = pt cosh(eta)
  utils/aoj.py:349-368       is<Flavor> selections times the mask, charge = +1 (tokens 4, 6, 8) / -1 (tokens 3, 5, 7)
  utils/aoj.py:452-463       JetFeatures: sums over the particle axis, pt, m, eta, phi
  utils/aoj.py:514-521       _jet_charge(kappa): sum_i Q_i pt_i^kappa / pt_jet^kappa  (kappa = 0: plain sum)
  utils/metrics.py:10-33     flavor_mutliplicities

Pinned by execution: tests/golden/observables.npz is written by tests/golden/make_golden_observables.py from the
reference's own ParticleClouds / JetFeatures / flavor_mutliplicities (imported behind stubs; only the fastjet substructure
call of JetFeatures.__post_init__ is skipped) and tests/test_oracle_golden.py replays it.
`dtype=torch.float64` gives the exact-arithmetic reading of the same formulas (the yardstick for fp32 rounding).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

COLUMNS = ("px", "py", "pz", "E", "pt", "m", "eta", "phi", "charge", "jet_charge", "multiplicity", "m2")


def jet_observables(x: torch.Tensor, k: Optional[torch.Tensor], mask: torch.Tensor, mean=None, std=None,
                    dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """x (B,D,3) standardised fp32, k (B,D) int64 or None, mask (B,D) int64 -> dict of (B,) tensors."""
    B, D = x.shape[:2]
    mask = mask.reshape(B, D)
    mb = mask > 0                                                     # aoj.py:336
    cont = x.float()
    if mean is not None or std is not None:                           # callbacks.py:52-55 (fp32, multiply then add)
        mu = torch.tensor([0.0, 0.0, 0.0] if mean is None else list(mean), dtype=torch.float32)
        sig = torch.tensor([1.0, 1.0, 1.0] if std is None else list(std), dtype=torch.float32)
        cont = cont * sig + mu
    cont = cont * mb.unsqueeze(-1)                                    # apply_mask(), callbacks.py:57
    cont = cont.to(dtype)
    pt, eta_rel, phi_rel = cont[..., 0], cont[..., 1], cont[..., 2]  # aoj.py:339-341
    px = pt * torch.cos(phi_rel)                                      # aoj.py:342-345
    py = pt * torch.sin(phi_rel)
    pz = pt * torch.sinh(eta_rel)
    E = pt * torch.cosh(eta_rel)
    out = {"px": px.sum(-1), "py": py.sum(-1), "pz": pz.sum(-1), "E": E.sum(-1)}      # aoj.py:455-458
    out["pt"] = torch.sqrt(out["px"] ** 2 + out["py"] ** 2)                            # aoj.py:459
    out["m2"] = out["E"] ** 2 - out["pt"] ** 2 - out["pz"] ** 2
    out["m"] = torch.sqrt(out["m2"])                                                   # aoj.py:460
    out["eta"] = 0.5 * torch.log((out["pt"] + out["pz"]) / (out["pt"] - out["pz"]))    # aoj.py:461
    out["phi"] = torch.atan2(out["py"], out["px"])                                     # aoj.py:462
    out["multiplicity"] = mask.sum(dim=1)                                              # aoj.py:337 / 452
    if k is not None:
        k = k.reshape(B, D)
        positive = ((k == 4) | (k == 6) | (k == 8)) & mb              # aoj.py:360-361 with _flavored_kinematics (:370-372)
        negative = ((k == 3) | (k == 5) | (k == 7)) & mb
        charge = torch.zeros_like(pt)                                 # aoj.py:365-367
        charge[positive] = 1
        charge[negative] = -1
        out["charge"] = charge.sum(dim=1)                             # kappa = 0, aoj.py:521
        out["jet_charge"] = (charge * pt).sum(dim=1) / out["pt"]      # kappa = 1, aoj.py:518-519
    return out


def flavor_mutliplicities(sample: torch.Tensor) -> Dict[str, torch.Tensor]:
    """sample (B,D) tokens with pads = 0: the reference's dictionary (metrics.py:10-33)."""
    neg = ((sample == 3) | (sample == 5) | (sample == 7)).sum(dim=1)
    pos = ((sample == 4) | (sample == 6) | (sample == 8)).sum(dim=1)
    return {
        "photons": (sample == 1).sum(dim=1), "h0": (sample == 2).sum(dim=1), "h-": (sample == 3).sum(dim=1),
        "h+": (sample == 4).sum(dim=1), "e-": (sample == 5).sum(dim=1), "e+": (sample == 6).sum(dim=1),
        "mu-": (sample == 7).sum(dim=1), "mu+": (sample == 8).sum(dim=1),
        "multiplicity": (sample > 0).sum(dim=1),
        "hadrons": ((sample >= 2) & (sample <= 4)).sum(dim=1),
        "leptons": (sample > 4).sum(dim=1),
        "neutrals": ((sample == 1) | (sample == 2)).sum(dim=1),
        "negatives": neg, "positives": pos,
        "isospin": (sample == 1).sum(dim=1) - (sample == 4).sum(dim=1),
        "net charge": neg - pos,
    }
