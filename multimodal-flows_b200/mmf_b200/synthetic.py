"""Deterministic synthetic inputs and weights.

No AOJ file and no trained checkpoint is available offline, so benches and
parity tests run on documented stand-ins:

  * multiplicity  n = clamp(round(55 + 18 z), 1, D), z ~ N(0,1), seed 1234
  * prefix masks  mask[i, :n_i] = 1          (reference ``utils/aoj.py:875-890``)
  * source state  x0 = randn * mask (seed 1235), k0 = randint(1, V) * mask (seed 1236)
                  (reference ``scripts/sample_mmf.py:82-84``)
  * weights       "default" = the reference initialisation (N(0, 0.02^2), LN = identity),
                  "wide"    = SURVEY.md appendix A-6 (5x wider matrices, random LN affine,
                  random biases) which gives peaked softmaxes and O(1) velocities.

Everything is drawn from CPU generators so the container and the GPU box see
the same numbers.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .param_spec import spec_for
from .tensorclass import DataCoupling, TensorMultiModal


def multiplicities(num_jets: int, max_particles: int = 150, seed: int = 1234,
                   dense: bool = False) -> torch.Tensor:
    if dense:
        return torch.full((num_jets,), max_particles, dtype=torch.int64)
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(num_jets, generator=g)
    return torch.clamp(torch.round(55.0 + 18.0 * z), 1, max_particles).to(torch.int64)


def prefix_masks(n: torch.Tensor, max_particles: int = 150) -> torch.Tensor:
    slots = torch.arange(max_particles).unsqueeze(0)
    return (slots < n.unsqueeze(1)).to(torch.int64).unsqueeze(-1)      # (B, D, 1)


def source_state(num_jets: int, max_particles: int = 150, vocab_size: int = 9,
                 dim_continuous: int = 3, dense: bool = False, seed: int = 1234,
                 with_discrete: bool = True) -> TensorMultiModal:
    n = multiplicities(num_jets, max_particles, seed, dense)
    mask = prefix_masks(n, max_particles)
    gx = torch.Generator().manual_seed(seed + 1)
    x0 = torch.randn(num_jets, max_particles, dim_continuous, generator=gx) * mask
    k0 = None
    if with_discrete:
        gk = torch.Generator().manual_seed(seed + 2)
        k0 = torch.randint(1, vocab_size, (num_jets, max_particles, 1), generator=gk) * mask
    return TensorMultiModal(time=None, continuous=x0, discrete=k0, mask=mask)


def source_batch(num_jets: int, **kw) -> DataCoupling:
    return DataCoupling(source=source_state(num_jets, **kw), target=TensorMultiModal())


def training_batch(num_jets: int, max_particles: int = 150, vocab_size: int = 9, seed: int = 1234, dense: bool = False) -> DataCoupling:
    """(source, target) pair of a training step (reference utils/datasets.py:8-41): the source as the sampler draws it, a
    synthetic "data" target on the same masks."""
    src = source_state(num_jets, max_particles, vocab_size, dense=dense, seed=seed)
    g = torch.Generator().manual_seed(seed + 7)
    x1 = (torch.randn(num_jets, max_particles, 3, generator=g) * 1.5 + 0.3) * src.mask
    k1 = torch.randint(1, vocab_size, (num_jets, max_particles, 1), generator=g) * src.mask
    return DataCoupling(source=src, target=TensorMultiModal(time=None, continuous=x1, discrete=k1, mask=src.mask))


def uniform_draws(num_steps: int, num_jets: int, max_particles: int = 150, vocab_size: int = 9,
                  seed: int = 1237) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.rand(num_steps, num_jets, max_particles, vocab_size, generator=g)


def make_state_dict(cfg, flavor: str = "wide", seed: int = 0,
                    scale: float = 5.0) -> Dict[str, torch.Tensor]:
    """fp32 state_dict with the reference's keys for ``cfg.model``."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    wide = flavor == "wide"
    if flavor not in ("wide", "default"):
        raise ValueError(flavor)
    for name, shape, kind in spec_for(cfg):
        if kind in ("w", "wn_v"):
            std = 0.02 * (scale if wide else 1.0)
            if kind == "wn_v":
                # torch's default Linear init is U(-1/sqrt(in), 1/sqrt(in)); any direction works,
                # a normal keeps the generator simple.
                std = (1.0 / shape[1] ** 0.5) * (1.5 if wide else 1.0)
            val = torch.randn(shape, generator=g) * std
        elif kind == "wn_g":
            # magnitude ~ row norm of a default Linear (= sqrt(1/3)), perturbed
            val = 0.58 * (1.0 + 0.2 * torch.randn(shape, generator=g)) * (1.5 if wide else 1.0)
        elif kind == "b":
            val = 0.05 * torch.randn(shape, generator=g) if wide else torch.zeros(shape)
        elif kind == "g":
            val = 1.0 + 0.2 * torch.randn(shape, generator=g) if wide else torch.ones(shape)
        elif kind == "s":
            val = 0.1 * torch.randn(shape, generator=g) if wide else torch.zeros(shape)
        else:
            raise ValueError(kind)
        out[name] = val.to(torch.float32).contiguous()
    return out


def state_dict_checksum(sd: Dict[str, torch.Tensor]) -> float:
    """Order-independent fp64 checksum, stored in fixtures to catch RNG drift."""
    tot = 0.0
    for name in sorted(sd):
        t = sd[name].double()
        tot += float((t * torch.arange(1, t.numel() + 1, dtype=torch.float64).reshape(t.shape).remainder(7.0)).sum())
    return tot


def to_checkpoint(cfg, sd: Dict[str, torch.Tensor], ema: Optional[Dict[str, torch.Tensor]] = None) -> dict:
    """Lightning-layout checkpoint dict (reference ``model/MMF.py:34,112-134``)."""
    ckpt = {
        "state_dict": {f"model.{k}": v for k, v in sd.items()},
        "hyper_parameters": dict(vars(cfg)),
    }
    if ema is not None:
        ckpt["callbacks"] = {"EMACallback": {"ema_state_dict": dict(ema)}}
    return ckpt
