"""Source state of the sampler, built on the device.

Host-side mirror of the reference's source construction:

  * ``sample_from_empirical_masks``  <- reference ``utils/aoj.py:875-890`` (same arguments; ``randomize_masks`` is not supported:
                                        the accelerated path works on prefix masks, which is what the reference samples with)
  * ``make_source``                  <- reference ``scripts/sample_mmf.py:82-87`` (``noise_continuous``, ``noise_discrete``, ``t0``,
                                        the ``TensorMultiModal`` source)

The draws come from the library's counter-based generator (``mmf_make_source``: Philox4x32-10 keyed on seed, GLOBAL jet index
and slot), so a sample is reproducible and independent of batch size and of the sharding over GPUs; the reference's torch
generator streams cannot be reproduced, parity is in distribution.  GPU only - there is no CPU fallback.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import _abi
from .tensorclass import TensorMultiModal


def empirical_multiplicity_probs(pad_masks: torch.Tensor, max_num_particles: int = 150) -> np.ndarray:
    """Density histogram of the multiplicities of ``pad_masks`` (B,D,1) - reference ``utils/aoj.py:876-877``."""
    nums = pad_masks.squeeze(-1).sum(1)
    probs, _ = np.histogram(nums.cpu().numpy(), bins=np.arange(0, max_num_particles + 2, 1), density=True)
    return probs.astype(np.float32)


def sample_from_empirical_masks(pad_masks: torch.Tensor, num_jets: int, max_num_particles: int = 150, randomize_masks: bool = False,
                                device="cuda", seed: int = 0, first_global_jet: int = 0) -> torch.Tensor:
    """(num_jets, D, 1) int64 prefix masks with multiplicities drawn from the empirical histogram of ``pad_masks``."""
    if randomize_masks:
        raise NotImplementedError("randomize_masks is not supported by the accelerated path (prefix masks only)")
    probs = empirical_multiplicity_probs(pad_masks, max_num_particles)
    _, _, mask, _ = _abi.make_source(probs, num_jets, max_num_particles, 2, seed, first_global_jet, device, discrete=False)
    return mask.unsqueeze(-1)


def make_source(mult_probs: Sequence[float], num_jets: int, max_num_particles: int = 150, vocab_size: int = 9, time_eps: float = 1e-5,
                seed: int = 0, first_global_jet: int = 0, device="cuda", discrete: bool = True) -> TensorMultiModal:
    """The reference's ``source`` (``sample_mmf.py:82-87``) on the device: time (B,) = eps, continuous (B,D,3) = N(0,1) * mask,
    discrete (B,D,1) = U{1..V-1} * mask (None for EPiC), mask (B,D,1)."""
    x0, k0, mask, _ = _abi.make_source(mult_probs, num_jets, max_num_particles, vocab_size, seed, first_global_jet, device, discrete)
    t0 = torch.full((num_jets,), float(time_eps), device=x0.device)
    return TensorMultiModal(time=t0, continuous=x0, discrete=None if k0 is None else k0.unsqueeze(-1), mask=mask.unsqueeze(-1))
