"""Golden fixture for the source construction, written by EXECUTING the reference's own sample_from_empirical_masks
(utils/aoj.py:875-890, imported behind the stubs of make_golden_observables.py).

    python tests/golden/make_golden_source.py

The reference draws from torch's global generator, so the fixture pins the LAW, not the draws: the empirical masks that went
in, and the multiplicities of 50 000 masks the reference produced from them (torch.manual_seed(0)).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_observables import reference_classes  # noqa: E402


def main():
    aoj, _, _ = reference_classes()
    D = 150
    g = torch.Generator().manual_seed(1234)
    n = torch.clamp(torch.round(55 + 18 * torch.randn(4000, generator=g)), 1, D).long()          # SURVEY 8(d) stand-in for the AOJ histogram
    emp = (torch.arange(D)[None, :] < n[:, None]).long().unsqueeze(-1)
    torch.manual_seed(0)
    masks = aoj.sample_from_empirical_masks(emp, 50000, D)
    assert masks.shape == (50000, D, 1) and masks.dtype == torch.int64
    nums = masks.squeeze(-1).sum(1)
    assert torch.equal(masks.squeeze(-1), (torch.arange(D)[None, :] < nums[:, None]).long())      # prefix masks
    np.savez_compressed(os.path.join(HERE, "source_law.npz"), empirical_n=n.numpy().astype(np.int16),
                        reference_sampled_n=nums.numpy().astype(np.int16))
    print("wrote source_law.npz", float(nums.float().mean()), float(n.float().mean()))


if __name__ == "__main__":
    main()
