"""Jet observables of a generated sample, computed on the device in one fused pass.

Host-side mirror of the reference's analysis containers for the quantities the parity / quality report uses:

  * ``JetFeatures``            <- reference ``utils/aoj.py:448-471`` (attributes ``px py pz E pt m eta phi charge jet_charge
                                  numParticles``) built on ``ParticleClouds`` ``utils/aoj.py:333-368``; the jet charge is
                                  ``_jet_charge`` ``utils/aoj.py:514-521``
  * ``flavor_mutliplicities``  <- reference ``utils/metrics.py:10-33`` (same keys, the reference's spelling included)

The sample is passed STANDARDISED together with ``metadata['mean'/'std']``: the de-standardisation of
``FlowGeneratorCallback`` (``utils/callbacks.py:52-56``) happens inside the kernel.  Substructure observables (fastjet
clustering, ``utils/aoj.py:536-571``) are out of scope.  There is no CPU path: tensors must live on a CUDA device.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import _abi
from .tensorclass import TensorMultiModal


class JetFeatures:
    def __init__(self, data: TensorMultiModal, mean: Optional[Sequence[float]] = None, std: Optional[Sequence[float]] = None,
                 vocab_size: int = 9):
        if data.continuous is None or data.mask is None:
            raise ValueError("JetFeatures needs the continuous features and the mask")
        if not data.continuous.is_cuda:
            raise RuntimeError("mmf_b200.observables runs on the GPU only (no CPU fallback): move the sample to a CUDA device")
        kin, counts = _abi.jet_observables(data.continuous.float(), data.discrete, data.mask, mean, std, vocab_size)
        self.kin, self.counts = kin, counts
        for i, name in enumerate(_abi.OBS_COLUMNS):
            if name not in ("multiplicity", "m2", "charge", "jet_charge"):
                setattr(self, name, kin[:, i])
        self.m2 = kin[:, _abi.OBS_COLUMNS.index("m2")]
        self.numParticles = kin[:, _abi.OBS_COLUMNS.index("multiplicity")].to(torch.int64)
        if data.discrete is not None:                          # reference: only with the discrete modality (aoj.py:466-470)
            self.charge = kin[:, _abi.OBS_COLUMNS.index("charge")]
            self.jet_charge = kin[:, _abi.OBS_COLUMNS.index("jet_charge")]

    def flavor_mutliplicities(self) -> Dict[str, torch.Tensor]:
        if self.counts is None:
            raise ValueError("the sample has no discrete modality")
        return flavor_mutliplicities(self.counts)


def flavor_mutliplicities(counts: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Per-jet token counts (B,V>=9) -> the reference's feature dictionary (``utils/metrics.py:10-33``)."""
    c = counts.to(torch.int64)
    neg, pos = c[:, 3] + c[:, 5] + c[:, 7], c[:, 4] + c[:, 6] + c[:, 8]
    return {
        "photons": c[:, 1], "h0": c[:, 2], "h-": c[:, 3], "h+": c[:, 4], "e-": c[:, 5], "e+": c[:, 6], "mu-": c[:, 7], "mu+": c[:, 8],
        "multiplicity": c[:, 1:].sum(dim=1),
        "hadrons": c[:, 2] + c[:, 3] + c[:, 4],
        "leptons": c[:, 5:].sum(dim=1),
        "neutrals": c[:, 1] + c[:, 2],
        "negatives": neg, "positives": pos,
        "isospin": c[:, 1] - c[:, 4],
        "net charge": neg - pos,
    }
