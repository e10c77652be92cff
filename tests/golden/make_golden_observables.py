"""Golden fixture for the jet observables, written by EXECUTING THE REFERENCE's own analysis classes.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_observables.py

``utils/aoj.py`` imports awkward / vector / fastjet / seaborn / matplotlib at module level (absent here); they are replaced
by inert stubs, exactly as ``ref_harness`` does for h5py / lightning / timm.  ``ParticleClouds``, ``JetFeatures`` and
``flavor_mutliplicities`` then run untouched, except that the fastjet substructure call at the end of
``JetFeatures.__post_init__`` (``utils/aoj.py:464``, out of scope) is skipped.  The sample goes through the reference's
post-processing first (``utils/callbacks.py:52-57``: de-standardise, apply_mask) and is squeezed like ``scripts/sample_mmf.py:125-129``.

Writes tests/golden/observables.npz: inputs (standardised x, k, mask, mean, std) and the reference's outputs.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def reference_classes():
    ref_harness.install()
    for name in ("awkward", "fastjet", "seaborn"):
        if name not in sys.modules:
            _stub(name)
    if "vector" not in sys.modules:
        _stub("vector", register_awkward=lambda: None)
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mp = _stub("matplotlib")
        mp.pyplot = _stub("matplotlib.pyplot", rcParams={})
    from utils import aoj                                       # type: ignore
    from utils.tensorclass import TensorMultiModal              # type: ignore
    try:
        from utils.metrics import flavor_mutliplicities         # type: ignore
    except Exception:                                           # scipy.stats is present; seaborn / matplotlib are stubbed above
        raise
    aoj.JetFeatures._substructure = lambda self, **kw: None     # fastjet clustering: out of scope
    return aoj, TensorMultiModal, flavor_mutliplicities


def make_inputs(seed=5, B=96, D=150, V=9):
    g = torch.Generator().manual_seed(seed)
    n = torch.clamp(torch.round(55 + 18 * torch.randn(B, generator=g)), 1, D).long()
    n[0], n[1], n[2], n[3] = 1, 2, D, 3                          # edge multiplicities
    mask = (torch.arange(D)[None, :] < n[:, None]).long().unsqueeze(-1)
    x = torch.randn(B, D, 3, generator=g)                        # standardised sample: pads NOT zeroed (the pipeline zeroes them)
    k = torch.randint(1, V, (B, D, 1), generator=g)
    mean = [1.9, 0.0, 0.0]
    std = [0.8, 0.11, 0.1]
    return x, k, mask, mean, std


def main():
    aoj, TensorMultiModal, flavor_mutliplicities = reference_classes()
    x, k, mask, mean, std = make_inputs()
    sample = TensorMultiModal(None, x.clone(), k.clone(), mask.clone())
    sample.continuous = (sample.continuous * torch.tensor(std)) + torch.tensor(mean)     # utils/callbacks.py:52-55
    sample.apply_mask()                                                                   # utils/callbacks.py:57
    sample = sample.squeeze(-1, "discrete")                                               # scripts/sample_mmf.py:129
    jf = aoj.JetFeatures(sample)
    out = {name: getattr(jf, name).detach().numpy() for name in ("px", "py", "pz", "E", "pt", "m", "eta", "phi", "charge", "jet_charge")}
    out["numParticles"] = jf.numParticles.numpy()
    fm = flavor_mutliplicities(sample.discrete)
    np.savez_compressed(os.path.join(HERE, "observables.npz"), x=x.numpy(), k=k.numpy(), mask=mask.numpy(),
                        mean=np.array(mean, dtype=np.float32), std=np.array(std, dtype=np.float32),
                        **{"ref_" + a: b for a, b in out.items()},
                        **{"flavor_" + a.replace(" ", "_"): b.numpy() for a, b in fm.items()})
    print("wrote observables.npz:", {a: b.shape for a, b in out.items()})


if __name__ == "__main__":
    main()
