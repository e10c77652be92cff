"""TEST INFRASTRUCTURE - import the UNMODIFIED reference behind stub modules.

Where the reference comes from, in order: ``$MMF_REFERENCE_ROOT``, ``/root/reference`` (build container only), or
``oracle/_ref`` - a byte-for-byte copy of the reference's ``multimodal_flows`` package made by the committed recipe
``oracle/make_ref.py`` (git-ignored, travels to the GPU box with the snapshot; the reference is pure Python, its
``setup.py`` installs no importable package because the sub-directories have no ``__init__.py``).

Used by ``tests/golden/make_golden.py`` to generate the committed fixtures, by ``tests/test_reference_live.py`` to re-check
the oracle against the live reference, and by ``bench.py --impl reference`` to time the reference's own sampler.

Recipe: SURVEY.md appendix A.  Missing third-party imports (h5py,
pytorch_lightning, lightning, timm) are replaced by inert stubs; the reference's
own code runs untouched.  ``torch.poisson`` is swapped for the uniform-driven
CDF inversion so that runs are reproducible.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root() -> str:
    for cand in (os.environ.get("MMF_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "multimodal_flows")):
            return cand
    return os.environ.get("MMF_REFERENCE_ROOT", "/root/reference")


REF_ROOT = _find_root()
REF_PKG = os.path.join(REF_ROOT, "multimodal_flows")


def available() -> bool:
    return os.path.isdir(REF_PKG)


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_installed = False


def install() -> None:
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")

    class LightningModule(torch.nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

        @property
        def device(self):
            return next(self.parameters()).device

    class _Empty:
        def __init__(self, *a, **k):
            pass

    if "h5py" not in sys.modules:
        try:
            import h5py  # noqa: F401
        except Exception:
            _mod("h5py")
    pl = _mod("pytorch_lightning", LightningModule=LightningModule, Callback=_Empty, Trainer=_Empty)
    pl.callbacks = _mod("pytorch_lightning.callbacks", RichProgressBar=_Empty, Callback=_Empty)
    pl.utilities = _mod("pytorch_lightning.utilities", rank_zero_only=lambda f: f)
    pl.loggers = _mod("pytorch_lightning.loggers", CometLogger=_Empty)
    _mod("lightning")
    _mod("lightning.pytorch")
    _mod("lightning.pytorch.callbacks")
    _mod("lightning.pytorch.callbacks.progress")
    _mod("lightning.pytorch.callbacks.progress.rich_progress", RichProgressBarTheme=_Empty)
    _mod("timm")
    _mod("timm.utils")
    _mod("timm.utils.model_ema", ModelEmaV2=_Empty)
    if REF_PKG not in sys.path:
        sys.path.insert(0, REF_PKG)
    _installed = True


def modules():
    """Returns a namespace with the reference classes used on the hot path."""
    install()
    from model.MMF import MultiModalFlowBridge            # type: ignore
    from model.CFM import ConditionalFlowMatching          # type: ignore
    from model.solvers import HybridSolver                 # type: ignore
    from networks.registry import MODEL_REGISTRY           # type: ignore
    from utils.tensorclass import TensorMultiModal         # type: ignore
    from utils.datasets import DataCoupling                # type: ignore
    return types.SimpleNamespace(
        MultiModalFlowBridge=MultiModalFlowBridge, ConditionalFlowMatching=ConditionalFlowMatching,
        HybridSolver=HybridSolver, MODEL_REGISTRY=MODEL_REGISTRY,
        TensorMultiModal=TensorMultiModal, DataCoupling=DataCoupling)


@contextlib.contextmanager
def supplied_uniforms(u_per_step):
    """Route ``torch.poisson`` through pre-drawn uniforms (one tensor per call, in order)."""
    it = iter(u_per_step)
    real = torch.poisson

    def fake(lam, generator=None):
        u = next(it).to(lam.device)
        assert u.shape == lam.shape, (u.shape, lam.shape)
        e = torch.exp(-lam)
        return (u >= e).to(lam.dtype) + (u >= e * (1.0 + lam)).to(lam.dtype)

    torch.poisson = fake
    try:
        yield
    finally:
        torch.poisson = real


@contextlib.contextmanager
def supplied_categorical(u_per_call, module: str = "model.solvers"):
    """Route ``Categorical(probs).sample()`` inside one module of the reference (``model.solvers`` or ``model.MJB``) through
    pre-drawn uniforms (one (B,D) tensor per call, in order): inverse CDF in channel order on the normalised probabilities
    torch itself stores."""
    install()
    import importlib
    ref_solvers = importlib.import_module(module)
    it = iter(u_per_call)
    real = ref_solvers.Categorical

    class FakeCategorical(real):
        def sample(self, sample_shape=torch.Size()):
            u = next(it).to(self.probs.device)
            assert u.shape == self.probs.shape[:-1], (u.shape, self.probs.shape)
            cum = self.probs.cumsum(-1)
            idx = (u.unsqueeze(-1) >= cum).sum(-1)
            last = (self.probs > 0).float().cumsum(-1).argmax(-1)
            return torch.minimum(idx, last)

    ref_solvers.Categorical = FakeCategorical
    try:
        yield
    finally:
        ref_solvers.Categorical = real


@contextlib.contextmanager
def supplied_rand(time, z):
    """``torch.rand`` -> the supplied per-jet uniforms behind ``time`` (model/MMF.py:146) and ``torch.randn_like`` -> the supplied
    bridge noise (model/CFM.py:182), once each, for a reproducible ``MultiModalFlowBridge.loss``."""
    real_rand, real_randn_like = torch.rand, torch.randn_like
    used = {"rand": 0, "randn_like": 0}

    def fake_rand(*size, **kw):
        used["rand"] += 1
        assert tuple(size) == tuple(time.shape) or (len(size) == 1 and tuple(size[0:1]) == tuple(time.shape)), size
        return time.clone()

    def fake_randn_like(x, **kw):
        used["randn_like"] += 1
        assert x.shape == z.shape, (x.shape, z.shape)
        return z.clone()

    torch.rand, torch.randn_like = fake_rand, fake_randn_like
    try:
        yield used
    finally:
        torch.rand, torch.randn_like = real_rand, real_randn_like
