"""Plugin table of the accelerated encoders.

``MODEL_REGISTRY[config.model](config)`` is the reference's plugin lookup
(``networks/registry.py:4-9``, used at ``model/MMF.py:30`` and ``model/CFM.py:23``).
The entries here are thin ``nn.Module`` shells: they own parameters with exactly
the reference's ``state_dict`` keys and shapes (``param_spec``), so reference
checkpoints and EMA dictionaries load unchanged, and their ``forward(state)``
hands raw pointers to ``libmmf_b200.so``.  No arithmetic of the forward pass is
done in Python.

``KinFormer`` / ``FlavorFormer`` (single-modality encoders that no script of the
reference drives) are outside the accelerated path; asking for them raises.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import nn

from . import _abi
from .param_spec import spec_for
from .tensorclass import TensorMultiModal


def _register(root: nn.Module, dotted: str, value: torch.Tensor) -> None:
    parts = dotted.split(".")
    mod = root
    for name in parts[:-1]:
        child = mod._modules.get(name)
        if child is None:
            child = nn.Module()
            mod.add_module(name, child)
        mod = child
    mod.register_parameter(parts[-1], nn.Parameter(value))


def _initial_value(shape, kind: str) -> torch.Tensor:
    # reference ParticleTransformers.py:135-142 (normal 0.02 / zeros), LayerNorm identity, torch Linear default for EPiC
    if kind == "w":
        return torch.randn(shape) * 0.02
    if kind == "wn_v":
        bound = 1.0 / shape[1] ** 0.5
        return (torch.rand(shape) * 2 - 1) * bound
    if kind == "wn_g":
        return torch.full(shape, 0.577)        # overwritten below with the row norms of weight_v
    if kind == "g":
        return torch.ones(shape)
    return torch.zeros(shape)


class EncoderShell(nn.Module):
    """Parameters in the reference layout + a lazily packed native model."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.n_embd = config.n_embd
        self.max_num_particles = config.max_num_particles
        spec = spec_for(config)
        for name, shape, kind in spec:
            _register(self, name, _initial_value(shape, kind))
        with torch.no_grad():                      # weight_norm initialises g to ||v||
            sd = dict(self.named_parameters())
            for name, _, kind in spec:
                if kind == "wn_g":
                    sd[name].copy_(sd[name[:-1] + "v"].norm(dim=1, keepdim=True))
        self._native: Optional[_abi.NativeModel] = None
        self._native_sig = None

    # -- native handle management ------------------------------------------------------------
    def _signature(self):
        params = list(self.parameters())
        return (params[0].device, sum(p._version for p in params), id(params[0]))

    def native(self) -> _abi.NativeModel:
        sig = self._signature()
        if self._native is None or sig != self._native_sig:
            device = sig[0]
            if device.type != "cuda":
                raise RuntimeError(
                    f"{type(self).__name__} runs through libmmf_b200.so on a CUDA device; parameters are on "
                    f"'{device}'. Move the module with .to('cuda') (there is no CPU fallback).")
            if self._native is not None:
                self._native.close()
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            self._native = _abi.NativeModel(self.config, sd, device)
            self._native_sig = sig
        return self._native

    def refresh(self) -> None:
        """Force re-packing of the weights (after in-place edits that bypass version counters)."""
        self._native_sig = None

    def forward(self, state: TensorMultiModal):
        nm = self.native()
        t = state.time
        if t is None:
            raise ValueError("state.time is required")
        vt, logits = nm.forward(state.continuous, state.discrete, state.mask, t.reshape(-1))
        return vt if logits is None else (vt, logits)


class ParticleFormer(EncoderShell):
    """Two 128-wide streams -> 256-wide fused stream -> two heads (reference ParticleTransformers.py:17-142)."""


class FusedParticleFormer(EncoderShell):
    """Single 256-wide stream (reference ParticleTransformers.py:145-219)."""


class EPiC(EncoderShell):
    """Deep-set encoder with masked mean/sum pooling (reference networks/EPiC.py:9-77)."""


def _not_accelerated(name):
    def ctor(config):
        raise NotImplementedError(
            f"'{name}' is outside the accelerated generation path (no script of the reference drives it); "
            "use ParticleFormer, FusedParticleFormer or EPiC")
    return ctor


MODEL_REGISTRY: Dict[str, object] = {
    "ParticleFormer": ParticleFormer,
    "KinFormer": _not_accelerated("KinFormer"),
    "FlavorFormer": _not_accelerated("FlavorFormer"),
    "FusedParticleFormer": FusedParticleFormer,
    "EPiC": EPiC,
}
