"""Solver front-ends with the reference's call signatures, backed by the fused CUDA step kernel.

  HybridSolver(model, config).fwd_step(state, delta_t) -> (state, rates)     reference model/solvers.py:7-60
  ContinuousSolver(model, config).fwd_step(state, delta_t) -> state          reference model/solvers.py:123-143

``model`` is anything callable as ``model(state) -> (vt, logits)`` (or ``vt``).  The update itself
(temperature, softmax, filters, telegraph rates, jump draws, Euler step) is ONE kernel launch.
"""
from __future__ import annotations

import torch

from . import _abi
from .tensorclass import TensorMultiModal


class HybridSolver:
    def __init__(self, model, config, seed: int = 0):
        self.method = "tauleap"                 # hard-coded in the reference (solvers.py:9)
        self.model = model
        self.config = config
        self.vocab_size = config.vocab_size
        self.seed = seed
        self.step_index = 0
        self.first_global_jet = 0

    def fwd_step(self, state: TensorMultiModal, delta_t, u=None):
        if self.method == "tauleap":
            return self.tauleap_step(state, delta_t, u=u)
        elif self.method == "euler":
            return self.euler_step(state, delta_t, u=u)

    def check(self, device) -> None:
        """The reference asserts 0 <= k < V inside every step (model/MJB.py:177-182, two host syncs per step); here
        the kernels raise a device flag instead and this call (one sync) turns it into the same error."""
        _abi.hybrid_step_status(device)

    @torch.no_grad()
    def tauleap_step(self, state: TensorMultiModal, delta_t, u=None):
        vt, logits = self.model(state)
        B, D = state.continuous.shape[:2]
        x = state.continuous.contiguous().float()
        k = state.discrete.reshape(B, D).contiguous().long()
        if x.data_ptr() == state.continuous.data_ptr():
            x = x.clone()                        # the reference rebinds state.continuous, it does not alias the input
        if k.data_ptr() == state.discrete.data_ptr():
            k = k.clone()
        opts = _abi.step_options(self.config, seed=self.seed, first_global_jet=self.first_global_jet)
        rates = _abi.hybrid_step(vt, logits, x, k, state.time.reshape(-1), float(delta_t), opts, u=u,
                                 step_index=self.step_index, want_rates=True)
        self.step_index += 1
        state.continuous = x
        state.discrete = k.unsqueeze(-1)
        return state, rates


    @torch.no_grad()
    def euler_step(self, state: TensorMultiModal, delta_t, u=None):
        """The categorical jump of reference model/solvers.py:62-91 (temperature 1; the reference's per-class
        ``_temperature_scaling`` only broadcasts for one batch shape).  ``u``: optional (B,D) supplied uniforms."""
        vt, logits = self.model(state)
        B, D = state.continuous.shape[:2]
        x = state.continuous.contiguous().float()
        k = state.discrete.reshape(B, D).contiguous().long()
        if x.data_ptr() == state.continuous.data_ptr():
            x = x.clone()
        if k.data_ptr() == state.discrete.data_ptr():
            k = k.clone()
        opts = _abi.step_options(self.config, seed=self.seed, first_global_jet=self.first_global_jet, method=1)
        if u is not None:                        # the kernel reads channel 0 of a (B,D,V) block
            uu = torch.zeros(B, D, self.vocab_size, device=x.device, dtype=torch.float32)
            uu[..., 0] = u.to(x.device)
            u = uu
        rates = _abi.hybrid_step(vt, logits, x, k, state.time.reshape(-1), float(delta_t), opts, u=u,
                                 step_index=self.step_index, want_rates=True)
        self.step_index += 1
        state.continuous = x
        state.discrete = k.unsqueeze(-1)
        return state, rates


class ContinuousSolver:
    def __init__(self, model, config):
        self.method = "euler"
        self.model = model

    def fwd_step(self, state: TensorMultiModal, delta_t):
        if not state.has_continuous:
            return state
        return self.euler_step(state, delta_t)

    @torch.no_grad()
    def euler_step(self, state: TensorMultiModal, delta_t):
        vt = self.model(state)
        x = state.continuous.contiguous().float()
        _abi.euler_step(vt, x, float(delta_t))
        state.continuous = x
        return state
