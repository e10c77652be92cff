"""Generative-model front-ends with the reference's entry points.

  MultiModalFlowBridge      reference model/MMF.py:20-200   (ParticleFormer / FusedParticleFormer)
  ConditionalFlowMatching   reference model/CFM.py:13-154   (EPiC carrier)

``simulate_dynamics`` / ``predict_step`` / ``forward`` keep the reference signatures.  The N-step loop
is ONE call into libmmf_b200.so: no per-step host synchronisation (the reference has three per step,
SURVEY.md section 9).  Lightning is optional: when ``pytorch_lightning`` is importable the classes
derive from ``LightningModule`` so ``Trainer.predict`` drives them unchanged; otherwise they are
plain ``nn.Module`` objects and ``load_from_checkpoint`` reads the ``.ckpt`` with ``torch.load``.

Training (``loss``, ``MultiTaskLoss``, optimisers) is not part of the accelerated path yet.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch
from torch import nn

from . import _abi
from .networks import MODEL_REGISTRY
from .tensorclass import DataCoupling, TensorMultiModal

try:                                            # pragma: no cover - not installed in the build container
    import pytorch_lightning as _L
    _Base = _L.LightningModule
except Exception:                               # noqa: BLE001
    _Base = nn.Module


def time_grid(config):
    """t_i = linspace(eps, 1-eps, N), dt = (t_{N-1}-t_0)/(N-1) exactly as reference model/MMF.py:181-184."""
    ts = torch.linspace(config.time_eps, 1.0 - config.time_eps, config.num_timesteps)
    dt = (ts[-1] - ts[0]) / (len(ts) - 1)
    return ts, float(dt)


def _as_namespace(config):
    return SimpleNamespace(**config) if isinstance(config, dict) else config


class _GenerativeBase(_Base):
    def __init__(self, config):
        super().__init__()
        config = _as_namespace(config)
        self.config = config
        self.model = MODEL_REGISTRY[config.model](config)
        self.ema_state_from_ckpt = None
        self.seed = int(getattr(config, "seed", 0) or 0)
        self._jet_cursor = 0
        if hasattr(self, "save_hyperparameters") and _Base is not nn.Module:   # pragma: no cover
            self.save_hyperparameters(vars(config))

    # Lightning gives modules a .device; provide it for the plain-torch build
    if _Base is nn.Module:
        @property
        def device(self):
            return next(self.parameters()).device

    def forward(self, state: TensorMultiModal):
        return self.model(state)

    # ---- checkpoints (reference model/MMF.py:112-134, scripts/sample_mmf.py:58-67) ----------------------
    def on_load_checkpoint(self, checkpoint: dict) -> None:
        self.ema_state_from_ckpt = None
        cb = checkpoint.get("callbacks", {})
        if "EMACallback" in cb:
            self.ema_state_from_ckpt = cb["EMACallback"]["ema_state_dict"]

    def use_ema_weights(self) -> bool:
        """What EMACallback.on_predict_start does (reference utils/callbacks.py:182-201)."""
        if self.ema_state_from_ckpt is None:
            return False
        self.model.load_state_dict(self.ema_state_from_ckpt, strict=True)
        return True

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location="cpu", config=None, strict=True, **_):
        ckpt = torch.load(checkpoint_path, map_location=map_location, weights_only=False)
        if config is None:
            config = SimpleNamespace(**ckpt["hyper_parameters"])
        obj = cls(config)
        sd = {k[len("model."):]: v for k, v in ckpt["state_dict"].items() if k.startswith("model.")}
        obj.model.load_state_dict(sd, strict=strict)
        obj.on_load_checkpoint(ckpt)
        return obj

    def _next_jet_offset(self, B: int, first_global_jet: Optional[int] = None, batch_idx: Optional[int] = None) -> int:
        """Global index of the first jet of this call.  Draws are keyed on it, so two calls must never share a range.
        Callers that shard explicitly (``mmf_b200.distributed.generate_sharded``) pass ``first_global_jet``.  Under
        Lightning's predict loop (``predict_step(batch, batch_idx)``) the range is ``(batch_idx * world + rank) * S`` with
        the configured ``config.batch_size`` as the stride S - independent of how many calls each rank has made, of a
        short last batch and of skipped batches.  Without either, a single-process run counts the jets it has handed out;
        a multi-process run cannot know what the other ranks drew and raises instead of silently re-using draws."""
        if first_global_jet is not None:
            return int(first_global_jet)
        rank, world = 0, 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            rank, world = torch.distributed.get_rank(), torch.distributed.get_world_size()
        stride = int(getattr(self.config, "batch_size", 0) or 0)
        if batch_idx is not None and stride >= B:
            return (int(batch_idx) * world + rank) * stride
        if world > 1:
            raise RuntimeError("cannot derive the global jet index of this batch on a multi-process run: pass "
                               "first_global_jet, or set config.batch_size (>= the batch length) and call predict_step with batch_idx")
        off = self._jet_cursor
        self._jet_cursor += B
        return off


class MultiModalFlowBridge(_GenerativeBase):
    """Hybrid continuous/discrete sampler (Euler ODE + telegraph tau-leap)."""

    @torch.no_grad()
    def simulate_dynamics(self, batch: DataCoupling, u: Optional[torch.Tensor] = None,
                          forced_k: Optional[torch.Tensor] = None,
                          first_global_jet: Optional[int] = None) -> DataCoupling:
        cfg = self.config
        ts, dt = time_grid(cfg)
        src = batch.source
        dev = self.device
        B = len(src)
        opts = _abi.step_options(cfg, seed=self.seed, first_global_jet=self._next_jet_offset(B, first_global_jet))
        nm = self.model.native()
        x, k, _ = nm.generate(src.continuous.to(dev), src.discrete.to(dev), src.mask.to(dev), ts, dt, opts,
                              u=None if u is None else u.to(dev), forced_k=None if forced_k is None else forced_k.to(dev))
        batch.target = TensorMultiModal(time=torch.full((B,), float(ts[-1]), device=dev), continuous=x,
                                        discrete=k.unsqueeze(-1), mask=src.mask.to(dev))
        return batch

    @torch.no_grad()
    def predict_step(self, batch: DataCoupling, batch_idx: Optional[int] = None, dataloader_idx: int = 0) -> TensorMultiModal:
        """Returns the generated sample on the HOST (reference model/MMF.py:70-75)."""
        src = batch.source
        if src.continuous.device.type == "cpu":
            cfg = self.config
            ts, dt = time_grid(cfg)
            B = len(src)
            opts = _abi.step_options(cfg, seed=self.seed, first_global_jet=self._next_jet_offset(B, batch_idx=batch_idx))
            x, k = self.model.native().generate_host(src.continuous, src.discrete, src.mask, ts, dt, opts)
            return TensorMultiModal(time=torch.full((B,), float(ts[-1])), continuous=x, discrete=k.unsqueeze(-1),
                                    mask=src.mask)
        return self.simulate_dynamics(batch, first_global_jet=self._next_jet_offset(len(src), batch_idx=batch_idx)).target.detach().cpu()


class ConditionalFlowMatching(_GenerativeBase):
    """Continuous-only sampler (Euler ODE), the carrier of EPiC."""

    @torch.no_grad()
    def simulate_dynamics(self, batch: DataCoupling) -> DataCoupling:
        ts, dt = time_grid(self.config)
        src = batch.source
        dev = self.device
        B = len(src)
        x, _, _ = self.model.native().generate(src.continuous.to(dev), None, src.mask.to(dev), ts, dt, None)
        batch.target = TensorMultiModal(time=torch.full((B,), float(ts[-1]), device=dev), continuous=x,
                                        mask=src.mask.to(dev))
        return batch

    @torch.no_grad()
    def predict_step(self, batch: DataCoupling, batch_idx: Optional[int] = None, dataloader_idx: int = 0) -> TensorMultiModal:
        return self.simulate_dynamics(batch).target.detach().cpu()
