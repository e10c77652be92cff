"""EMA of the weights: the reference's ``EMACallback`` (``utils/callbacks.py:152-226``) and the ``timm.utils.model_ema.ModelEmaV2``
it drives, with the same hooks and the same checkpoint hand-over:

  * ``on_fit_start``            EMA copy of ``pl_module.model`` (resumed from a cached ``ema_state_dict``)
  * ``on_train_batch_end``      ``ema <- decay * ema + (1 - decay) * p`` for every ``state_dict`` entry - here one fused kernel
                                launch per tensor (``mmf_ema_update``), bit-identical to the torch expression
  * ``on_validation_epoch_*`` / ``on_predict_*``   swap the EMA module in and the trained module back
  * the EMA dictionary travels in the checkpoint as ``callbacks.EMACallback.ema_state_dict`` (reference ``model/MMF.py:112-134``);
    ``MultiModalFlowBridge.on_load_checkpoint`` caches it as ``ema_state_from_ckpt`` exactly like the reference.
"""
from __future__ import annotations

import copy
from typing import Optional

import torch

from . import _abi

try:                                            # pragma: no cover - not installed in the build container
    from pytorch_lightning import Callback as _CallbackBase
except Exception:                               # noqa: BLE001
    _CallbackBase = object


class ModelEma:
    """``timm.utils.model_ema.ModelEmaV2``: a deep copy in eval mode, ``update(model)`` over the zipped ``state_dict`` values."""

    def __init__(self, model: torch.nn.Module, decay: float = 0.9999):
        self.module = copy.deepcopy(model)
        self.module.eval()
        self.decay = decay

    def to(self, device):
        self.module.to(device)
        return self

    @torch.no_grad()
    def update(self, model: torch.nn.Module) -> None:
        for ema_v, model_v in zip(self.module.state_dict().values(), model.state_dict().values()):
            if ema_v.is_cuda and ema_v.dtype == torch.float32 and ema_v.is_contiguous():
                _abi.ema_update(ema_v, model_v, self.decay)
            else:
                raise RuntimeError("the EMA update runs on CUDA fp32 parameters (no CPU fallback)")
        if hasattr(self.module, "refresh"):
            self.module.refresh()                # the packed native weights of the EMA copy are stale now


class EMACallback(_CallbackBase):
    def __init__(self, config):
        super().__init__()
        self.decay = config.ema_decay
        self.use_ema = config.use_ema_weights
        self.ema_model: Optional[ModelEma] = None
        self.model_backup = None
        self.ema_state_to_load = None

    @staticmethod
    def _device_of(module: torch.nn.Module) -> torch.device:
        try:
            return next(module.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def on_fit_start(self, trainer, pl_module) -> None:
        if self.use_ema:
            self.ema_model = ModelEma(pl_module.model, decay=self.decay).to(self._device_of(pl_module.model))
            if self.ema_state_to_load:
                dev = self._device_of(pl_module.model)
                self.ema_model.module.load_state_dict({k: v.to(dev) for k, v in self.ema_state_to_load.items()})
                self.ema_state_to_load = None

    def on_load_checkpoint(self, trainer, pl_module, callback_state: dict) -> None:
        if self.use_ema and callback_state and "ema_state_dict" in callback_state:
            self.ema_state_to_load = callback_state["ema_state_dict"]

    def state_dict(self) -> dict:
        """What lands under ``checkpoint['callbacks']['EMACallback']`` (the reference writes it in ``MMF.on_save_checkpoint``)."""
        if self.ema_model is None:
            return {}
        return {"ema_state_dict": {k: v.detach().cpu() for k, v in self.ema_model.module.state_dict().items()}}

    def on_train_batch_end(self, trainer, pl_module, outputs=None, batch=None, batch_idx=0) -> None:
        if self.use_ema and self.ema_model is not None:
            self.ema_model.update(pl_module.model)

    def _swap_in(self, pl_module) -> None:
        self.model_backup = pl_module.model
        pl_module.model = self.ema_model.module

    def _swap_out(self, pl_module) -> None:
        if self.use_ema and self.model_backup is not None:
            pl_module.model = self.model_backup
            self.model_backup = None

    def on_validation_epoch_start(self, trainer, pl_module) -> None:
        if self.use_ema and self.ema_model is not None:
            self._swap_in(pl_module)

    def on_validation_epoch_end(self, trainer, pl_module) -> None:
        self._swap_out(pl_module)

    def on_predict_start(self, trainer, pl_module) -> None:
        if not self.use_ema:
            return
        if self.ema_model is None:
            self.ema_model = ModelEma(pl_module.model, decay=self.decay).to(self._device_of(pl_module.model))
        cached = getattr(pl_module, "ema_state_from_ckpt", None)
        if not cached:
            return                               # (the reference warns and samples with the trained weights)
        dev = self._device_of(pl_module.model)
        self.ema_model.module.load_state_dict({k: v.to(dev) for k, v in cached.items()})
        self._swap_in(pl_module)

    def on_predict_end(self, trainer, pl_module) -> None:
        self._swap_out(pl_module)
