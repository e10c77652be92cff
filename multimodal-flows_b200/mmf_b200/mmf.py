"""Generative-model front-ends with the reference's entry points.

  MultiModalFlowBridge      reference model/MMF.py:20-200   (ParticleFormer / FusedParticleFormer)
  ConditionalFlowMatching   reference model/CFM.py:13-154   (EPiC carrier)

``simulate_dynamics`` / ``predict_step`` / ``forward`` keep the reference signatures.  The N-step loop
is ONE call into libmmf_b200.so: no per-step host synchronisation (the reference has three per step,
SURVEY.md section 9).  Lightning is optional: when ``pytorch_lightning`` is importable the classes
derive from ``LightningModule`` so ``Trainer.predict`` drives them unchanged; otherwise they are
plain ``nn.Module`` objects and ``load_from_checkpoint`` reads the ``.ckpt`` with ``torch.load``.

Training: ``MultiModalFlowBridge.loss`` / ``validation_step`` run the forward half of the reference's training step
(bridge sampling, encoder forward, masked MSE + CE, MultiTaskLoss) through the library (``mmf_bridge_sample``,
``mmf_encoder_forward``, ``mmf_multitask_loss``).  After ``configure_training()`` ``training_step`` runs the whole step on
the device - forward, backward, DDP gradient average, norm clipping, Adam - through ``mmf_b200.training.TrainEngine``
(``include/mmf_b200_train.h``); there is no torch autograd on the path, so under Lightning the module declares
``automatic_optimization = False``.
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
        lc = {k[len("loss_combine."):]: v for k, v in ckpt["state_dict"].items() if k.startswith("loss_combine.")}
        if lc and hasattr(obj, "loss_combine"):
            obj.loss_combine.load_state_dict(lc, strict=strict)
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


class MultiTaskLoss(nn.Module):
    """Parameter shell of the reference's loss combiner (model/MMF.py:203-233): ``uncertainty_net`` = MLP(n_embd, n_embd, 2)
    for "time-weighted" (``c_proj.bias`` starts at 0: balanced L = L_mse + L_ce), ``loss_weights`` for "weighted", nothing for
    "sum".  Same state_dict keys as the reference (``loss_combine.uncertainty_net.c_fc.weight`` ...); the arithmetic runs in
    ``mmf_multitask_loss``."""

    def __init__(self, config):
        super().__init__()
        self.mode = config.multitask_loss
        E = config.n_embd
        if self.mode == "weighted":
            self.loss_weights = nn.Parameter(torch.tensor([0.0, 0.0]))
        elif self.mode == "time-weighted":
            net = nn.Module()
            net.c_fc = nn.Linear(E, E)
            net.c_proj = nn.Linear(E, 2)
            nn.init.constant_(net.c_proj.bias, 0.0)
            self.uncertainty_net = net
        elif self.mode != "sum":
            raise ValueError(f"unknown multitask_loss '{self.mode}'")

    def net_tensors(self):
        n = self.uncertainty_net
        return n.c_fc.weight, n.c_fc.bias, n.c_proj.weight, n.c_proj.bias


class MultiModalFlowBridge(_GenerativeBase):
    """Hybrid continuous/discrete sampler (Euler ODE + telegraph tau-leap)."""

    def __init__(self, config):
        super().__init__(config)
        self.loss_combine = MultiTaskLoss(self.config)

    # ---- forward half of the training step (reference model/MMF.py:42-68, 138-170) ----------------------------------
    @torch.no_grad()
    def loss(self, batch: DataCoupling, time: Optional[torch.Tensor] = None, z: Optional[torch.Tensor] = None,
             u: Optional[torch.Tensor] = None):
        """(loss, loss_mse, loss_ce, w_mse, w_ce) exactly as the reference returns them.  ``time`` (B,), ``z`` (B,D,3) and
        ``u`` (B,D) replace the reference's torch.rand / randn_like / Categorical.sample draws when given (parity tests);
        otherwise time comes from torch's generator like the reference and the bridge noise from the library's Philox."""
        cfg = self.config
        dev = self.device
        B, V, eps = len(batch), cfg.vocab_size, cfg.time_eps
        if cfg.multitask_loss == "weighted":
            raise NotImplementedError("loss() covers 'time-weighted' and 'sum'; multitask_loss='weighted' runs through configure_training() "
                                      "(training_step / validation_step)")
        if time is None:
            time = eps + (1.0 - eps) * torch.rand(B, device=dev)
        tgt, src = batch.target.to(dev), batch.source.to(dev) if batch.source is not None else TensorMultiModal()
        if not src.has_continuous:                       # reference model/CFM.py:175-177
            src.continuous = torch.randn_like(tgt.continuous) * tgt.mask
        if not src.has_discrete:                         # reference model/MJB.py:201-203
            src.discrete = torch.randint_like(tgt.discrete, 1, V) * tgt.mask
        xt, kt = _abi.bridge_sample(src.continuous, tgt.continuous, src.discrete, tgt.discrete, time.to(dev), cfg.sigma, cfg.beta, V,
                                    z=None if z is None else z.to(dev), u=None if u is None else u.to(dev), seed=self.seed,
                                    first_global_jet=self._jet_cursor)
        self._jet_cursor += B
        state = TensorMultiModal(continuous=xt, discrete=kt, mask=tgt.mask, time=time.to(dev))
        vt, logits = self.model(state)
        net = self.loss_combine.net_tensors() if cfg.multitask_loss == "time-weighted" else None
        out, _ = _abi.multitask_loss(vt, logits, src.continuous, tgt.continuous, tgt.discrete, tgt.mask, time.to(dev),
                                     cfg.multitask_loss, cfg.n_embd, net)
        if cfg.multitask_loss == "sum":
            return out[0], out[1], out[2], None, None
        return out[0], out[1], out[2], out[3], out[4]

    def _log(self, name, value, **kw):
        if value is not None and _Base is not nn.Module:      # pragma: no cover - Lightning only
            self.log(name, value, **kw)

    def configure_training(self, lr: Optional[float] = None, max_norm: float = 1.0, use_graphs: bool = True):
        """Creates the device training engine (reference configure_optimizers model/MMF.py:77-78: Adam(lr=config.lr); Trainer
        gradient_clip_val=1.0 scripts/train_mmf.py:166).  From here on the parameters live in the engine's flat buffers."""
        from .training import TrainEngine
        self.automatic_optimization = False                 # Lightning: the step owns backward and the optimiser
        self._engine = TrainEngine(self, lr=lr, max_norm=max_norm, use_graphs=use_graphs)
        return self._engine

    def training_step(self, batch: DataCoupling, batch_idx: int = 0, lr: Optional[float] = None):
        engine = getattr(self, "_engine", None)
        if engine is not None:
            out = engine.train_step(batch, lr=lr)
            weighted = self.config.multitask_loss != "sum"
            loss, loss_mse, loss_ce, w_mse, w_ce = out[0], out[1], out[2], out[3] if weighted else None, out[4] if weighted else None
        else:
            loss, loss_mse, loss_ce, w_mse, w_ce = self.loss(batch)
        for name, v in (("train_loss", loss), ("train_loss_ce", loss_ce), ("train_loss_mse", loss_mse), ("train_weight_mse", w_mse),
                        ("train_weight_ce", w_ce)):
            self._log(name, v, on_epoch=True, sync_dist=True, batch_size=len(batch))
        return {"loss": loss}

    def validation_step(self, batch: DataCoupling, batch_idx: int = 0):
        engine = getattr(self, "_engine", None)
        if engine is not None and self.model is engine.model_ref:   # (an EMA module swapped in by EMACallback validates through loss())
            # the training engine's forward (any multitask_loss mode), no gradients
            out = engine.loss_only(batch)
            weighted = self.config.multitask_loss != "sum"
            loss, loss_mse, loss_ce, w_mse, w_ce = out[0], out[1], out[2], out[3] if weighted else None, out[4] if weighted else None
        else:
            loss, loss_mse, loss_ce, w_mse, w_ce = self.loss(batch)
        for name, v in (("val_loss", loss), ("val_loss_ce", loss_ce), ("val_loss_mse", loss_mse), ("val_weight_mse", w_mse),
                        ("val_weight_ce", w_ce)):
            self._log(name, v, on_epoch=True, sync_dist=True, batch_size=len(batch))
        return {"val_loss": loss}

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
