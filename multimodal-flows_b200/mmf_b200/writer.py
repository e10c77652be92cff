"""Output side of a generation run: the replacement of the reference's ``FlowGeneratorCallback``.

reference ``utils/callbacks.py:14-62``                      here
  on_predict_batch_end: keep every batch (host tensors)       keep every batch ON THE DEVICE
  _save_results_local:  one temp ``.h5`` per rank             --
  barrier + rank 0 globs and re-reads the temp files          ONE collective of narrow per-jet records
  continuous * std + mean, apply_mask on the host             fused into the record kernel (``mmf_pack_sample``)
  ``generated_sample.h5`` via ``TensorMultiModal.save_to``    ``write_generated_sample``: same datasets, shapes, dtypes

File layout (reference ``utils/tensorclass.py:197-201``, ``utils/callbacks.py:44-58``):

  <dir>/<project>/<experiment_id>/generation_results<_tag>/configs.yaml
  <dir>/<project>/<experiment_id>/generation_results<_tag>/generated_sample.h5
      time (N,) f32 | continuous (N,D,3) f32 de-standardised, pads zero | discrete (N,D,1) i64 | mask (N,D,1) i64

``h5py`` is an optional dependency (absent in the build container): when it cannot be imported the same dataset names,
shapes and dtypes go to ``generated_sample.h5.npz`` and ``TensorMultiModal.load_from`` reads either.
"""
from __future__ import annotations

import os
from typing import List, Optional

import numpy as np
import torch

from . import _abi
from .distributed import gather_records, shard_bounds, world
from .tensorclass import TensorMultiModal

try:                                            # pragma: no cover - not installed in the build container
    from pytorch_lightning import Callback as _CallbackBase
except Exception:                               # noqa: BLE001
    _CallbackBase = object


def records_to_arrays(rec: np.ndarray, D: int, discrete: bool = True):
    """Host view of (N, R) uint8 records as the reference's arrays: continuous (N,D,3) f32, discrete (N,D,1) i64 or None,
    mask (N,D,1) i64.  Pure re-interpretation of bytes plus the widening of the token / mask byte."""
    N = rec.shape[0]
    x = np.ascontiguousarray(rec[:, : D * 12]).view(np.float32).reshape(N, D, 3)
    kb = rec[:, D * 12: D * 13]
    k = (kb & 0x7F).astype(np.int64)[..., None] if discrete else None
    mask = (kb >> 7).astype(np.int64)[..., None]
    return x, k, mask


def write_generated_sample(path: str, rec, D: int, time_value: float, discrete: bool = True) -> str:
    """Write the gathered records in the ``generated_sample.h5`` layout.  Returns the path written."""
    if isinstance(rec, torch.Tensor):
        rec = rec.detach().cpu().numpy()
    x, k, mask = records_to_arrays(rec, D, discrete)
    data = {"time": np.full((rec.shape[0],), time_value, np.float32), "continuous": x}
    if k is not None:
        data["discrete"] = k
    data["mask"] = mask
    try:
        import h5py  # type: ignore
        if not hasattr(h5py, "File"):
            raise ImportError("h5py stub")
    except Exception:                           # noqa: BLE001
        out = path if path.endswith(".npz") else path + ".npz"
        np.savez(out, **data)
        return out
    with h5py.File(path, "w") as f:
        for key, arr in data.items():
            f.create_dataset(key, data=arr)
    return path


class FlowGeneratorCallback(_CallbackBase):
    """Drop-in for the reference callback of the same name (same constructor argument and hooks)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.experiment_dir = os.path.join(str(config.dir), str(config.project), str(config.experiment_id))
        self.tag = f"_{config.tag}" if getattr(config, "tag", None) else ""
        self.batched_data: List[TensorMultiModal] = []
        self.device = None

    def on_predict_start(self, trainer=None, pl_module=None):
        self.batched_data = []
        if pl_module is not None:
            self.device = pl_module.device

    def on_predict_batch_end(self, trainer, pl_module, outputs, batch=None, batch_idx=0, dataloader_idx=0):
        self.batched_data.append(outputs)

    def finalize(self, num_jets_total: Optional[int] = None) -> Optional[str]:
        """Pack this rank's batches, gather once, write on rank 0.  ``num_jets_total`` gives the shard sizes of a
        ``generate_sharded`` run; without it every rank is taken to hold the same number of jets."""
        rank, ws = world()
        dev = self.device or torch.device("cuda", torch.cuda.current_device())
        local = TensorMultiModal.cat([b.to(dev) for b in self.batched_data], dim=0)
        md = getattr(self.config, "metadata", None) or {}
        rec = _abi.pack_sample(local.continuous, local.discrete, local.mask, md.get("mean"), md.get("std"))
        n_local = rec.shape[0]
        counts = ([shard_bounds(num_jets_total, r, ws)[1] - shard_bounds(num_jets_total, r, ws)[0] for r in range(ws)]
                  if num_jets_total is not None else [n_local] * ws)
        rec = gather_records(rec, counts)
        if rank != 0:
            return None
        out_dir = os.path.join(self.experiment_dir, f"generation_results{self.tag}")
        os.makedirs(out_dir, exist_ok=False)             # the reference's os.mkdir fails on an existing directory too
        import yaml
        with open(os.path.join(out_dir, "configs.yaml"), "w") as f:
            yaml.dump(dict(vars(self.config)), f, sort_keys=False)
        D = local.continuous.shape[1]
        t_end = float(local.time[0]) if local.time is not None and local.time.numel() else 1.0 - float(getattr(self.config, "time_eps", 1e-5))
        return write_generated_sample(os.path.join(out_dir, "generated_sample.h5"), rec, D, t_end, discrete=local.discrete is not None)

    def on_predict_end(self, trainer=None, pl_module=None):
        return self.finalize()
