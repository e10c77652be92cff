"""Boundary value types of the generation hot path.

These mirror the *interface* of the reference containers so that code written
against the reference keeps working when it is pointed at this package:

  * ``TensorMultiModal``  <- reference ``utils/tensorclass.py:12-250``
  * ``DataCoupling``      <- reference ``utils/datasets.py:8-41``

Fields, shapes and dtypes (reference ``utils/tensorclass.py:13-16``):

  time        (B,)       float32
  continuous  (B, D, 3)  float32   (pT, eta_rel, phi_rel), standardised
  discrete    (B, D, 1)  int64     token in 0..V-1, 0 = pad
  mask        (B, D, 1)  int64     1 = real particle

Only the container behaviour is reproduced here (it is host-side glue); the
arithmetic of the hot path lives in the CUDA library behind ``mmf_b200._abi``.
"""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Callable, List, Optional

import torch

_MODES = ("time", "continuous", "discrete")          # order matters for __len__/ndim
_ALL = ("time", "continuous", "discrete", "mask")


@dataclass
class TensorMultiModal:
    time: Optional[torch.Tensor] = None
    continuous: Optional[torch.Tensor] = None
    discrete: Optional[torch.Tensor] = None
    mask: Optional[torch.Tensor] = None

    # ---- introspection -------------------------------------------------
    def available_modes(self) -> List[str]:
        """Non-empty modes in the fixed order time, continuous, discrete
        (reference ``tensorclass.py:189-195``; ``mask`` is not a mode)."""
        return [m for m in _MODES if getattr(self, m) is not None]

    def _last(self) -> Optional[torch.Tensor]:
        modes = self.available_modes()
        return getattr(self, modes[-1]) if modes else None

    @property
    def ndim(self) -> int:
        last = self._last()
        return 0 if last is None else last.dim()

    @property
    def shape(self):
        last = self._last()
        return None if last is None else last.shape[:-1]

    def __len__(self) -> int:
        last = self._last()
        return 0 if last is None else len(last)

    @property
    def has_continuous(self) -> bool:
        return self.continuous is not None

    @property
    def has_discrete(self) -> bool:
        return self.discrete is not None

    # ---- element-wise plumbing ------------------------------------------
    def _map(self, fn: Callable[[torch.Tensor], torch.Tensor]) -> "TensorMultiModal":
        kw = {}
        for name in _ALL:
            val = getattr(self, name)
            kw[name] = fn(val) if isinstance(val, torch.Tensor) else None
        return TensorMultiModal(**kw)

    def to(self, device) -> "TensorMultiModal":
        return self._map(lambda x: x.to(device))

    def cpu(self) -> "TensorMultiModal":
        return self._map(lambda x: x.cpu())

    def detach(self) -> "TensorMultiModal":
        return self._map(lambda x: x.detach())

    def clone(self) -> "TensorMultiModal":
        return self._map(lambda x: x.clone())

    def pin_memory(self) -> "TensorMultiModal":
        return self._map(lambda x: x.pin_memory())

    def __getitem__(self, idx) -> "TensorMultiModal":
        return self._map(lambda x: x[idx])

    def _op(self, op, *args, mode: Optional[str] = None, **kw) -> "TensorMultiModal":
        if mode is not None and mode not in _ALL:
            raise ValueError(f"Invalid mode '{mode}'. Choose from {list(_ALL)}")
        out = {}
        for name in _ALL:
            val = getattr(self, name)
            hit = val is not None and (mode is None or mode == name)
            out[name] = op(val, *args, **kw) if hit else val
        return TensorMultiModal(**out)

    def unsqueeze(self, dim: int, mode: Optional[str] = None):
        return self._op(torch.unsqueeze, dim, mode=mode)

    def squeeze(self, dim: Optional[int] = None, mode: Optional[str] = None):
        if dim is None:
            return self._op(torch.squeeze, mode=mode)
        return self._op(torch.squeeze, dim, mode=mode)

    def reshape(self, *shape, mode: Optional[str] = None):
        return self._op(torch.reshape, shape, mode=mode)

    def expand(self, *sizes, mode: Optional[str] = None):
        return self._op(torch.Tensor.expand, *sizes, mode=mode)

    def repeat(self, *reps, mode: Optional[str] = None):
        return self._op(torch.Tensor.repeat, *reps, mode=mode)

    def broadcast_time(self) -> None:
        """(B,1) -> (B,D,1), reference ``tensorclass.py:90-95``."""
        D = self.shape[-1]
        self.time = self.time.unsqueeze(1).repeat(1, D, 1)

    def apply_mask(self, condition=None, include_time: bool = False) -> None:
        """Zero the padded slots in place (reference ``tensorclass.py:97-108``)."""
        m = self.mask if condition is None else condition
        if include_time and self.time is not None:
            self.time *= m
        if self.continuous is not None:
            self.continuous *= m
        if self.discrete is not None:
            self.discrete = (self.discrete * m).long()

    # ---- batching --------------------------------------------------------
    @staticmethod
    def _join(states, fn, dim):
        kw = {}
        for name in _ALL:
            parts = [getattr(s, name, None) for s in states]
            parts = [p for p in parts if p is not None]
            kw[name] = fn(parts, dim=dim) if parts else None
        return TensorMultiModal(**kw)

    @staticmethod
    def cat(states: List["TensorMultiModal"], dim: int = 0) -> "TensorMultiModal":
        return TensorMultiModal._join(states, torch.cat, dim)

    @staticmethod
    def stack(states: List["TensorMultiModal"], dim: int = 0) -> "TensorMultiModal":
        return TensorMultiModal._join(states, torch.stack, dim)

    # ---- files -----------------------------------------------------------
    # The reference writes HDF5 with one dataset per mode plus ``mask``
    # (``tensorclass.py:197-201``).  h5py is an optional dependency here: it is
    # used when importable, otherwise the same dataset names go to a ``.npz``.
    def save_to(self, path: str) -> str:
        data = {m: getattr(self, m).detach().cpu().numpy() for m in self.available_modes()}
        if self.mask is not None:
            data["mask"] = self.mask.detach().cpu().numpy()
        try:
            import h5py  # type: ignore
            if not hasattr(h5py, "File"):          # an inert stub module (test harnesses install one)
                raise ImportError("h5py stub")
        except Exception:
            import numpy as np
            out = path if path.endswith(".npz") else path + ".npz"
            np.savez(out, **data)
            return out
        with h5py.File(path, "w") as f:
            for key, arr in data.items():
                f.create_dataset(key, data=arr)
        return path

    @classmethod
    def load_from(cls, path: str, device=None, transform=None) -> "TensorMultiModal":
        tensors = {k: None for k in _ALL}
        if path.endswith(".npz"):
            import numpy as np
            with np.load(path) as f:
                for key in _ALL:
                    if key in f.files:
                        tensors[key] = torch.from_numpy(f[key])
        else:
            import h5py  # type: ignore
            with h5py.File(path, "r") as f:
                for key in _ALL:
                    if key in f:
                        tensors[key] = torch.from_numpy(f[key][:])
        if callable(transform):
            tensors = {k: (transform(v) if v is not None else None) for k, v in tensors.items()}
        elif isinstance(transform, dict):
            for key, fn in transform.items():
                if tensors.get(key) is not None and callable(fn):
                    tensors[key] = fn(tensors[key])
        state = cls(**tensors)
        return state.to(device) if device else state


@dataclass
class DataCoupling:
    """source / target / context triple (reference ``utils/datasets.py:8-41``)."""
    source: Optional[TensorMultiModal] = None
    target: Optional[TensorMultiModal] = None
    context: Optional[TensorMultiModal] = None

    def __len__(self) -> int:
        return len(self.target)

    @property
    def ndim(self) -> int:
        return self.target.ndim

    @property
    def shape(self):
        return self.target.shape

    @property
    def has_source(self) -> bool:
        return bool(self.source)

    @property
    def has_target(self) -> bool:
        return bool(self.target)

    @property
    def has_context(self) -> bool:
        return bool(self.context)


def field_names():
    return [f.name for f in fields(TensorMultiModal)]
