"""Jet-sharded generation across the GPUs of one box.

Jets are independent (attention, pooling and the step are strictly intra-jet), so the sampler shards the
global jet index contiguously over ranks with NO collective inside the time loop; every rank runs the
whole N-step sampler on its slice.  One gather at the very end (``all_gather`` over NCCL on GPUs, gloo in
the CPU tests) replaces the reference's per-rank temp ``.h5`` + barrier + rank-0 glob
(reference ``utils/callbacks.py:27-58``).  Draws are keyed on the GLOBAL jet index
(``MmfStepOptions.first_global_jet``) so the generated sample does not depend on the world size.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist

from .tensorclass import TensorMultiModal

RunBatch = Callable[[TensorMultiModal, int], TensorMultiModal]


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(num_jets: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of the global jet index owned by ``rank``; sizes differ by at most one."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(num_jets, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _gather_rows(t: Optional[torch.Tensor], counts: List[int]) -> Optional[torch.Tensor]:
    """all_gather of a (n_rank, ...) tensor with per-rank row counts ``counts`` (padded to the maximum)."""
    if t is None:
        return None
    rank, ws = world()
    if ws == 1:
        return t
    cap = max(counts)
    pad = torch.zeros((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def gather_sample(local: TensorMultiModal, counts: List[int]) -> TensorMultiModal:
    """The single end-of-run collective: every rank receives the full sample in global jet order."""
    return TensorMultiModal(time=_gather_rows(local.time, counts), continuous=_gather_rows(local.continuous, counts),
                            discrete=_gather_rows(local.discrete, counts), mask=_gather_rows(local.mask, counts))


def gather_records(rec: torch.Tensor, counts: List[int]) -> torch.Tensor:
    """THE end-of-run collective: one ``all_gather_into_tensor`` of the per-jet records of this rank's shard
    ((n_rank, R) uint8, ``_abi.pack_sample``; shards are padded to the largest).  Returns the records of the whole
    sample in global jet order on every rank.  13 bytes per slot instead of the 28 of the fp32 / int64 / int64 tensors."""
    rank, ws = world()
    if ws == 1:
        return rec
    cap, R = max(counts), rec.shape[1]
    pad = rec
    if rec.shape[0] != cap:
        pad = torch.zeros(cap, R, dtype=rec.dtype, device=rec.device)
        pad[: rec.shape[0]] = rec
    out = torch.empty(ws * cap, R, dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, pad.contiguous())
    if all(c == cap for c in counts):
        return out
    return torch.cat([out[r * cap: r * cap + c] for r, c in enumerate(counts)], dim=0)


def gather_packed_sample(local: TensorMultiModal, counts: List[int], mean=None, std=None, unpack: bool = True):
    """De-standardise + mask + narrow on the device (``mmf_pack_sample``), ONE collective, and (optionally) the reference's
    tensors again.  Replaces FlowGeneratorCallback's temp-file merge and its host-side post-processing
    (reference ``utils/callbacks.py:27-58``).  Returns a ``TensorMultiModal`` (``unpack=True``) or the (N, R) records."""
    from . import _abi
    D = local.continuous.shape[1]
    rec = _abi.pack_sample(local.continuous, None if local.discrete is None else local.discrete, local.mask, mean, std)
    rec = gather_records(rec, counts)
    if not unpack:
        return rec
    x, k, mask = _abi.unpack_sample(rec, D, discrete=local.discrete is not None)
    t = local.time
    if t is not None:
        t = torch.full((x.shape[0],), float(t[0]) if t.numel() else 0.0, device=x.device, dtype=t.dtype)
    return TensorMultiModal(time=t, continuous=x, discrete=None if k is None else k.unsqueeze(-1), mask=mask.unsqueeze(-1))


def generate_sharded(run_batch: RunBatch, source: TensorMultiModal, batch_size: int,
                     gather: bool = True) -> TensorMultiModal:
    """Generate ``len(source)`` jets with the global source state replicated on every rank.

    ``run_batch(src_slice, first_global_jet)`` runs the sampler on one batch of this rank's shard (e.g.
    ``lambda s, g0: bridge.simulate_dynamics(DataCoupling(source=s), first_global_jet=g0).target``).
    Returns the gathered sample (global order) on every rank, or the local shard when ``gather=False``.
    """
    rank, ws = world()
    n = len(source)
    lo, hi = shard_bounds(n, rank, ws)
    outs = []
    for b0 in range(lo, hi, batch_size):
        b1 = min(b0 + batch_size, hi)
        outs.append(run_batch(source[b0:b1], b0))
    if outs:
        local = TensorMultiModal.cat(outs, dim=0)
    else:                                   # more ranks than jets: an empty shard with the right trailing shapes
        local = source[0:0].clone()
        if local.continuous is not None and local.time is None:
            local.time = torch.zeros(0, dtype=local.continuous.dtype, device=local.continuous.device)
    if not gather:
        return local
    counts = [shard_bounds(n, r, ws)[1] - shard_bounds(n, r, ws)[0] for r in range(ws)]
    return gather_sample(local, counts)


def generate_from_device_source(run_batch: RunBatch, mult_probs, num_jets: int, batch_size: int, max_num_particles: int = 150,
                                vocab_size: int = 9, time_eps: float = 1e-5, seed: int = 0, device="cuda", discrete: bool = True,
                                gather: bool = True) -> TensorMultiModal:
    """``generate_sharded`` without a host-side source: every batch of this rank's shard is BUILT ON THE DEVICE
    (``mmf_b200.source.make_source``: multiplicities from ``mult_probs``, prefix masks, noise and tokens keyed on
    (seed, GLOBAL jet index, slot)), so nothing but the multiplicities ever crosses PCIe and the sample is the same for every
    batch size and world size.  Replaces the replicated ``_make_source_dataloader`` of ``scripts/sample_mmf.py:70-92``."""
    from .source import make_source
    rank, ws = world()
    lo, hi = shard_bounds(num_jets, rank, ws)
    outs = []
    for b0 in range(lo, hi, batch_size):
        b1 = min(b0 + batch_size, hi)
        src = make_source(mult_probs, b1 - b0, max_num_particles, vocab_size, time_eps, seed=seed, first_global_jet=b0, device=device,
                          discrete=discrete)
        outs.append(run_batch(src, b0))
    if outs:
        local = TensorMultiModal.cat(outs, dim=0)
    else:                                   # more ranks than jets: an empty shard (no library call for zero jets)
        dev, D = torch.device(device), max_num_particles
        local = TensorMultiModal(time=torch.zeros(0, device=dev), continuous=torch.zeros(0, D, 3, device=dev),
                                 discrete=torch.zeros(0, D, 1, dtype=torch.int64, device=dev) if discrete else None,
                                 mask=torch.zeros(0, D, 1, dtype=torch.int64, device=dev))
    if not gather:
        return local
    counts = [shard_bounds(num_jets, r, ws)[1] - shard_bounds(num_jets, r, ws)[0] for r in range(ws)]
    return gather_sample(local, counts)
