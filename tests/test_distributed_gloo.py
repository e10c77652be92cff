"""world_size-2 `gloo` run of the jet-sharded sampler driver (host logic of SURVEY.md 8(e)) on CPU.

The CUDA sampler cannot run here, so the per-batch function is a CPU stand-in with the same contract: its
output depends only on (source slice, GLOBAL jet index), like the Philox keying of the real kernels.  The
test proves the sharding + single end-of-run gather reproduce the single-process result exactly.
"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmf_b200 import synthetic
from mmf_b200.distributed import generate_sharded
from mmf_b200.tensorclass import TensorMultiModal


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_batch(src: TensorMultiModal, g0: int) -> TensorMultiModal:
    """Stand-in sampler: per-jet 'draws' keyed on the global jet index only."""
    B = len(src)
    gid = torch.arange(g0, g0 + B, dtype=torch.float32)
    x = src.continuous + torch.sin(gid)[:, None, None] * src.mask
    k = (src.discrete + (torch.arange(g0, g0 + B) % 7)[:, None, None] * src.mask) % 9
    return TensorMultiModal(time=torch.full((B,), 1.0), continuous=x, discrete=k, mask=src.mask)


def _worker(rank, world, port, n_jets, batch, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    src = synthetic.source_state(n_jets, seed=77)
    calls = []

    def rb(s, g0):
        calls.append((g0, len(s)))
        return _run_batch(s, g0)

    out = generate_sharded(rb, src, batch)
    torch.save({"out": out, "calls": calls}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_is_world_size_invariant(tmp_path):
    n_jets, batch, world = 37, 8, 2                      # ragged: shards of 19 and 18 jets, partial last batches
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_jets, batch, str(tmp_path)), nprocs=world, join=True)
    src = synthetic.source_state(n_jets, seed=77)
    single = generate_sharded(_run_batch, src, batch)    # world size 1, no process group
    seen = []
    for r in range(world):
        blob = torch.load(os.path.join(tmp_path, f"r{r}.pt"), weights_only=False)
        out = blob["out"]
        for name in ("time", "continuous", "discrete", "mask"):
            assert torch.equal(getattr(out, name), getattr(single, name)), (r, name)
        seen += blob["calls"]
    # every jet generated exactly once across the ranks, batches never cross a shard boundary
    covered = sorted(j for g0, b in seen for j in range(g0, g0 + b))
    assert covered == list(range(n_jets))
    assert all(b <= batch for _, b in seen)


def test_more_ranks_than_jets(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 1, 4, str(tmp_path)), nprocs=2, join=True)
    out = torch.load(os.path.join(tmp_path, "r1.pt"), weights_only=False)["out"]
    assert len(out) == 1
