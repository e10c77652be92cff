"""Output side of a run: per-jet records, the single gather, and the generated_sample.h5 layout
(reference utils/callbacks.py:14-62, utils/tensorclass.py:197-201)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmf_b200 import synthetic
from mmf_b200.distributed import gather_records, shard_bounds
from mmf_b200.tensorclass import TensorMultiModal
from mmf_b200.writer import records_to_arrays, write_generated_sample
from oracle import sample_oracle


def _sample(n_jets, seed=5):
    src = synthetic.source_state(n_jets, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(n_jets, 150, 3, generator=g)              # padded slots deliberately NOT zero: the reference lets them evolve
    k = torch.randint(0, 9, (n_jets, 150, 1), generator=g)
    return TensorMultiModal(time=torch.full((n_jets,), 1 - 1e-5), continuous=x, discrete=k, mask=src.mask)


def test_record_layout_and_host_view():
    s = _sample(7)
    mean, std = [0.5, -0.25, 2.0], [3.0, 0.5, 1.5]
    rec = sample_oracle.pack_records(s.continuous.numpy(), s.discrete.numpy(), s.mask.numpy(), mean, std)
    assert rec.shape == (7, 1952) and sample_oracle.record_bytes(150) == 1952
    x, k, m = records_to_arrays(rec, 150)
    xr, kr, mr = sample_oracle.postprocess(s.continuous.numpy(), s.discrete.numpy(), s.mask.numpy(), mean, std)
    assert x.dtype == np.float32 and k.dtype == np.int64 and m.dtype == np.int64
    assert np.array_equal(x, xr) and np.array_equal(k, kr) and np.array_equal(m, mr)
    # the reference's own arithmetic: continuous * sig + mu, then apply_mask (callbacks.py:52-57)
    ref = TensorMultiModal(time=s.time, continuous=s.continuous * torch.tensor(std) + torch.tensor(mean), discrete=s.discrete.clone(), mask=s.mask)
    ref.apply_mask()
    assert torch.equal(torch.from_numpy(x), ref.continuous) and torch.equal(torch.from_numpy(k), ref.discrete)


def test_generated_sample_file_layout(tmp_path):
    """Datasets, shapes and dtypes of generated_sample.h5 as TensorMultiModal.load_from reads them (npz when h5py is absent)."""
    s = _sample(5)
    rec = sample_oracle.pack_records(s.continuous.numpy(), s.discrete.numpy(), s.mask.numpy())
    path = write_generated_sample(os.path.join(tmp_path, "generated_sample.h5"), rec, 150, 1 - 1e-5)
    back = TensorMultiModal.load_from(path)
    assert back.time.shape == (5,) and back.time.dtype == torch.float32
    assert back.continuous.shape == (5, 150, 3) and back.continuous.dtype == torch.float32
    assert back.discrete.shape == (5, 150, 1) and back.discrete.dtype == torch.int64
    assert back.mask.shape == (5, 150, 1) and back.mask.dtype == torch.int64
    ref = s.clone()
    ref.apply_mask()
    assert torch.equal(back.continuous, ref.continuous) and torch.equal(back.discrete, ref.discrete) and torch.equal(back.mask, ref.mask)
    # the same datasets the reference's save_to writes (tensorclass.py:197-201)
    ref_path = ref.save_to(os.path.join(tmp_path, "reference_layout.h5"))
    a, b = TensorMultiModal.load_from(path), TensorMultiModal.load_from(ref_path)
    for name in ("time", "continuous", "discrete", "mask"):
        assert getattr(a, name).shape == getattr(b, name).shape and getattr(a, name).dtype == getattr(b, name).dtype, name


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_jets, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = _sample(n_jets)
    lo, hi = shard_bounds(n_jets, rank, world)
    rec = torch.from_numpy(sample_oracle.pack_records(s.continuous[lo:hi].numpy(), s.discrete[lo:hi].numpy(), s.mask[lo:hi].numpy()))
    counts = [shard_bounds(n_jets, r, world)[1] - shard_bounds(n_jets, r, world)[0] for r in range(world)]
    out = gather_records(rec, counts)
    torch.save(out, os.path.join(out_dir, f"rec{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_jets", [9, 8, 1])
def test_single_collective_gathers_ragged_shards(tmp_path, n_jets):
    """world_size 2 over gloo: ONE all_gather_into_tensor of the records reproduces the global sample in jet order."""
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_jets, str(tmp_path)), nprocs=2, join=True)
    s = _sample(n_jets)
    want = torch.from_numpy(sample_oracle.pack_records(s.continuous.numpy(), s.discrete.numpy(), s.mask.numpy()))
    for r in range(2):
        got = torch.load(os.path.join(tmp_path, f"rec{r}.pt"))
        assert torch.equal(got, want), r


@pytest.mark.gpu
def test_pack_kernel_matches_oracle_bytes_and_roundtrips():
    from mmf_b200 import _abi
    s = _sample(33)
    mean, std = [0.5, -0.25, 2.0], [3.0, 0.5, 1.5]
    dev = "cuda:0"
    rec = _abi.pack_sample(s.continuous.to(dev), s.discrete.to(dev), s.mask.to(dev), mean, std)
    want = sample_oracle.pack_records(s.continuous.numpy(), s.discrete.numpy(), s.mask.numpy(), mean, std)
    assert rec.shape == want.shape and np.array_equal(rec.cpu().numpy(), want)
    x, k, m = _abi.unpack_sample(rec, 150)
    xr, kr, mr = sample_oracle.postprocess(s.continuous.numpy(), s.discrete.numpy(), s.mask.numpy(), mean, std)
    assert np.array_equal(x.cpu().numpy(), xr) and np.array_equal(k.cpu().numpy()[..., None], kr) and np.array_equal(m.cpu().numpy()[..., None], mr)
    # EPiC: no tokens
    rec2 = _abi.pack_sample(s.continuous.to(dev), None, s.mask.to(dev))
    assert np.array_equal(rec2.cpu().numpy(), sample_oracle.pack_records(s.continuous.numpy(), None, s.mask.numpy()))
    x2, k2, m2 = _abi.unpack_sample(rec2, 150, discrete=False)
    assert k2 is None and torch.equal(m2.cpu(), s.mask.squeeze(-1))


@pytest.mark.gpu
def test_flow_generator_callback_writes_the_reference_layout(tmp_path):
    """FlowGeneratorCallback(config): collect device batches, pack + gather + write on predict end."""
    from types import SimpleNamespace
    from mmf_b200.writer import FlowGeneratorCallback
    cfg = SimpleNamespace(dir=str(tmp_path), project="proj", experiment_id="exp1", tag="t0", time_eps=1e-5,
                          metadata={"mean": [1.0, 2.0, 3.0], "std": [2.0, 2.0, 0.5]})
    os.makedirs(os.path.join(tmp_path, "proj", "exp1"))
    cb = FlowGeneratorCallback(cfg)
    cb.on_predict_start()
    parts = [_sample(6, seed=1), _sample(3, seed=2)]
    for i, p in enumerate(parts):
        cb.on_predict_batch_end(None, None, p.to("cuda:0"), None, i)
    path = cb.on_predict_end()
    assert os.path.dirname(path).endswith(os.path.join("proj", "exp1", "generation_results_t0"))
    assert os.path.exists(os.path.join(os.path.dirname(path), "configs.yaml"))
    back = TensorMultiModal.load_from(path)
    ref = TensorMultiModal.cat(parts)
    ref.continuous = ref.continuous * torch.tensor(cfg.metadata["std"]) + torch.tensor(cfg.metadata["mean"])
    ref.apply_mask()
    assert torch.equal(back.continuous, ref.continuous) and torch.equal(back.discrete, ref.discrete) and torch.equal(back.mask, ref.mask)
    assert back.time.shape == (9,)
