"""Whole device-side pipeline of a sampling run (scripts/sample_mmf.py without the file I/O), per GPU:
   source on the device (mmf_make_source) -> N-step sampler (mmf_generate) -> jet observables (mmf_jet_observables),
in batches, with NO host<->device copy of the sample; only the per-batch multiplicities travel (tile planning).
usage: pipeline_rate.py [model] [total_jets] [batch] [timesteps]      (BASELINE config #4: EPiC 1048576 16384 100)
Under torchrun (one rank per GPU) the global jet index is sharded in contiguous slices, every draw is keyed on the GLOBAL jet
index (the sample does not depend on the world size), and ONE all_gather of the per-jet observables ends the run; the time is
the max over ranks of the device time, gather included."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-flows_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from mmf_b200 import _abi, synthetic
from mmf_b200.param_spec import make_config
from mmf_b200.mmf import time_grid

model = sys.argv[1] if len(sys.argv) > 1 else "EPiC"
total = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
N = int(sys.argv[4]) if len(sys.argv) > 4 else 100
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=dev)
from mmf_b200.distributed import shard_bounds
lo, hi = shard_bounds(total, rank, world)
cfg = make_config(model, num_timesteps=N)
sd = synthetic.make_state_dict(cfg, "wide", 0)
nm = _abi.NativeModel(cfg, sd, dev)
ts, dt = time_grid(cfg)
D, V = cfg.max_num_particles, cfg.vocab_size
n = np.clip(np.round(55 + 18 * np.random.default_rng(0).standard_normal(200000)), 1, D).astype(int)     # SURVEY 8(d) stand-in
probs = (np.bincount(n, minlength=D + 1) / len(n)).astype(np.float32)
epic = model == "EPiC"

def run(first, B):
    x0, k0, mask, _ = _abi.make_source(probs, B, D, V, 1, first, dev, discrete=not epic)
    opts = None if epic else _abi.step_options(cfg, seed=3, first_global_jet=first)
    x, k, _ = nm.generate(x0, k0, mask, ts, dt, opts)
    return _abi.jet_observables(x, k, mask, [1.9, 0.0, 0.0], [0.8, 0.11, 0.1], V)

run(0, batch); torch.cuda.synchronize()
if world > 1:
    torch.distributed.barrier()
mass_hist = torch.zeros(64, device=dev)
mult_hist = torch.zeros(D + 1, device=dev)
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
done = lo
kins = []
while done < hi:
    B = min(batch, hi - done)
    kin, counts = run(done, B)
    kins.append(kin)
    done += B
kin = torch.cat(kins) if kins else torch.zeros(0, len(_abi.OBS_COLUMNS), device=dev)
if world > 1:                                   # the one collective of the run: per-jet observables of every shard
    sizes = [shard_bounds(total, r, world) for r in range(world)]
    pad = max(b - a for a, b in sizes)
    buf = torch.zeros(pad, kin.shape[1], device=dev); buf[: kin.shape[0]] = kin
    out = [torch.empty_like(buf) for _ in range(world)]
    torch.distributed.all_gather(out, buf)
    kin = torch.cat([o[: b - a] for o, (a, b) in zip(out, sizes)])
mass_hist += torch.histc(kin[:, 5].nan_to_num(0.0), bins=64, min=0.0, max=64.0)
mult_hist += torch.bincount(kin[:, 10].long(), minlength=D + 1).float()
e1.record(); torch.cuda.synchronize()
wall = time.perf_counter() - t0
tm = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
if world > 1:
    torch.distributed.all_reduce(tm, op=torch.distributed.ReduceOp.MAX)
ms = float(tm.item())
if world > 1:
    torch.distributed.destroy_process_group()
if rank == 0:
  print(json.dumps({"n_gpus": world, "pipeline": "make_source -> generate -> jet_observables, all on the device", "model": model, "jets": total, "batch": batch,
                  "timesteps": N, "device_s": ms * 1e-3, "wall_s": wall, "jets_per_s": total / (ms * 1e-3),
                  "mean_multiplicity": float((mult_hist * torch.arange(D + 1, device=dev)).sum() / mult_hist.sum()),
                  "jets_with_finite_mass": int(mass_hist.sum())}))
