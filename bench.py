#!/usr/bin/env python
"""Headline benchmark: generated jets / second at 100 steps (ParticleFormer, <=150 particles).

Contract (see the task description): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.
  step      = one pass of the hot path over one batch: the whole 100-timestep sampler on 256 synthetic
              AOJ-shaped jets per GPU (BASELINE.json configs[1])
  value     = whole-job jets/s with inputs resident in HBM when the timed region starts (CUDA events, max over ranks)
  e2e       = same metric through the drop-in API with HOST buffers (H2D and D2H copies inside the timed region)
  roofline  = the dominant kernel class, timed live with CUDA events, against MEASURED_PEAKS.json
  roofline_dense = the same kernel on the dense worst case (every jet 150 particles: CTA-pair tiles)
  cpu_baseline = the UNMODIFIED reference's own sampler (oracle/_ref, copied by oracle/make_ref.py; the oracle port when that
              copy is absent) on this box's host cores: the WHOLE batch for a bounded number of timesteps
  gpu_eager_baseline = the fp32 oracle port (plain torch, TF32 off) on the same GPU, whole batch, bounded timesteps
  extra_models = FusedParticleFormer / EPiC on one GPU; whole_run = source -> sampler -> records -> ONE collective, timed
              end to end on every rank (BASELINE configs #3 and #4)
  training  = BASELINE config #5: the device training step (forward + backward + gradient all-reduce + clip + Adam) on every rank,
              with the unmodified reference's own step on the host cores and fp32 eager torch on the same GPU beside it
`--impl reference` times the reference's own CPU sampler as the reference arm (rank 0 only): every step is the whole batch
for `--ref-sample-timesteps` timesteps, nothing is extrapolated over jets.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "multimodal-flows_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "generated jets/sec @100 steps (ParticleFormer, 150 particles)"


def metric_name(args):
    return METRIC if args.model == "ParticleFormer" else f"generated jets/sec @{args.timesteps} steps ({args.model}, 150 particles)"
FLOPS_PER_PARTICLE = {"ParticleFormer": 10_630_656, "FusedParticleFormer": 5_649_920, "EPiC": 2_404_960}
FLOPS_PER_JET_CONST = {"ParticleFormer": 65_536, "FusedParticleFormer": 0, "EPiC": 1_794_048}
ATTN_FLOPS_PER_N2 = {"ParticleFormer": 11_264, "FusedParticleFormer": 5_120, "EPiC": 0}
TC_CLASSES = ("gemm_embed", "gemm_qkv", "attention", "gemm_attn_proj_resln", "gemm_mlp_fc_gelu", "gemm_mlp_out_resln",
              "gemm_head_fc_gelu")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="ParticleFormer", choices=["ParticleFormer", "FusedParticleFormer", "EPiC"])
    ap.add_argument("--batch", type=int, default=256, help="jets per GPU per step")
    ap.add_argument("--timesteps", type=int, default=100)
    ap.add_argument("--temperature", type=float, default=1.0)
    ap.add_argument("--dense", action="store_true", help="every jet has 150 particles (worst case)")
    ap.add_argument("--cpu-sample-jets", type=int, default=32,
                    help="jets of the CPU sample (0 = the whole batch).  32 jets is the CPU's best case: its fp32 attention tensors still "
                         "fit in cache, the whole batch of 256 runs ~20x slower per jet (see cpu_baseline.whole_batch_probe)")
    ap.add_argument("--cpu-sample-timesteps", type=int, default=100,
                    help="timesteps of the cpu_baseline sample (32 jets x all 100 timesteps: 5-15 s of CPU work, no extrapolation)")
    ap.add_argument("--ref-sample-timesteps", type=int, default=2,
                    help="timesteps of ONE step of --impl reference / cpu_baseline: the WHOLE batch is stepped through this many grid "
                         "points (about a second each on 16 cores), so K + W steps fit in minutes")
    ap.add_argument("--eager-timesteps", type=int, default=3, help="timesteps of the gpu_eager_baseline sample (whole batch)")
    ap.add_argument("--whole-run-jets", type=int, default=4096, help="jets per rank of the whole_run legs (x4 for EPiC)")
    ap.add_argument("--no-extras", action="store_true", help="skip roofline_dense, extra_models, whole_run, gpu_eager_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-step-roofline", action="store_true")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except Exception:
                continue
            for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6), ("sw_power_cap", 7)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # median over the upper half: samples taken between launches idle down
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_flops_per_timestep(model: str, n: torch.Tensor) -> float:
    nf = n.double()
    return float((FLOPS_PER_PARTICLE[model] * nf + FLOPS_PER_JET_CONST[model] + ATTN_FLOPS_PER_N2[model] * nf * nf).sum())


def cpu_port_rate(args, cfg, sd, sample_jets, sample_timesteps, repeats=1, device="cpu"):
    """jets/s at `timesteps` steps of the oracle port (torch fp32), from `sample_timesteps` timesteps of `sample_jets` jets."""
    from mmf_b200 import synthetic
    from oracle import mmf_oracle as orc
    if device == "cpu":
        torch.set_num_threads(os.cpu_count() or 1)
    src = synthetic.source_state(sample_jets, cfg.max_num_particles, cfg.vocab_size, dense=args.dense).to(device)
    u = synthetic.uniform_draws(sample_timesteps + 1, sample_jets, cfg.max_num_particles, cfg.vocab_size).to(device)
    sdd = {k: v.to(device) for k, v in sd.items()}

    def run(uu, steps):
        if cfg.model == "EPiC":
            return orc.simulate_dynamics_cfm(sdd, cfg, src.continuous, src.mask, max_steps=steps)
        return orc.simulate_dynamics(sdd, cfg, src.continuous, src.discrete, src.mask, u=uu, max_steps=steps)

    def sync():
        if device != "cpu":
            torch.cuda.synchronize()

    run(u, 1)                                                                                         # warm-up
    sync()
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        run(u[1:], sample_timesteps)
        sync()
        dt = (time.perf_counter() - t0) / sample_timesteps
        best = dt if best is None else min(best, dt)
    return sample_jets / (best * cfg.num_timesteps), best


class ReferenceSampler:
    """The UNMODIFIED reference (oracle/_ref or /root/reference behind the stub modules of oracle/ref_loader.py): its own
    MultiModalFlowBridge / ConditionalFlowMatching.simulate_dynamics on the whole batch, fp32, all host threads."""

    def __init__(self, args, cfg, sd):
        from oracle import ref_loader
        self.ok = ref_loader.available()
        self.where = ref_loader.REF_ROOT
        if not self.ok:
            return
        from mmf_b200 import synthetic
        torch.set_num_threads(os.cpu_count() or 1)
        self.ref = ref_loader.modules()
        self.args, self.cfg = args, cfg
        self.epic = cfg.model == "EPiC"
        self.sd = sd
        self.src = synthetic.source_state(args.batch, cfg.max_num_particles, cfg.vocab_size, dense=args.dense)
        self._models = {}

    def _model(self, timesteps):
        if timesteps not in self._models:
            import copy
            c = copy.copy(self.cfg)
            c.num_timesteps = timesteps
            m = (self.ref.ConditionalFlowMatching if self.epic else self.ref.MultiModalFlowBridge)(c).eval()
            m.model.load_state_dict(self.sd, strict=True)
            self._models[timesteps] = m
        return self._models[timesteps]

    def step(self, timesteps):
        """One bounded sample: the reference's simulate_dynamics over `timesteps` grid points of the whole batch. Seconds."""
        m = self._model(timesteps)
        ref = self.ref
        src = ref.TensorMultiModal(continuous=self.src.continuous.clone(), discrete=None if self.epic else self.src.discrete.clone(),
                                   mask=self.src.mask.clone())
        batch = ref.DataCoupling(source=src, target=ref.TensorMultiModal())
        t0 = time.perf_counter()
        with torch.no_grad():
            m.simulate_dynamics(batch)
        return time.perf_counter() - t0


def cpu_reference_measure(args, cfg, sd, steps, warmup):
    """(value jets/s @ cfg.num_timesteps, seconds per bench step, kind, sample text).  Every step = whole batch x S timesteps."""
    S = max(2, min(args.ref_sample_timesteps, cfg.num_timesteps))      # (S = 1 divides by zero in the reference's own dt, MMF.py:183-185)
    cores = os.cpu_count() or 1
    rs = ReferenceSampler(args, cfg, sd)
    if rs.ok:
        for _ in range(warmup):
            rs.step(S)
        secs = [rs.step(S) for _ in range(max(steps, 1))]
        kind = "reference"
        what = f"the unmodified reference ({'oracle/_ref' if rs.where.endswith('_ref') else rs.where}) simulate_dynamics"
    else:
        for _ in range(warmup):
            cpu_port_rate(args, cfg, sd, args.batch, 1)
        secs = [cpu_port_rate(args, cfg, sd, args.batch, S)[1] * S for _ in range(max(steps, 1))]
        kind = "port"
        what = "oracle/mmf_oracle.py (port; oracle/_ref absent)"
    per_step = sum(secs) / len(secs)
    value = args.batch / (per_step / S * cfg.num_timesteps)
    sample = (f"{what}, torch fp32, {cores} threads: the WHOLE batch of {args.batch} jets x {S} timesteps per step "
              f"({per_step:.2f} s measured), scaled by {cfg.num_timesteps}/{S} timesteps to the metric's {cfg.num_timesteps} steps; no extrapolation over jets")
    return value, per_step, kind, sample


def run_reference_arm(args, rank):
    if rank != 0:
        return
    from mmf_b200 import synthetic
    from mmf_b200.param_spec import make_config
    cfg = make_config(args.model, num_timesteps=args.timesteps, temperature=args.temperature)
    sd = synthetic.make_state_dict(cfg, flavor="wide", seed=0)
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    value, per_step, kind, sample = cpu_reference_measure(args, cfg, sd, args.steps, max(args.warmup, 0))
    wall = time.perf_counter() - t0
    emit({
        "impl": "reference", "metric": metric_name(args), "value": value, "unit": "jets/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "wall_s": wall,
                   "step": f"one step = the whole batch for {max(2, min(args.ref_sample_timesteps, args.timesteps))} of the {args.timesteps} timesteps (bounded sample)"},
        "cpu_baseline": {"value": value, "unit": "jets/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "jets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def timed_generate(nm, src_dev, ts, dt, cfg, steps, warmup, flush, epic):
    """ms per call of the device-resident sampler (CUDA events on the launching stream, L2 flushed between calls)."""
    from mmf_b200 import _abi

    def call(i):
        if epic:
            return nm.generate(src_dev.continuous, None, src_dev.mask, ts, dt, None)
        return nm.generate(src_dev.continuous, src_dev.discrete, src_dev.mask, ts, dt, _abi.step_options(cfg, seed=7, first_global_jet=i * len(src_dev)))
    for i in range(warmup):
        call(i)
    torch.cuda.synchronize()
    ms = 0.0
    for i in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call(warmup + i)
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    return ms / steps


def side_model(name, args, dev):
    from mmf_b200 import synthetic
    from mmf_b200.mmf import ConditionalFlowMatching, MultiModalFlowBridge, time_grid
    from mmf_b200.param_spec import make_config
    cfg = make_config(name, num_timesteps=args.timesteps, temperature=args.temperature)
    bridge = (ConditionalFlowMatching if name == "EPiC" else MultiModalFlowBridge)(cfg)
    bridge.model.load_state_dict(synthetic.make_state_dict(cfg, flavor="wide", seed=0), strict=True)
    bridge = bridge.to(dev)
    ts, dt = time_grid(cfg)
    return cfg, bridge, bridge.model.native(), ts, dt


def extra_model_lines(args, peaks, dev, flush, rank):
    """FusedParticleFormer and EPiC on one GPU, same batch shape and timing rules as the headline (a few steps each)."""
    from mmf_b200 import synthetic
    out = {}
    for name in ("FusedParticleFormer", "EPiC"):
        cfg, bridge, nm, ts, dt = side_model(name, args, dev)
        src = synthetic.source_state(args.batch, cfg.max_num_particles, cfg.vocab_size, seed=1234 + 10 * rank)
        n = src.mask.squeeze(-1).sum(1)
        ms = timed_generate(nm, src.to(dev), ts, dt, cfg, 3, 3, flush, name == "EPiC")
        tf = algorithmic_flops_per_timestep(name, n) * args.timesteps / (ms * 1e-3) / 1e12
        out[name] = {"value": args.batch / (ms * 1e-3), "unit": "jets/s", "ms_per_step": ms, "steps": 3, "warmup": 3,
                     "roofline": {"bound": "tensor", "kernel": "epic_tile_kernel" if name == "EPiC" else "tf_tile_kernel", "achieved": tf,
                                  "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": tf / peaks["bf16_tflops_sustained"]}}
        del bridge, nm
    return out


def whole_run_lines(args, dev, rank, world):
    """BASELINE configs #3 / #4 as a whole run on every rank, everything inside the clock: source built on the device
    (mmf_make_source) -> sampler (mmf_generate_n, asynchronous) -> de-standardise + mask + narrow (mmf_pack_sample) ->
    ONE all_gather_into_tensor of the per-jet records (13 B/slot).  jets/s = all jets of all ranks / max-over-ranks time."""
    from mmf_b200 import _abi, distributed as mdist
    out = {}
    for name, jets, batch in (("FusedParticleFormer", args.whole_run_jets, 1024), ("EPiC", 4 * args.whole_run_jets, 4096)):
        cfg, bridge, nm, ts, dt = side_model(name, args, dev)
        epic = name == "EPiC"
        D, V = cfg.max_num_particles, cfg.vocab_size
        probs = torch.exp(-0.5 * ((torch.arange(D + 1, dtype=torch.float64) - 55.0) / 18.0) ** 2)
        probs[0] = 0.0
        probs = probs.float().tolist()
        total = jets * world

        def run():
            lo, hi = mdist.shard_bounds(total, rank, world)
            recs = []
            for b0 in range(lo, hi, batch):
                b1 = min(b0 + batch, hi)
                x0, k0, mask, n = _abi.make_source(probs, b1 - b0, D, V, 11, b0, dev, discrete=not epic)
                x, k, _ = nm.generate(x0, k0, None, ts, dt, None if epic else _abi.step_options(cfg, seed=11, first_global_jet=b0),
                                      n_per_jet=n.cpu())
                recs.append(_abi.pack_sample(x, k, mask, [1.9, 0.0, 0.0], [0.8, 0.11, 0.1]))
            rec = torch.cat(recs, 0)
            counts = [mdist.shard_bounds(total, r, world)[1] - mdist.shard_bounds(total, r, world)[0] for r in range(world)]
            return mdist.gather_records(rec, counts)

        run()                                                  # warm-up (workspaces, NCCL communicator)
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        rec = run()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        tmax = torch.tensor([e0.elapsed_time(e1) * 1e-3, wall], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
        nm.status()
        assert rec.shape[0] == total
        out[name] = {"jets": total, "jets_per_rank": jets, "batch": batch, "timesteps": args.timesteps, "temperature": args.temperature,
                     "seconds_device": float(tmax[0]), "seconds_wall": float(tmax[1]), "value": total / float(tmax[0]), "unit": "jets/s",
                     "collective": "one all_gather_into_tensor of (jets, %d) uint8 records inside the timed region" % rec.shape[1],
                     "gathered_bytes": int(rec.numel())}
        del bridge, nm, rec
    return out


def training_flops(model: str, n: torch.Tensor) -> float:
    """Algorithmic FLOPs of one training step on real particles: forward + backward = 3 x forward
    (every product appears once more for the data gradient and once for the weight gradient)."""
    return 3.0 * algorithmic_flops_per_timestep(model, n)


def training_lines(args, peaks, dev, rank, world):
    """BASELINE config #5 on every rank: one step = bridge sampling + bf16 forward + backward + (N > 1: ONE all-reduce of the flat
    gradient) + norm clipping + Adam on a batch of `--batch` AOJ-shaped jets, all inside the clock.  jets/s = jets of all ranks /
    max-over-ranks CUDA-event time.  Rank 0 of a single-GPU run adds the unmodified reference's own training step on the host
    cores and fp32 eager torch (autograd over the oracle port + torch.optim.Adam) on the same GPU."""
    from mmf_b200 import synthetic
    from mmf_b200.mmf import MultiModalFlowBridge
    from mmf_b200.param_spec import make_config
    out = {}
    for name, steps in (("ParticleFormer", 8), ("FusedParticleFormer", 8)):
        cfg = make_config(name, lr=1e-3, lr_final=1e-5, max_epochs=100, warmup_epochs=5)
        sd = synthetic.make_state_dict(cfg, flavor="wide", seed=0)
        bridge = MultiModalFlowBridge(cfg)
        bridge.model.load_state_dict(sd, strict=True)
        bridge = bridge.to(dev)
        eng = bridge.configure_training(lr=cfg.lr)
        batch = synthetic.training_batch(args.batch, cfg.max_num_particles, cfg.vocab_size, seed=1234 + 10 * rank)
        batch.source, batch.target = batch.source.pin_memory(), batch.target.pin_memory()     # as a DataLoader hands it over
        h2d = sum(t.numel() * t.element_size() for tm in (batch.source, batch.target) for t in (tm.continuous, tm.discrete))
        for _ in range(3):
            eng.train_step(batch)
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        l0 = eng.ops.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            out5 = eng.train_step(batch)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        eng.check_tokens()
        tmax = torch.tensor([e0.elapsed_time(e1) * 1e-3, wall], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
        sec = float(tmax[0]) / steps
        fl = training_flops(name, torch.as_tensor(eng.last_plan.n))
        tf = fl / sec / 1e12
        out[name] = {"workload": f"{name} training step: MultiModalFlowBridge.loss ({cfg.multitask_loss}) forward + backward, gradient clip 1.0, Adam; "
                                 f"{args.batch} AOJ-shaped jets per GPU, bf16 operands / fp32 accumulation, master weights and gradients fp32",
                     "value": world * args.batch / sec, "unit": "jets/s", "ms_per_step": 1e3 * sec, "steps": steps, "warmup": 3,
                     "wall_ms_per_step": 1e3 * float(tmax[1]) / steps, "gpu_launches_per_step": (eng.ops.launches - l0) / steps,
                     "parameters": int(sum(int(np.prod(eng.shape[n])) for n in eng.names)), "loss_after": float(out5[0]),
                     "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 0,
                     "api": "MultiModalFlowBridge.configure_training() / TrainEngine.train_step(batch) on a pinned HOST batch (copies inside the clock)",
                     "collective": None if world == 1 else f"one all_reduce of the flat fp32 gradient ({4 * eng.total} bytes) per step, inside the timed region",
                     "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                                  "frac": tf / peaks["bf16_tflops_sustained"],
                                  "note": "algorithmic FLOPs = 3 x the forward count of SURVEY 8(d) on real particles, whole step (all kernels)"}}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            out[name].update(training_baselines(args, cfg, sd, dev, name))
        del eng, bridge
    return out


def training_baselines(args, cfg, sd, dev, name):
    """The same training step (same batch, same Adam + clip) by (a) the UNMODIFIED reference on the host cores and (b) fp32 eager
    torch on this GPU (autograd over oracle/mmf_oracle.py, TF32 off).  One measured step each after one warm-up step."""
    from mmf_b200 import synthetic
    from oracle import mmf_oracle as orc, ref_loader
    res = {}
    batch = synthetic.training_batch(args.batch, cfg.max_num_particles, cfg.vocab_size, seed=1234)
    if ref_loader.available():
        torch.set_num_threads(os.cpu_count() or 1)
        ref = ref_loader.modules()
        m = ref.MultiModalFlowBridge(cfg)
        m.model.load_state_dict(sd, strict=True)
        opt = m.configure_optimizers()["optimizer"]
        mk = lambda t: ref.TensorMultiModal(continuous=t.continuous.clone(), discrete=t.discrete.clone(), mask=t.mask.clone())
        secs = []
        for _ in range(2):
            b = ref.DataCoupling(source=mk(batch.source), target=mk(batch.target))
            t0 = time.perf_counter()
            opt.zero_grad()
            loss = m.loss(b)[0]
            loss.backward()
            torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
            opt.step()
            secs.append(time.perf_counter() - t0)
        res["cpu_baseline"] = {"value": args.batch / secs[-1], "unit": "jets/s", "cores": os.cpu_count() or 1, "kind": "reference",
                               "sample": f"the unmodified reference (oracle/_ref): MultiModalFlowBridge.loss + backward + clip_grad_norm_(1.0) + its own "
                                         f"configure_optimizers() Adam step, torch fp32, one step on the whole batch of {args.batch} jets ({secs[-1]:.2f} s)"}
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sdg = {k: torch.nn.Parameter(v.to(dev).clone()) for k, v in sd.items()}
    E = cfg.n_embd
    g = torch.Generator().manual_seed(3)
    slg = {"uncertainty_net.c_fc.weight": torch.randn(E, E, generator=g) * 0.02, "uncertainty_net.c_fc.bias": torch.zeros(E),
           "uncertainty_net.c_proj.weight": torch.randn(2, E, generator=g) * 0.02, "uncertainty_net.c_proj.bias": torch.zeros(2)}
    slg = {k: torch.nn.Parameter(v.to(dev)) for k, v in slg.items()}
    params = list(sdg.values()) + list(slg.values())
    opt = torch.optim.Adam(params, lr=cfg.lr)
    src, tgt = batch.source.to(dev), batch.target.to(dev)
    B, D = src.continuous.shape[:2]
    ms = []
    for _ in range(3):
        t = cfg.time_eps + (1.0 - cfg.time_eps) * torch.rand(B, device=dev)
        z, u = torch.randn(B, D, 3, device=dev), torch.rand(B, D, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        opt.zero_grad()
        loss = orc.training_loss(sdg, slg, cfg, src.continuous, src.discrete, tgt.continuous, tgt.discrete, src.mask, t, z, u)[0]
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    res["gpu_eager_baseline"] = {"value": args.batch / (min(ms[1:]) * 1e-3), "unit": "jets/s", "ms_per_step": min(ms[1:]),
                                 "kind": "port on cuda (autograd over oracle/mmf_oracle.py + torch.optim.Adam + clip_grad_norm_, eager torch fp32, TF32 off)"}
    return res


def workload_name(args):
    return (f"{args.model} sampler, batch {args.batch} jets/GPU x {args.timesteps} timesteps, D=150, V=9, "
            + ("dense n=150" if args.dense else "AOJ-shaped n~clamp(round(55+18z),1,150)") + f", T={args.temperature}")


def step_kernel_roofline(peaks, dev):
    """Standalone fused step kernel on a batch larger than L2: 65536 jets x 150 slots."""
    from mmf_b200 import _abi
    B, D, V = 65536, 150, 9
    g = torch.Generator(device=dev).manual_seed(3)
    vt = torch.randn(B, D, 3, device=dev, generator=g)
    logits = torch.randn(B, D, V, device=dev, generator=g)
    x = torch.randn(B, D, 3, device=dev, generator=g)
    k = torch.randint(0, V, (B, D), device=dev, generator=g)
    t = torch.full((B,), 0.5, device=dev)
    opts = _abi.MmfStepOptions(1.0, 0.075, 0, 0.0, 0, 1, 0)
    out = {}
    for mode, u in (("philox", None), ("supplied_u", torch.rand(B, D, V, device=dev, generator=g))):
        for _ in range(3):
            _abi.hybrid_step(vt, logits, x, k, t, 0.0101, opts, u=u, want_rates=False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for i in range(reps):
            _abi.hybrid_step(vt, logits, x, k, t, 0.0101, opts, u=u, step_index=i, want_rates=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        bytes_per_particle = 12 + 36 + 12 + 8 + 12 + 8 + (36 if u is not None else 0)
        gbs = B * D * bytes_per_particle / (ms * 1e-3) / 1e9
        out[mode] = {"bound": "hbm", "kernel": "hybrid_step_prod_kernel" if u is None else "hybrid_step_kernel",
                     "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                     "traffic": None, "bytes_per_particle": bytes_per_particle, "ms_per_launch": ms, "particles": B * D}
        try:                                  # DRAM bytes per launch of the same shape from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                tr = json.load(f).get(out[mode]["kernel"])
            if tr:
                out[mode]["traffic"] = tr["dram_bytes_per_launch"]
                out[mode]["traffic_note"] = tr["note"]
        except (OSError, ValueError, KeyError):
            pass
        del u
    del vt, logits, x, k
    out["jet_observables"] = observables_roofline(peaks, dev)
    return out


def observables_roofline(peaks, dev):
    """mmf_jet_observables on 2^19 AOJ-shaped jets (2.2 GB of input, far larger than L2).
    Algorithmic bytes: 8 (mask) per slot + 20 (x 12 + k 8) per real particle in, 48 + 4 V per jet out."""
    from mmf_b200 import _abi
    B, D, V = 1 << 19, 150, 9
    g = torch.Generator(device=dev).manual_seed(2)
    n = torch.clamp(torch.round(55 + 18 * torch.randn(B, device=dev, generator=g)), 1, D).long()
    mask = (torch.arange(D, device=dev)[None, :] < n[:, None]).long()
    x = torch.randn(B, D, 3, device=dev, generator=g)
    k = torch.randint(1, V, (B, D), device=dev, generator=g)
    for _ in range(3):
        _abi.jet_observables(x, k, mask, [1.9, 0.0, 0.0], [0.8, 0.11, 0.1])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        _abi.jet_observables(x, k, mask, [1.9, 0.0, 0.0], [0.8, 0.11, 0.1])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = 8 * B * D + 20 * int(n.sum()) + (48 + 4 * V) * B
    gbs = nbytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "jet_observables_kernel", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": gbs / peaks["hbm_gbs"], "traffic": None, "algorithmic_bytes": nbytes, "ms_per_launch": ms, "jets": B,
            "jets_per_s": B / (ms * 1e-3)}


_JSON_OUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the process's original stdout; everything else that writes to file descriptor 1 (NCCL prints
    its version banner there from C code, whatever NCCL_DEBUG_FILE says) has been pointed at stderr by main()."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _JSON_OUT
    args = parse_args()
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.cpu_sample_jets <= 0:
        args.cpu_sample_jets = args.batch
    args.cpu_sample_timesteps = max(1, min(args.cpu_sample_timesteps, args.timesteps))
    args.ref_sample_timesteps = max(1, min(args.ref_sample_timesteps, args.timesteps))
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the accelerated path has no CPU fallback")
    from mmf_b200 import _abi, synthetic
    from mmf_b200.mmf import ConditionalFlowMatching, MultiModalFlowBridge, time_grid
    from mmf_b200.param_spec import make_config
    from mmf_b200.tensorclass import DataCoupling, TensorMultiModal

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # with NCCL_DEBUG set (VERSION and up) NCCL writes "NCCL version ..." to STDOUT, next to the one JSON line: keep
        # whatever level the operator asked for, but send it to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        torch.distributed.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()

    cfg = make_config(args.model, num_timesteps=args.timesteps, temperature=args.temperature, batch_size=args.batch)
    sd = synthetic.make_state_dict(cfg, flavor="wide", seed=0)
    epic = args.model == "EPiC"
    bridge = (ConditionalFlowMatching if epic else MultiModalFlowBridge)(cfg)
    bridge.model.load_state_dict(sd, strict=True)
    bridge = bridge.to(dev)
    nm = bridge.model.native()
    ts, dt = time_grid(cfg)

    # every rank generates its own shard of jets; draws are keyed on the global jet index (world-size invariant)
    B = args.batch
    src_host = synthetic.source_state(B, cfg.max_num_particles, cfg.vocab_size, dense=args.dense, seed=1234 + 10 * rank).pin_memory()
    src_dev = src_host.to(dev)
    n_per_jet = src_host.mask.squeeze(-1).sum(1)
    flops_step = algorithmic_flops_per_timestep(args.model, n_per_jet) * args.timesteps
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def one_step(i):
        if epic:
            return nm.generate(src_dev.continuous, None, src_dev.mask, ts, dt, None)
        opts = _abi.step_options(cfg, seed=7, first_global_jet=(i * world + rank) * B)
        return nm.generate(src_dev.continuous, src_dev.discrete, src_dev.mask, ts, dt, opts)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        one_step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = nm.launches
    evs = []
    barrier()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()                                                            # L2 flush between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        one_step(args.warmup + i)
        e1.record()
        evs.append((e0, e1))
    barrier()
    wall = time.perf_counter() - wall0
    launches = nm.launches - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    tmax = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
    total_ms = float(tmax.item())
    value = world * B * args.steps / (total_ms * 1e-3)

    # ---- end-to-end through the drop-in API with pinned HOST buffers (H2D + D2H inside the timed region) ----
    def e2e_step(i):
        batch = DataCoupling(source=src_host, target=TensorMultiModal())
        out = bridge.predict_step(batch, i)              # (batch_idx -> global jet offset, as under Lightning's predict loop)
        return out
    for i in range(2):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        out = e2e_step(2 + i)
    torch.cuda.synchronize()
    e2e_t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(e2e_t, op=torch.distributed.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(e2e_t.item())
    slots = B * cfg.max_num_particles
    h2d = slots * (3 * 4 + (0 if epic else 8))         # x0 f32 + k0 i64 (the mask stays on the host: it only feeds the planner)
    d2h = slots * (3 * 4 + (0 if epic else 8))
    clocks = sampler.stop() if rank == 0 else None

    # ---- whole runs with the single end-of-run collective INSIDE the clock (all ranks) -----------------------
    whole_run = None if args.no_extras else whole_run_lines(args, dev, rank, world)
    training = None if (args.no_extras or epic) else training_lines(args, peaks, dev, rank, world)

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    # ---- live per-kernel-class profile of one more step (CUDA events on the launching stream) -------------
    prof = {}
    if not epic:
        nm.profile(True)
        nm.profile_read(reset=True)
        one_step(10_000)
        prof = nm.profile_read(reset=True)
        nm.profile(False)
    tc = {k: v for k, v in prof.items() if k in TC_CLASSES and v["launches"]}
    if not tc:
        # one persistent kernel runs the whole sampler (all timesteps in one launch): the step time IS that kernel's time
        ach = flops_step / (total_ms / args.steps * 1e-3) / 1e12
        name = "epic_tile_kernel" if epic else "tf_tile_kernel"
        roofline = {"bound": "tensor", "kernel": f"{name} (persistent: one CTA per 128-row tile of whole jets, all timesteps in one launch)",
                    "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops_sustained"],
                    "traffic": None, "peak_source": peaks["source"] + " sustained", "avg_launch_us": 1e3 * total_ms / args.steps,
                    "launches_per_step": launches / max(args.steps, 1), "share_of_step": 1.0}
    else:
        tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
        shares = {k: round(v["ms"] / tot_ms, 4) for k, v in prof.items() if v["launches"]}
        dom = max(tc, key=lambda k: tc[k]["ms"])
        d = tc[dom]
        ach = d["flops"] / d["launches"] / (d["ms"] / d["launches"] * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": ach / peaks["bf16_tflops_sustained"], "traffic": None, "peak_source": peaks["source"] + " sustained",
                    "avg_launch_us": 1e3 * d["ms"] / d["launches"], "launches_per_step": d["launches"],
                    "share_of_step": shares[dom], "kernel_time_shares": shares}
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (bytes per launch; null when absent)
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tr = json.load(f).get(roofline["kernel"].split(" ")[0])
        if tr:
            roofline["traffic"] = tr["dram_bytes_per_launch"]
            roofline["traffic_note"] = tr["note"]
    except (OSError, ValueError, KeyError):
        pass
    path_tflops = flops_step / (total_ms / args.steps * 1e-3) / 1e12
    roofline["whole_path"] = {"algorithmic_tflops": path_tflops, "frac_of_sustained_peak": path_tflops / peaks["bf16_tflops_sustained"]}

    line = {
        "metric": metric_name(args), "value": value, "unit": "jets/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(args), "jets_per_gpu_per_step": B, "timesteps": args.timesteps,
                   "real_particles_per_step": int(n_per_jet.sum()), "weights": "synthetic wide init seed 0 (random, no checkpoint offline)",
                   "l2": "256 MiB buffer written between timed iterations (L2 flush)", "rng": "in-kernel Philox4x32-10",
                   "wall_s_timed_region": wall},
        "e2e": {"value": e2e_value, "unit": "jets/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": ("mmf_b200.mmf.ConditionalFlowMatching.predict_step (pinned host batch -> mmf_generate -> .cpu())" if epic else
                        "mmf_b200.mmf.MultiModalFlowBridge.predict_step (pinned host batch -> mmf_generate_host)")},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
    }
    if whole_run is not None:
        line["whole_run"] = whole_run
    if training is not None:
        line["training"] = training
    # The legs below only add context to the line (rank 0, no collective).  One of them failing must not cost the headline
    # measured above: the failure is recorded in its place ({"error": ...}) and the line is still printed.
    def guarded(key, fn):
        try:
            fn()
        except Exception as exc:                       # noqa: BLE001 - reported, not hidden
            line[key] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
            sys.stderr.write(f"bench.py: leg {key} failed: {exc}\n")

    def dense_and_full_chip():
        # the dense worst case (every jet 150 particles -> CTA-pair tiles) on the same kernel, same timing rules
        dsrc = synthetic.source_state(B, cfg.max_num_particles, cfg.vocab_size, dense=True, seed=1234)
        dms = timed_generate(nm, dsrc.to(dev), ts, dt, cfg, 2, 2, flush, False)
        dn = dsrc.mask.squeeze(-1).sum(1)
        dtf = algorithmic_flops_per_timestep(args.model, dn) * args.timesteps / (dms * 1e-3) / 1e12
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                dtraffic = json.load(f).get("tf_tile_kernel_pair", {}).get("dram_bytes_per_launch")
        except (OSError, ValueError):
            dtraffic = None
        line["roofline_dense"] = {"bound": "tensor", "kernel": "tf_tile_kernel (pair tiles: one 150-particle jet per 2-CTA cluster)", "traffic": dtraffic,
                                  "workload": f"{args.model}, {B} jets of 150 particles x {args.timesteps} timesteps", "value": B / (dms * 1e-3),
                                  "unit_value": "jets/s", "ms_per_step": dms, "achieved": dtf, "peak": peaks["bf16_tflops_sustained"],
                                  "unit": "TFLOP/s", "frac": dtf / peaks["bf16_tflops_sustained"], "steps": 2, "warmup": 2}
        # the same kernel with every SM busy (4096 jets = 1760 tiles = 11.9 waves of 148): what the design sustains per SM when
        # the batch does not leave a quarter of the chip without a tile (256 jets = 110 tiles)
        fsrc = synthetic.source_state(4096, cfg.max_num_particles, cfg.vocab_size, seed=1234)
        fms = timed_generate(nm, fsrc.to(dev), ts, dt, cfg, 2, 1, flush, False)
        ftf = algorithmic_flops_per_timestep(args.model, fsrc.mask.squeeze(-1).sum(1)) * args.timesteps / (fms * 1e-3) / 1e12
        line["roofline_full_chip"] = {"bound": "tensor", "kernel": "tf_tile_kernel", "workload": f"{args.model}, 4096 AOJ-shaped jets x {args.timesteps} timesteps",
                                      "value": 4096 / (fms * 1e-3), "unit_value": "jets/s", "ms_per_step": fms, "achieved": ftf,
                                      "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": ftf / peaks["bf16_tflops_sustained"],
                                      "steps": 2, "warmup": 1}
        del fsrc

    def extra_models():
        line["extra_models"] = extra_model_lines(args, peaks, dev, flush, rank)

    def step_roofline():
        line["roofline_step_kernel"] = step_kernel_roofline(peaks, dev)

    def cpu_baseline():
        cores = os.cpu_count() or 1
        value_cpu, per_step, kind, sample = cpu_reference_measure(args, cfg, sd, 1, 1)
        line["cpu_baseline"] = {"value": value_cpu, "unit": "jets/s", "cores": cores, "kind": kind, "sample": sample}

    def gpu_eager():
        # the same fp32 algorithm as plain eager torch on THIS GPU (TF32 off, the torch default)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        S = max(1, min(args.eager_timesteps, args.timesteps))
        rate, s_per_ts = cpu_port_rate(args, cfg, sd, args.batch, S, device=str(dev))
        line["gpu_eager_baseline"] = {"value": rate, "unit": "jets/s", "kind": "port on cuda (oracle/mmf_oracle.py, eager torch fp32, TF32 off)",
                                      "sample": f"the whole batch of {args.batch} jets x {S} timesteps ({s_per_ts * 1e3:.1f} ms/timestep), scaled to {args.timesteps} timesteps"}

    if not args.no_extras:
        if not epic and not args.dense:
            guarded("roofline_dense", dense_and_full_chip)
        guarded("extra_models", extra_models)
    if not args.no_step_roofline and not epic:
        guarded("roofline_step_kernel", step_roofline)
    if world == 1 and not args.no_cpu_baseline:
        guarded("cpu_baseline", cpu_baseline)
        if not args.no_extras:
            guarded("gpu_eager_baseline", gpu_eager)
    emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
