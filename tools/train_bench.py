"""Timing of the device training step (BASELINE config #5): whole-step CUDA-event time and a per-operator breakdown.
usage: train_bench.py [model] [jets] [steps]   (MMF_TRAIN_PROFILE=1 brackets every operator with events: slower, for the table)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-flows_b200")); sys.path.insert(0, ROOT)
import torch
from mmf_b200 import synthetic
from mmf_b200.mmf import MultiModalFlowBridge
from mmf_b200.param_spec import make_config

model = sys.argv[1] if len(sys.argv) > 1 else "ParticleFormer"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device("cuda:0")
cfg = make_config(model, lr=1e-3)
bridge = MultiModalFlowBridge(cfg)
bridge.model.load_state_dict(synthetic.make_state_dict(cfg, "wide", 0))
bridge = bridge.to(dev)
eng = bridge.configure_training(lr=1e-3, use_graphs=not (os.environ.get('MMF_TRAIN_EAGER') or os.environ.get('MMF_TRAIN_PROFILE')))
batch = synthetic.training_batch(B)
if os.environ.get("MMF_TRAIN_DEVICE_BATCH"):
    batch.source, batch.target = batch.source.to(dev), batch.target.to(dev)
else:                                   # the DataLoader's view: a pinned host batch, copied to the device inside the step
    batch.source, batch.target = batch.source.pin_memory(), batch.target.pin_memory()

prof = {}
if os.environ.get("MMF_TRAIN_PROFILE"):
    ops = eng.ops
    for name in [n for n in dir(ops) if not n.startswith("_") and callable(getattr(ops, n))]:
        fn = getattr(ops, name)
        def wrap(fn=fn, name=name):
            def inner(*a, **k):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); r = fn(*a, **k); e1.record()
                prof.setdefault(name, []).append((e0, e1))
                return r
            return inner
        setattr(ops, name, wrap())

for _ in range(3):
    eng.train_step(batch)
torch.cuda.synchronize()
prof.clear()
l0 = eng.ops.launches
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(steps):
    out = eng.train_step(batch)
e1.record()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
n = eng.last_plan.n
fl = {"ParticleFormer": lambda n: 10630656 * n + 65536 + 11264 * n * n, "FusedParticleFormer": lambda n: 5649920 * n + 5120 * n * n}[model]
flops = 3.0 * float(sum(fl(int(v)) for v in n))
res = {"model": model, "jets": B, "rows": int(eng.last_plan.M), "ms_per_step": ms, "host_ms_per_step": 1e3 * t_host / steps, "jets_per_s": B / ms * 1e3,
       "launches_per_step": (eng.ops.launches - l0) / steps, "algorithmic_tflops": flops / ms / 1e9, "loss": float(out[0])}
if eng.use_graphs:
    slot = next(iter(eng._slots.values()))
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        slot.graph.replay()
    e1.record()
    torch.cuda.synchronize()
    res["graph_replay_ms"] = e0.elapsed_time(e1) / steps
    e0.record()
    for _ in range(steps):
        eng.optimizer_step()
    e1.record()
    torch.cuda.synchronize()
    res["optimizer_ms"] = e0.elapsed_time(e1) / steps
if prof:
    res["ops_ms_per_step"] = {k: round(sum(a.elapsed_time(b) for a, b in v) / steps, 4) for k, v in sorted(prof.items())}
    res["ops_calls_per_step"] = {k: len(v) / steps for k, v in sorted(prof.items())}
print(json.dumps(res))
