"""Per-stage clock trace of the transformer tile kernel (debugging aid): MMF_TRACE=<file> python tools/tf_trace.py [model]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-flows_b200")); sys.path.insert(0, ROOT)
import torch
from mmf_b200 import _abi, synthetic
from mmf_b200.param_spec import make_config
from mmf_b200.mmf import time_grid
model = sys.argv[1] if len(sys.argv) > 1 else "ParticleFormer"
cfg = make_config(model, num_timesteps=4)
sd = synthetic.make_state_dict(cfg, "wide", 0)
nm = _abi.NativeModel(cfg, sd, torch.device("cuda:0"))
dense = len(sys.argv) > 2 and sys.argv[2] == "dense"          # dense: pair tiles only (every jet 150 particles)
src = (synthetic.source_state(4, dense=True) if dense else synthetic.source_state(256)).to("cuda:0")
ts, dt = time_grid(cfg)
for _ in range(2):
    nm.generate(src.continuous, src.discrete, src.mask, ts, dt, _abi.step_options(cfg))
torch.cuda.synchronize()
text = open(os.environ["MMF_TRACE"]).read()
print(text)
# summary of timestep 1: cycles the epilogue warps spent waiting on each completion barrier (a stamp tagged "doneN arrived"
# follows a stamp tagged "before a wait", so its delta is a pure wait) against everything else (epilogue work)
import re
waits, work, nwait = {}, 0, {}
for line in text.splitlines():
    m = re.match(r"step 1 mark\s+\d+\s+\+(\d+) cycles \(total (\d+)\)(?: \[(.*)\])?", line)
    if not m:
        continue
    d, tag = int(m.group(1)), m.group(3)
    if tag and tag.endswith("arrived"):
        waits[tag] = waits.get(tag, 0) + d
        nwait[tag] = nwait.get(tag, 0) + 1
    else:
        work += d
total = work + sum(waits.values())
if total:
    print(f"summary of timestep 1 ({model}): {total} cycles; epilogue work {work} ({100 * work / total:.1f} %)")
    for tag in sorted(waits):
        print(f"  waiting, {tag}: {waits[tag]} cycles in {nwait[tag]} waits ({100 * waits[tag] / total:.1f} %)")
