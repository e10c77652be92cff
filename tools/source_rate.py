"""mmf_make_source on 2^20 jets: time per call (two launches) and HBM write fraction. Algorithmic bytes: 28 per slot + 4 per jet."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-flows_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from mmf_b200 import _abi
dev = torch.device("cuda:0")
peaks = bench.load_peaks()
B, D, V = 1 << 20, 150, 9
n = np.clip(np.round(55 + 18 * np.random.default_rng(0).standard_normal(100000)), 1, D).astype(int)
probs = (np.bincount(n, minlength=D + 1) / len(n)).astype(np.float32)
lib = _abi.lib()
import ctypes
x0 = torch.empty(B, D, 3, device=dev); k0 = torch.empty(B, D, device=dev, dtype=torch.int64)
mask = torch.empty(B, D, device=dev, dtype=torch.int64); nn = torch.empty(B, device=dev, dtype=torch.int32)
pc = (ctypes.c_float * (D + 1))(*[float(v) for v in probs])
def call(seed):
    _abi.check(lib.mmf_make_source(pc, B, D, V, seed, 0, _abi.ptr(x0), _abi.ptr(k0), _abi.ptr(mask), _abi.ptr(nn), 0, _abi.stream_handle(dev)))
for i in range(3):
    call(i)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for i in range(reps):
    call(10 + i)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
nbytes = 28 * B * D + 4 * B
gbs = nbytes / (ms * 1e-3) / 1e9
print(json.dumps({"kernel": "source_fill_kernel (+ source_mult_kernel)", "jets": B, "slots": B * D, "real_particles": int(nn.sum()), "ms_per_call": ms,
                  "jets_per_s": B / (ms * 1e-3), "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                  "algorithmic_bytes": nbytes}))
