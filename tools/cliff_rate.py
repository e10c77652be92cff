"""Cost of a few jets of more than 128 particles in a batch (they take the layered kernels): the same batch with 0 / 1 / 4 such
jets, with the tile kernel and the layered path overlapped on two streams (default) and serialised (MMF_NO_OVERLAP=1)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-flows_b200")); sys.path.insert(0, ROOT)
import torch
from mmf_b200 import _abi, synthetic
from mmf_b200.param_spec import make_config
from mmf_b200.mmf import time_grid
dev = torch.device("cuda:0")
for model in ("ParticleFormer", "FusedParticleFormer"):
    cfg = make_config(model, num_timesteps=100)
    nm = _abi.NativeModel(cfg, synthetic.make_state_dict(cfg, "wide", 0), dev)
    ts, dt = time_grid(cfg)
    for B in (256, 4096):
        for big in (0, 1, 4):
            src = synthetic.source_state(B).to(dev)
            for j in range(big):                                   # jets 0..big-1 get 140 particles
                src.mask[j, :140] = 1
                src.continuous[j, :140] = torch.randn(140, 3, device=dev)
                src.discrete[j, :140] = torch.randint(1, 9, (140, 1), device=dev)
            res = {}
            for mode in ("overlap", "serial"):
                if mode == "serial": os.environ["MMF_NO_OVERLAP"] = "1"
                else: os.environ.pop("MMF_NO_OVERLAP", None)
                for _ in range(2):
                    x, k, _ = nm.generate(src.continuous, src.discrete, src.mask, ts, dt, _abi.step_options(cfg, seed=3))
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(2):
                    x, k, _ = nm.generate(src.continuous, src.discrete, src.mask, ts, dt, _abi.step_options(cfg, seed=3))
                e1.record(); torch.cuda.synchronize()
                res[mode] = e0.elapsed_time(e1) / 2
                res[mode + "_x"] = x.clone()
            same = bool(torch.equal(res.pop("overlap_x"), res.pop("serial_x")))
            print(json.dumps({"model": model, "jets": B, "jets_over_128": big, "ms_overlap": round(res["overlap"], 2), "ms_serial": round(res["serial"], 2),
                              "identical_output": same}), flush=True)
    nm.close()
