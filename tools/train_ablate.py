"""Which operators carry the replay time of the training step?  Each run captures the graph with one operator class replaced by a
no-op (values are garbage, timing is what is measured) and reports the replay time.  usage: train_ablate.py [model] [jets]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-flows_b200")); sys.path.insert(0, ROOT)
import torch
from mmf_b200 import synthetic
from mmf_b200.mmf import MultiModalFlowBridge
from mmf_b200.param_spec import make_config
model = sys.argv[1] if len(sys.argv) > 1 else "ParticleFormer"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda:0")
GROUPS = {"none": [], "wgrad (gemm_tn)": ["gemm_tn"], "attention": ["attn_tc_fwd", "attn_tc_bwd"], "layernorm": ["ln_fwd", "ln_bwd"],
          "gemm (fwd + dgrad)": ["gemm", "gemm_qkv"], "qkln_bwd": ["qkln_bwd"], "cast / colsum": ["cast_transpose"],
          "heads + loss + embed": ["head_fwd", "head_bwd", "loss_fwd", "loss_bwd", "loss_combine", "embed_x_fwd", "embed_x_bwd", "embed_y_fwd", "embed_y_bwd", "sgemm", "jet_sum"]}
res = {}
for name, ops_off in GROUPS.items():
    cfg = make_config(model, lr=1e-3)
    bridge = MultiModalFlowBridge(cfg)
    bridge.model.load_state_dict(synthetic.make_state_dict(cfg, "wide", 0))
    bridge = bridge.to(dev)
    eng = bridge.configure_training(lr=1e-3)
    for o in ops_off:
        setattr(eng.ops, o, lambda *a, **k: None)
    batch = synthetic.training_batch(B)
    batch.source, batch.target = batch.source.pin_memory(), batch.target.pin_memory()
    for _ in range(2):
        eng.loss_and_grad(batch)
    slot = next(iter(eng._slots.values()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        slot.graph.replay()
    e1.record(); torch.cuda.synchronize()
    res[name] = round(e0.elapsed_time(e1) / 20, 3)
    del eng, bridge
    torch.cuda.empty_cache()
base = res["none"]
print(json.dumps({"model": model, "jets": B, "replay_ms": res, "saved_ms_when_removed": {k: round(base - v, 3) for k, v in res.items() if k != "none"}}))
