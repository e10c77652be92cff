"""DDP check of the device training step under torchrun (NCCL): after a few steps on DIFFERENT per-rank batches the replicas hold
bit-identical parameters, and they equal (to accumulation order) a single-process engine that averages the per-rank gradients itself.
usage: torchrun --nproc-per-node 2 tools/ddp_check.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-flows_b200")); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from mmf_b200 import synthetic
from mmf_b200.mmf import MultiModalFlowBridge
from mmf_b200.param_spec import make_config
from mmf_b200.training import TrainEngine
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=dev)
cfg = make_config("FusedParticleFormer", lr=1e-3, n_layer=3)
sd = synthetic.make_state_dict(cfg, "wide", 0)
def make():
    b = MultiModalFlowBridge(cfg)
    b.model.load_state_dict(sd)
    return b.to(dev)
eng = TrainEngine(make(), lr=1e-3, use_graphs=True)      # broadcasts rank 0's parameters (the loss net is randomly initialised per process)
P_start = eng.P.clone()
B, steps = 32, 4
gen = torch.Generator().manual_seed(5)
draws = [[(torch.rand(B, generator=gen), torch.randn(B, 150, 3, generator=gen), torch.rand(B, 150, generator=gen)) for _ in range(world)] for _ in range(steps)]
batches = [[synthetic.training_batch(B, seed=300 + 10 * s + r) for r in range(world)] for s in range(steps)]
for s in range(steps):
    t, z, u = draws[s][rank]
    eng.train_step(batches[s][rank], time=t, z=z, u=u)
torch.cuda.synchronize()
mine = eng.P.clone()
gathered = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(gathered, mine)
ok_replicas = all(torch.equal(gathered[0], g) for g in gathered)
# single-process emulation on rank 0: gradients of every rank's batch averaged by hand, world size 1 optimiser
msg = ""
if rank == 0:
    dist_backup = torch.distributed.is_initialized
    torch.distributed.is_initialized = lambda: False          # the solo engine must neither broadcast nor all-reduce
    solo = TrainEngine(make(), lr=1e-3, use_graphs=False)
    solo.P.copy_(P_start)
    solo.refresh_operands()
    try:
        for s in range(steps):
            acc = torch.zeros_like(solo.G)
            for r in range(world):
                t, z, u = draws[s][r]
                solo.loss_and_grad(batches[s][r], time=t, z=z, u=u)
                acc += solo.G
            solo.G.copy_(acc / world)
            solo.optimizer_step()
    finally:
        torch.distributed.is_initialized = dist_backup
    # Adam turns rounding-level differences of near-zero gradient entries into +-lr steps, so compare the UPDATES in norm
    upd = float(((mine - P_start) - (solo.P - P_start)).norm() / (solo.P - P_start).norm())
    msg = f"replicas identical: {ok_replicas}; relative difference of the {steps}-step update to the single-process run: {upd:.3e}"
    print(msg, flush=True)
    assert ok_replicas and upd < 0.1, msg
dist.barrier()
dist.destroy_process_group()
