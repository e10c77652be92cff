"""Split-K sweep of the weight-gradient GEMM (mmf_tr_gemm_tn) and timing of the forward / data-gradient shapes of one training step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-flows_b200"))
import torch
from mmf_b200._train_abi import Ops
dev = torch.device("cuda:0")
ops = Ops(dev)
K = 13819
def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("wgrad TN: (out, in) x tokens 13819; us per call by ksplit")
for (M, N) in [(384, 128), (128, 128), (512, 128), (128, 512), (768, 256), (256, 256), (512, 256), (256, 512)]:
    dy, x = torch.randn(K, M, device=dev).bfloat16(), torch.randn(K, N, device=dev).bfloat16()
    C = torch.zeros(M, N, device=dev)
    row = []
    for ks in (2, 4, 8, 12, 16, 24, 36, 54, 108, 216):
        row.append((ks, round(timeit(lambda: ops.gemm_tn(dy, x, C, ks)), 1)))
    print((M, N), row)
print("forward / dgrad NT: tokens x (N, K); us per call  mode0 (bf16) / mode1 (fp32)")
for (N, Kd) in [(384, 128), (128, 128), (512, 128), (128, 512), (768, 256), (256, 256), (512, 256), (256, 512), (128, 384), (256, 768)]:
    A, B = torch.randn(K, Kd, device=dev).bfloat16(), torch.randn(N, Kd, device=dev).bfloat16()
    C0, C1 = torch.zeros(K, N, device=dev, dtype=torch.bfloat16), torch.zeros(K, N, device=dev)
    print((N, Kd), round(timeit(lambda: ops.gemm(A, B, C0, None, 0)), 1), round(timeit(lambda: ops.gemm(A, B, C1, None, 1)), 1))
