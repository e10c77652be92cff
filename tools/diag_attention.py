"""Diagnostic: run the attention kernel on one jet configuration (separate process per configuration)."""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "multimodal-flows_b200"))
import torch
from mmf_b200 import _abi

C, hs = int(sys.argv[1]), int(sys.argv[2])
jets = [int(a) for a in sys.argv[3:]]
dev = torch.device("cuda:0")
rows = sum(jets); M = (rows + 127) // 128 * 128
g = torch.Generator().manual_seed(1)
q = torch.randn(M, C, generator=g).bfloat16().to(dev); k = torch.randn(M, C, generator=g).bfloat16().to(dev)
v = torch.randn(M, C, generator=g).bfloat16().to(dev); vT = v.T.contiguous()
out = torch.zeros(M, C, device=dev, dtype=torch.bfloat16)
jn = (ctypes.c_int32 * len(jets))(*jets)
_abi.check(_abi.lib().mmf_dbg_attention(q.data_ptr(), k.data_ptr(), vT.data_ptr(), jn, len(jets), M, C, hs, out.data_ptr(), 0, None))
torch.cuda.synchronize()
H = C // hs; ref = torch.zeros(M, C, device=dev); s = 0; errs = []
for n in jets:
    qq = q[s:s+n].float().view(n, H, hs).transpose(0, 1); kk = k[s:s+n].float().view(n, H, hs).transpose(0, 1)
    vv = v[s:s+n].float().view(n, H, hs).transpose(0, 1)
    ref[s:s+n] = (torch.softmax(qq @ kk.transpose(1, 2) / hs ** 0.5, -1) @ vv).transpose(0, 1).reshape(n, C)
    errs.append(round(float((out[s:s+n].float() - ref[s:s+n]).norm() / ref[s:s+n].norm()), 4)); s += n
print("OK", C, hs, jets, errs)
