"""Probe of the MN-major tcgen05 contraction (agcn_gemm_tn): identity-like operands make a wrong shared-memory layout
visible as a permutation of the output.  Needs a GPU.

    python tools/dbg_tn.py
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import agcn_b200
from agcn_b200 import _lib
L = _lib.lib()
def run(M, Kd, N, S, A, D, tc):
    sb = L.agcn_gemm_tn_scratch_bytes(M, Kd, N, S)
    scr = torch.zeros(sb, dtype=torch.uint8, device='cuda')
    out = torch.full((Kd * S, N), -7.0, device='cuda')
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(L.agcn_gemm_tn(vp(A), vp(A), vp(D), vp(out), M, Kd, N, S, vp(scr), tc, None))
    torch.cuda.synchronize()
    return out
torch.manual_seed(0)
M, Kd, N = 64, 64, 64
A = torch.zeros(M, Kd, device='cuda'); D = torch.zeros(M, N, device='cuda')
A[torch.arange(M), torch.arange(M) % Kd] = 1.0          # A[r, r] = 1
D[:] = torch.arange(M, device='cuda')[:, None] * 100 + torch.arange(N, device='cuda')[None, :]
ref = A.t() @ D
o0 = run(M, Kd, N, 1, A, D, 0); o1 = run(M, Kd, N, 1, A, D, 1)
print("simt ok", torch.allclose(o0, ref))
print("tc  out[0:4,0:6]\n", o1[:4, :6]); print("ref out[0:4,0:6]\n", ref[:4, :6])
print("tc nonzero count", int((o1 != 0).sum()), "of", o1.numel(), "minus7 count", int((o1 == -7).sum()))
A2 = torch.randn(300, 128, device='cuda'); D2 = torch.randn(300, 128, device='cuda')
r2 = A2.double().t() @ D2.double()
o2 = run(300, 128, 128, 1, A2, D2, 1)
print("rand rel err", float((o2.double() - r2).abs().max() / r2.abs().max()))
