"""Is the 3xTF32 error a truncation bias of the TMEM accumulation?  Signed error of C = A W^T for positive / negated /
mixed operands, and the same reduction issued as chunks of 128 with an FP32 (RN) reduce-add between the chunks."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dxvae_b200 import _lib
L = _lib.require_cuda()
st = lambda: torch.cuda.current_stream().cuda_stream

def dgrad(var, M, N, K, dY, ldy, W, dX, acc):
    _lib.check(L.dxvae_test_gemm(var, M, N, K, dY.data_ptr(), ldy, W.data_ptr(), K, dX.data_ptr(), K, None, 0, acc, st()), "g")

M, N, K = 4096, 1024, 512           # reduction over N
g = torch.Generator().manual_seed(1)
for name, mk in (("positive", lambda *s: torch.rand(*s, generator=g)), ("negA", None), ("mixed", lambda *s: torch.randn(*s, generator=g))):
    if name == "negA":
        dY = -dY
    else:
        dY = mk(M, N).cuda(); W = mk(N, K).cuda()
    ref = dY.double() @ W.double()
    for var, vn in ((1, "ffma"), (33, "x3")):
        dX = torch.zeros(M, K, device="cuda")
        dgrad(var, M, N, K, dY, N, W, dX, 0)
        e = (dX.double() - ref) / ref.abs().max()
        print("%-8s %-4s single  : max %.2e  mean signed %.2e  mean signed rel-to-elem %.2e" % (name, vn, e.abs().max().item(), e.mean().item(), ((dX.double() - ref) / ref).mean().item()))
        for ch in (128, 64):
            dX = torch.zeros(M, K, device="cuda")
            for c in range(0, N, ch):
                dgrad(var, M, ch, K, dY[:, c:], N, W[c:c + ch], dX, 1)
            e = (dX.double() - ref) / ref.abs().max()
            print("%-8s %-4s chunk%-3d: max %.2e  mean signed %.2e  mean signed rel-to-elem %.2e" % (name, vn, ch, e.abs().max().item(), e.mean().item(), ((dX.double() - ref) / ref).mean().item()))
