"""tcgen05 GEMM: TF32 operands (fp32 in HBM) vs BF16 operands on the path's large forward shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dxvae_b200 import _lib
lib = _lib.require_cuda()
st = torch.cuda.current_stream().cuda_stream
for M, N, K in ((32768, 1536, 512), (32768, 2048, 512), (32768, 1024, 1024), (32768, 512, 2048)):
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); C = torch.empty(M, N, device="cuda")
    A16, W16 = A.to(torch.bfloat16), W.to(torch.bfloat16)
    for name, var, a, w in (("tf32", 16, A, W), ("bf16", 64, A16, W16)):
        for _ in range(3):
            _lib.check(lib.dxvae_test_gemm(var, M, N, K, a.data_ptr(), K, w.data_ptr(), K, C.data_ptr(), N, None, 0, 0, st), "g")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            _lib.check(lib.dxvae_test_gemm(var, M, N, K, a.data_ptr(), K, w.data_ptr(), K, C.data_ptr(), N, None, 0, 0, st), "g")
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("M=%d N=%d K=%d %s: %.3f ms  %.0f TFLOP/s" % (M, N, K, name, ms, 2.0 * M * N * K / ms / 1e9), flush=True)
