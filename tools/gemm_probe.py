"""FP32 FFMA GEMM microbenchmark through dxvae_test_gemm (variant 0 = forward form)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dxvae_b200 import _lib
lib = _lib.require_cuda()
st = torch.cuda.current_stream().cuda_stream
for M, N, K in ((32768, 1536, 512), (16384, 2048, 512), (16384, 1024, 1024)):
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
    C = torch.empty(M, N, device="cuda")
    for _ in range(3):
        _lib.check(lib.dxvae_test_gemm(0, M, N, K, A.data_ptr(), K, W.data_ptr(), K, C.data_ptr(), N, b.data_ptr(), 0, 0, st), "g")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        _lib.check(lib.dxvae_test_gemm(0, M, N, K, A.data_ptr(), K, W.data_ptr(), K, C.data_ptr(), N, b.data_ptr(), 0, 0, st), "g")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    ref = A[:64].double() @ W.double().t() + b.double()
    err = (C[:64].double() - ref).abs().max().item()
    print("M=%d N=%d K=%d: %.3f ms  %.1f TFLOP/s  err %.2e  map=%s" % (M, N, K, ms, 2.0 * M * N * K / ms / 1e9, err,
          "2x16" if os.environ.get("DX_GEMM_MAP_2x16") else "4x8"), flush=True)
# reference point: cuBLAS SGEMM (TF32 off) on the same shapes
torch.backends.cuda.matmul.allow_tf32 = False
for M, N, K in ((32768, 1536, 512), (16384, 1024, 1024)):
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda")
    for _ in range(3):
        C = A @ W.t()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        C = A @ W.t()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("cuBLAS sgemm M=%d N=%d K=%d: %.3f ms  %.1f TFLOP/s" % (M, N, K, ms, 2.0 * M * N * K / ms / 1e9), flush=True)
