"""3xTF32 (in-kernel hi/lo split) probe: accuracy against float64 and speed against plain TF32 for the three operand
forms on the path's shapes.  Run on the GPU box: python tools/x3_probe.py [quick]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dxvae_b200 import _lib  # noqa: E402

L = _lib.require_cuda()
st = lambda: torch.cuda.current_stream().cuda_stream


def run(variant, M, N, K, A, lda, B, ldb, C, ldc, bias=None, act=0, acc=0):
    _lib.check(L.dxvae_test_gemm(variant, M, N, K, A.data_ptr(), lda, B.data_ptr(), ldb, C.data_ptr(), ldc,
                                 None if bias is None else bias.data_ptr(), act, acc, st()), "gemm")


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def probe(M, N, K, timing=True):
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).cuda(); W = torch.randn(N, K, generator=g).cuda(); X = torch.randn(M, K, generator=g).cuda()
    dY = torch.randn(M, N, generator=g).cuda()
    out = []
    # forward
    C = torch.full((M, N), float("nan"), device="cuda")
    ref = A.double() @ W.double().t()
    for name, var in (("fp32", 0), ("tf32", 16), ("x3", 32)):
        run(var, M, N, K, A, K, W, K, C, N)
        err = (C.double() - ref).abs().max().item() / ref.abs().max().item()
        t = timeit(lambda: run(var, M, N, K, A, K, W, K, C, N)) if timing else 0
        out.append("fwd %-4s err %.2e %7.1f us %6.1f TF" % (name, err, t * 1e3, 2.0 * M * N * K / (t * 1e-3 + 1e-30) / 1e12))
    dX = torch.full((M, K), float("nan"), device="cuda")
    ref = dY.double() @ W.double()
    for name, var in (("fp32", 1), ("tf32", 17), ("x3", 33)):
        run(var, M, N, K, dY, N, W, K, dX, K)
        err = (dX.double() - ref).abs().max().item() / ref.abs().max().item()
        t = timeit(lambda: run(var, M, N, K, dY, N, W, K, dX, K)) if timing else 0
        out.append("dgrad %-4s err %.2e %7.1f us %6.1f TF" % (name, err, t * 1e3, 2.0 * M * N * K / (t * 1e-3 + 1e-30) / 1e12))
    ref = dY.double().t() @ X.double()
    for name, var in (("fp32", 2), ("tf32", 18), ("x3", 34)):
        dW = torch.zeros(N, K, device="cuda")
        run(var, M, N, K, dY, N, X, K, dW, K)
        err = (dW.double() - ref).abs().max().item() / ref.abs().max().item()
        t = timeit(lambda: run(var, M, N, K, dY, N, X, K, dW, K)) if timing else 0
        out.append("wgrad %-4s err %.2e %7.1f us %6.1f TF" % (name, err, t * 1e3, 2.0 * M * N * K / (t * 1e-3 + 1e-30) / 1e12))
    print("M=%d N=%d K=%d" % (M, N, K))
    for o in out:
        print("   ", o)
    sys.stdout.flush()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        # one designated product per operand form, twice each (the second launch is the one to capture under ncu):
        #   python tools/x3_probe.py one M N K [variant-base: 32 = 3xTF32 (default), 16 = TF32]
        M, N, K = (int(a) for a in sys.argv[2:5])
        vb = int(sys.argv[5]) if len(sys.argv) > 5 else 32
        g = torch.Generator().manual_seed(1)
        A = torch.randn(M, K, generator=g).cuda(); W = torch.randn(N, K, generator=g).cuda(); dY = torch.randn(M, N, generator=g).cuda()
        C = torch.empty(M, N, device="cuda"); dX = torch.empty(M, K, device="cuda"); dW = torch.zeros(N, K, device="cuda")
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        for _ in range(2):
            flush.zero_()                                   # the operands do not sit in L2 from the previous launch
            run(vb, M, N, K, A, K, W, K, C, N)
            flush.zero_()
            run(vb + 1, M, N, K, dY, N, W, K, dX, K)
            flush.zero_()
            run(vb + 2, M, N, K, dY, N, A, K, dW, K)
        torch.cuda.synchronize()
        print("one", M, N, K, vb)
        sys.exit(0)
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    print("DX_X3_INPLACE =", os.environ.get("DX_X3_INPLACE"))
    shapes = [(256, 256, 64), (1000, 1024, 1024), (300, 2048, 512)] if quick else \
        [(32768, 1536, 512), (32768, 2048, 512), (32768, 1024, 1024), (32768, 512, 512), (32768, 1536, 32), (8192, 1536, 512),
         (4096, 512, 512), (128, 1536, 512), (32768, 27, 1024), (32768, 1, 1024), (2000, 1536, 512), (4000, 512, 1536), (10000, 1536, 512), (10000, 2048, 512),
         (14000, 1536, 512), (20000, 1536, 512)]
    for s in shapes:
        probe(*s, timing=not quick)
