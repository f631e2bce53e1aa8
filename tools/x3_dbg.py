import sys, torch
sys.path.insert(0, '.')
from dxvae_b200 import _lib
lib = _lib.require_cuda()
st = torch.cuda.current_stream().cuda_stream
for (M, N, K) in [(192, 1024, 512), (192, 1024, 1024), (192, 27, 1024), (192, 55, 1024), (192, 2048, 512), (192, 512, 128),
                  (20000, 1024, 1024), (20000, 27, 1024), (192, 1536, 512), (192, 128, 512), (192, 100, 1024)]:
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).cuda(); W = torch.randn(N, K, generator=g).cuda(); b = torch.randn(N, generator=g).cuda()
    ref = A.double() @ W.double().t() + b.double()
    out = {}
    for var in (0, 16, 32):
        C = torch.full((M, N), float('nan'), device='cuda')
        _lib.check(lib.dxvae_test_gemm(var, M, N, K, A.data_ptr(), K, W.data_ptr(), K, C.data_ptr(), N, b.data_ptr(), 0, 0, st), 'g')
        out[var] = (C.double() - ref).abs().max().item() / ref.abs().max().item()
    print("M=%5d N=%4d K=%4d  rel err: fp32 %.2e  tf32 %.2e  3xtf32 %.2e" % (M, N, K, out[0], out[16], out[32]))
