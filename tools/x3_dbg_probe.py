"""Times one large 3xTF32 forward product under the DX_X3_DBG experiment switches of k_tc_gemm_x3 (set in the environment
by the caller): which stage of the pipeline bounds it?"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dxvae_b200 import _lib
L = _lib.require_cuda()
M, N, K = 32768, 1536, 512
if len(sys.argv) > 3: M, N, K = map(int, sys.argv[1:4])
A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); C = torch.empty(M, N, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def run(): _lib.check(L.dxvae_test_gemm(32, M, N, K, A.data_ptr(), K, W.data_ptr(), K, C.data_ptr(), N, None, 0, 0, st), "g")
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 10
print("DBG=%-3s CHUNK=%-2s V1=%s  %d x %d x %d : %.1f us  %.1f TFLOP/s" % (os.environ.get("DX_X3_DBG", "0"), os.environ.get("DX_X3_CHUNK", "2"),
      os.environ.get("DX_X3_V1", "-"), M, N, K, t * 1e3, 2.0 * M * N * K / (t * 1e-3) / 1e12))
