import torch, ctypes, sys
sys.path.insert(0,'.')
from dxvae_b200 import _lib
lib=_lib.require_cuda()
st=torch.cuda.current_stream().cuda_stream
for (M,N,K) in [(8192,1536,512),(8192,1536,512),(8192,2048,512)]:
    A=torch.randn(M,K,device='cuda'); W=torch.randn(N,K,device='cuda'); C=torch.empty(M,N,device='cuda')
    _lib.check(lib.dxvae_test_gemm(16,M,N,K,A.data_ptr(),K,W.data_ptr(),K,C.data_ptr(),N,None,0,0,st),'g')
torch.cuda.synchronize()
