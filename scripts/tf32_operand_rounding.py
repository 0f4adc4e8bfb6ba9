import sys; sys.path.insert(0,'/root/repo')
import torch, numpy as np
from d3p_b200 import _native as _n
dev=torch.device('cuda')
g=torch.Generator(device='cuda').manual_seed(1)
M,N,K=256,224,512
A=torch.randn((M,K),device=dev,generator=g); B=torch.randn((N,K),device=dev,generator=g)
def trunc(x): return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)
def rn(x):   # round to nearest even at 10 mantissa bits
    i=x.view(torch.int32); r=i + 0xFFF + ((i>>13)&1); return (r & ~0x1FFF).view(torch.float32)
want=A.double()@B.double().T
scale=(A.abs().max()*B.abs().max()*np.sqrt(K)).item()
def run(a_hi,a_lo,b_hi,b_lo):
    out=torch.empty((1,M,N),device=dev)
    _n.check(_n.lib().d3p_gemm_tf32x3(_n.ptr(a_hi),_n.ptr(a_lo),0,K,_n.ptr(b_hi),_n.ptr(b_lo),0,K,M,N,K,1,224,_n.ptr(out),N,M*N,0,_n.stream_ptr()))
    torch.cuda.synchronize(); return ((out[0].double()-want).abs().max().item()/scale)
print('masked hi, lo=x-trunc      ',run(trunc(A),A-trunc(A),trunc(B),B-trunc(B)))
print('RAW hi,    lo=x-trunc      ',run(A.clone(),A-trunc(A),B.clone(),B-trunc(B)))
print('RAW hi,    lo=x-rn         ',run(A.clone(),A-rn(A),B.clone(),B-rn(B)))
print('RAW hi only                ',run(A.clone(),None,B.clone(),None))
print('trunc hi only              ',run(trunc(A),None,trunc(B),None))
print('rn hi only                 ',run(rn(A),None,rn(B),None))
