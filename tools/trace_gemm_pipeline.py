import sys,os; sys.path.insert(0,'.')
import torch, ctypes
import satk_path; satk=satk_path.load()
from importlib import import_module
O=import_module("self-attention-tacotron_b200.ops"); L=import_module("self-attention-tacotron_b200.lib")
dev="cuda"
for (M,N,K) in ((12800,1024,544),(4736,128,1024),(12800,256,1024)):
    A=torch.randn(M,K,device=dev); W=torch.randn(N,K,device=dev); C=torch.empty(M,N,device=dev)
    for _ in range(3): O.gemm(A,W,C,M,N,K,lda=K,ldb=K,ldc=N,transB=True,engine=2)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): O.gemm(A,W,C,M,N,K,lda=K,ldb=K,ldc=N,transB=True,engine=2)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    o=(ctypes.c_longlong*16)(); L.load().satk_debug_phase_cycles(3, o); o=list(o)
    print(M,N,K,"%.1f us"%(ms*1e3), "eff TF/s (1x) %.1f"%(2*M*N*K/ms/1e9))
    for i in range(4): print("  it",6+i,"P(tma issue) %d  F(full seen) %d  X(xform seen) %d  E(mma issued) %d"%tuple(o[i*4:i*4+4]))
