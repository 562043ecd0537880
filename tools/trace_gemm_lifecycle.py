import sys,os; sys.path.insert(0,'.')
import torch, ctypes
import satk_path; satk=satk_path.load()
from importlib import import_module
O=import_module("self-attention-tacotron_b200.ops"); L=import_module("self-attention-tacotron_b200.lib")
dev="cuda"
names=["setup","tma0 issued","tile0 landed","split0 done","mma0 start","last commit","accum done","epilogue done","teardown","ldtm0 done","staged0","rows0 done","rows1 done"]
for (M,N,K,kw) in ((4736,128,128,{}),(12800,256,256,{}),(12800,256,256,{"beta":1.0})):
    A=torch.randn(M,K,device=dev); W=torch.randn(N,K,device=dev); C=torch.zeros(M,N,device=dev)
    for _ in range(3): O.gemm(A,W,C,M,N,K,lda=K,ldb=K,ldc=N,transB=True,engine=2,**kw)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(4e6)); e0.record()
    for _ in range(10): O.gemm(A,W,C,M,N,K,lda=K,ldb=K,ldc=N,transB=True,engine=2,**kw)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    o=(ctypes.c_longlong*16)(); L.load().satk_debug_phase_cycles(3, o); o=list(o)
    print(M,N,K,kw,"%.1f us"%(ms*1e3))
    print("   "+"  ".join("%s=%d"%(n,v) for n,v in zip(names,o[:13])))
