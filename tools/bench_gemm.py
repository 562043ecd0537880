"""Device-time micro-benchmark of the dense tile at the step's recurring shapes (CUDA events, warm L2 / rotating buffers)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import satk_path
satk = satk_path.load()
from importlib import import_module
O = import_module("self-attention-tacotron_b200.ops")

def timeit(fn, reps=20):
    for _ in range(3): fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(1.2e7))     # keep the GPU busy (~6 ms) while the host queues every call: device time only
    e0.record()
    for i in range(reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

NB = 6   # rotate buffers so that operands do not sit in L2 from the previous call
def case(M, N, K, engine, **kw):
    As = [torch.randn(M, K, device="cuda") for _ in range(NB)]
    Wt = torch.randn(N, K, device="cuda")
    Cs = [torch.zeros(M, N, device="cuda") for _ in range(NB)]
    res = torch.randn(M, N, device="cuda") if kw.pop("residual", False) else None
    bias = torch.randn(N, device="cuda") if kw.pop("bias", False) else None
    def f(i):
        O.gemm(As[i % NB], Wt, Cs[i % NB], M, N, K, lda=K, ldb=K, ldc=N, transB=True, residual=res, ldres=N, bias=bias, engine=engine, **kw)
    us = timeit(f)
    print(f"M={M:6d} N={N:5d} K={K:5d} engine={engine} {kw} res={res is not None} bias={bias is not None}: {us:8.1f} us  {2*M*N*K/us/1e6:7.1f} TFLOP/s", flush=True)

for eng in (2,):
    case(12800, 256, 256, eng)
    case(12800, 256, 256, eng, bias=True, act="tanh")
    case(12800, 256, 256, eng, residual=True)
    case(12800, 256, 256, eng, beta=1.0)
    case(12800, 1024, 544, eng, bias=True)
    case(12800, 1024, 256, eng)
    case(12800, 544, 1024, eng)
    case(12800, 160, 256, eng, bias=True)
    case(4736, 128, 128, eng)
    case(4736, 512, 128, eng, bias=True)
    case(4736, 128, 2048, eng)
    case(12800, 256, 256, eng, split_k=2)
    case(4736, 128, 128, eng, split_k=4)
    case(4736, 2048, 128, eng)
