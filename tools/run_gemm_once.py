"""Runs the dense tile once per shape (used under ncu --set full to capture the tcgen05 kernel)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import satk_path; satk = satk_path.load()
from importlib import import_module
O = import_module("self-attention-tacotron_b200.ops")
for (M, N, K) in ((12800, 1024, 544), (12800, 256, 256)):
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); C = torch.empty(M, N, device="cuda")
    b = torch.randn(N, device="cuda")
    for _ in range(2):
        O.gemm(A, W, C, M, N, K, lda=K, ldb=K, ldc=N, transB=True, bias=b, engine=2)
torch.cuda.synchronize()
