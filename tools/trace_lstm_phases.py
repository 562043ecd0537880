import sys,os; sys.path.insert(0,'.')
import torch, ctypes
import satk_path; satk=satk_path.load()
from importlib import import_module
O=import_module("self-attention-tacotron_b200.ops"); L=import_module("self-attention-tacotron_b200.lib")
for H,T,B in ((256,400,28),(128,148,32)):
    dev="cuda"
    xg = torch.randn(T*B,4*H,device=dev)*0.3; Wh=torch.randn(H,4*H,device=dev)*0.05; out=torch.empty(T,B,H,device=dev)
    gates, cp, hp = torch.empty(T*B,4*H,device=dev), torch.empty(T*B,H,device=dev), torch.empty(T*B,H,device=dev)
    mc=(torch.rand(T,B,H,device=dev)<0.9).to(torch.uint8); mh=(torch.rand(T,B,H,device=dev)<0.9).to(torch.uint8)
    for variant in ("full","nosave","nosave_nomask"):
        kw=dict(mask_c=mc,mask_h=mh,gates=gates,c_prev=cp,h_prev=hp)
        if variant!="full": kw.update(gates=None,c_prev=None,h_prev=None)
        if variant=="nosave_nomask": kw.update(mask_c=None,mask_h=None,zc=0.1,zh=0.1)
        for _ in range(3): O.lstm_seq_fwd(xg,Wh,out,T,B,H,**kw)
        torch.cuda.synchronize()
        o=(ctypes.c_longlong*16)(); L.load().satk_debug_phase_cycles(0, o)
        print(H, variant, "wait,gemm,sync1,pointwise+send,sync2,looptop:", list(o)[:6], "sum", sum(list(o)[:6]))
