import sys,os; sys.path.insert(0,'.')
import torch, ctypes
import satk_path; satk=satk_path.load()
from importlib import import_module
E=import_module("self-attention-tacotron_b200.engine"); L=import_module("self-attention-tacotron_b200.lib")
hp = satk.load_hparams("examples/ljspeech_self-attention-tacotron.json")
eng = E.TacotronEngine(hp, "cuda", seed=1)
f, l = satk.synthetic_batch(hp, 28, 148, 800, seed=3, device="cuda")
for _ in range(2):
    eng.forward(f, l, True); eng.backward()
torch.cuda.synchronize()
o=(ctypes.c_longlong*16)(); L.load().satk_debug_phase_cycles(1, o)
names=["waitX","P1gemm","sync1","pointwise+send","sync2","saverA+waitO","fS+qpart","qfin+2sync","energies","fence+sync","bulk+qsave+waitE","softmax","sync3","ctxpart+sync","ctx send","top"]
print("FWD", [(n,v) for n,v in zip(names,list(o))], "sum", sum(list(o)))
L.load().satk_debug_phase_cycles(2, o)
names=["top+waitC+sync","BA1+bulk","waitW","softmax_bwd","energy_bwd","finalize+send","waitQ","BB","waitG(+saver)","BC"]+["-"]*5+["looptop"]
print("BWD", [(n,v) for n,v in zip(names,list(o)) if n!="-"], "sum", sum(list(o)))
