"""Runs the full-size forward pass a few times (used under ncu to capture the attention-RNN kernels)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import satk_path; satk = satk_path.load()
from importlib import import_module
E = import_module("self-attention-tacotron_b200.engine")
hp = satk.load_hparams("examples/ljspeech_self-attention-tacotron.json")
eng = E.TacotronEngine(hp, "cuda", seed=1)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 28
f, l = satk.synthetic_batch(hp, B, 148, 800, seed=3, device="cuda")
for _ in range(2):
    eng.forward(f, l, True)
    if "bwd" in sys.argv: eng.backward()
torch.cuda.synchronize()
