"""Attention-RNN kernel time as a function of the source lengths (single wave, B=28): all 148 / all 100 / mixed."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import satk_path
satk = satk_path.load()
from importlib import import_module
E = import_module("self-attention-tacotron_b200.engine")
hp = satk.load_hparams(os.path.join(ROOT, "examples", "ljspeech_self-attention-tacotron.json"))
eng = E.TacotronEngine(hp, "cuda", seed=1)
for B in (28, 32):
    for name, L in (("all 148", 148), ("all 128", 128), ("all 100", 100), ("all 64", 64)):
        f, l = satk.synthetic_batch(hp, B, 148, 800, seed=3, device="cuda", full_length=True)
        sl = torch.full_like(f.source_length, L)
        src = f.source.clone(); src[:, L:] = 0
        f = f._replace(source=src, source_length=sl)
        eng.timers = {}
        for _ in range(4):
            eng.forward(f, l, True); eng.backward()
        torch.cuda.synchronize()
        ts = {k: min(a.elapsed_time(b) for a, b in v[1:]) for k, v in eng.timers.items() if k.startswith("attn")}
        print(f"B={B} {name}: " + "  ".join(f"{k} {v:.3f} ms" for k, v in ts.items()), flush=True)
