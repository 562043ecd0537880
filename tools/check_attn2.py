"""A/B of the attention-RNN kernel generations on the GPU: same inputs / weights / masks through the first-generation kernels
(SATK_ATTN_GEN=1) and the second-generation ones (4 and 5 utterances per cluster), outputs and saved tensors compared, kernel
times from CUDA events.  `python tools/check_attn2.py [bwd] [phases]`."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import satk_path  # noqa: E402

satk = satk_path.load()
from importlib import import_module  # noqa: E402

E = import_module("self-attention-tacotron_b200.engine")
L = import_module("self-attention-tacotron_b200.lib")

FWD_BUFS = ["dec.x2", "dec.align1", "dec.align2", "dec.soft1", "dec.qsave", "dec.gates1", "dec.cprev1", "dec.hprev1"]
BWD_BUFS = ["dec.dgates1", "dec.dq", "dec.dkeys1", "dec.dkeys2", "dec.dx_lstm2"]
GRADS = ["att1.v", "att2.v", "att1.loc_conv.W", "att1.loc_conv.b", "att1.loc_layer.W", "att1.b"]


def run(eng, f, l, masks, gen, nb, bwd):
    os.environ["SATK_ATTN_GEN"] = str(gen)
    if nb:
        os.environ["SATK_ATTN_NB"] = str(nb)
    else:
        os.environ.pop("SATK_ATTN_NB", None)
    eng.timers = {}
    out = {}
    for it in range(3):
        eng.ps.grad.zero_()
        eng.forward(f, l, True, masks)
        if bwd:
            eng.backward()
    torch.cuda.synchronize()
    for k in FWD_BUFS + (BWD_BUFS if bwd else []):
        if k in eng._bufs:
            out[k] = eng._bufs[k].clone()
    if bwd:
        for k in GRADS:
            out["g:" + k] = eng.ps.g[k].clone()
        out["g:flat"] = eng.ps.grad.clone()
    times = {k: min(a.elapsed_time(b) for a, b in evs[1:]) for k, evs in eng.timers.items() if k.startswith("attn")}
    return out, times


def main():
    bwd = "bwd" in sys.argv
    hp = satk.load_hparams(os.path.join(ROOT, "examples", "ljspeech_self-attention-tacotron.json"))
    d = satk.dims_from_hparams(hp)
    print(L.device_info(), flush=True)
    shapes = [(5, 30, 40), (7, 148, 120), (3, 61, 64), (32, 148, 800)]
    if "big" in sys.argv:
        shapes = [(32, 148, 800), (64, 148, 800)]
    if "small" in sys.argv:
        shapes = [(5, 30, 40)]
    for (B, Tt, Tm) in shapes:
        ps = satk.ParamStore(d).init(7, "random")
        eng = E.TacotronEngine(hp, "cuda", params=ps)
        f, l = satk.synthetic_batch(hp, B, Tt, Tm, seed=3, device="cuda")
        masks = {k: v.cuda() for k, v in satk.make_masks(d, B, Tt, Tm // d.r, seed=5).items()}
        ref, t1 = run(eng, f, l, masks, 1, 0, bwd)
        print(f"B={B} Tt={Tt} Tm={Tm}: gen1 {t1}", flush=True)
        for nb in (4, 5):
            got, t2 = run(eng, f, l, masks, 2, nb, bwd)
            worst = 0.0
            msg = []
            for k in ref:
                a, b = ref[k], got[k]
                err = (a - b).abs().max().item()
                sc = a.abs().max().item()
                rel = err / max(sc, 1e-30)
                worst = max(worst, rel)
                msg.append(f"{k}:{err:.1e}/{sc:.1e}")
            print(f"  gen2 NB={nb}: {t2}  worst rel {worst:.2e}", flush=True)
            print("    " + " ".join(msg), flush=True)
    if "phases" in sys.argv:
        o = (ctypes.c_longlong * 16)()
        for which, nm in ((4, "FWD2"), (5, "BWD2")):
            try:
                L.check(L.load().satk_debug_phase_cycles(which, o), "phase")
                print(nm, list(o), "sum", sum(list(o)))
            except Exception as ex:  # noqa: BLE001
                print(nm, "no phase data:", ex)


if __name__ == "__main__":
    main()
