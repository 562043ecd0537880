"""Diagnostic: gradient difference between the attention-RNN backward walk over all Td steps and the walk that starts at the last
loss step (satk_attn_rnn_bwd_desc.step_end), per parameter tensor, next to the run-to-run noise of the full walk."""
import os
import sys
from importlib import import_module

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import satk_path  # noqa: E402

satk = satk_path.load()

E = import_module("self-attention-tacotron_b200.engine")
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
hp = satk.load_hparams(os.path.join(root, "examples", "ljspeech_self-attention-tacotron.json"))
d = satk.dims_from_hparams(hp)
ps = satk.ParamStore(d).init(5, "glorot")
f, l = satk.synthetic_batch(hp, 32, 148, 800, seed=int(os.environ.get("DIAG_SEED", "78")), device="cuda")
masks = satk.make_masks(d, 32, 148, 400, seed=3, device="cuda")
res = []
for skip, tc in (("0", "1"), ("0", "1"), ("1", "1"), ("0", "1"), ("0", "0"), ("0", "0")):
    os.environ["SATK_STEP_END"] = skip
    os.environ["SATK_ATTN_TC"] = tc
    eng = E.TacotronEngine(hp, "cuda", params=ps)
    eng.forward(f, l, True, masks)
    eng.backward()
    torch.cuda.synchronize()
    res.append({n: g.clone() for n, g in eng.ps.g.items()})
    if skip == "1":
        se = eng.saved["step_end"]
        print("step_end", se.tolist())
        for k in [k for k in eng._bufs if k.startswith("dec.dx_lstm") or k.startswith("dec.dgates")] + ["dec.dmel_tm", "dec.dstop_tm", "dec.dproj_in"]:
            v = eng._bufs[k].view(400, 32, -1)
            worst = max(float(v[int(se[b]):, b].abs().max()) if int(se[b]) < 400 else 0.0 for b in range(32))
            print("max |%s| in skipped rows:" % k, worst)
def rel(a, b, n):
    return (res[a][n] - res[b][n]).double().norm().item() / (res[b][n].double().norm().item() + 1e-30)


rows = [(n, rel(1, 0, n), rel(3, 1, n), rel(2, 1, n), rel(5, 4, n), rel(4, 1, n), res[0][n].double().norm().item()) for n in res[0]]
rows.sort(key=lambda r: -r[1])
print("%-34s %10s %10s %10s %10s %10s %10s" % ("tensor", "run1-run0", "run3-run1", "skip-run1", "simt5-4", "simt-tc", "norm"))
for r in rows[:12]:
    print("%-34s %10.2e %10.2e %10.2e %10.2e %10.2e %10.2e" % r)
rows.sort(key=lambda r: -r[3])
print("sorted by skip-run1")
for r in rows[:12]:
    print("%-34s %10.2e %10.2e %10.2e %10.2e %10.2e %10.2e" % r)
tot = lambda a, b: (sum(((res[a][n] - res[b][n]).double() ** 2).sum().item() for n in res[0]) / sum((res[b][n].double() ** 2).sum().item() for n in res[0])) ** 0.5  # noqa: E731
print("whole-vector: run1-run0 %.2e run3-run1 %.2e skip-run1 %.2e simt5-simt4 %.2e simt-tc %.2e" % (tot(1, 0), tot(3, 1), tot(2, 1), tot(5, 4), tot(4, 1)))
