"""Race check of the multi-stream backward pass: gradients of the same batch with and without the side / auxiliary streams must
agree to reduction-order noise, over repeated runs; then a 40-step soak at full size (finite, decreasing loss)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import satk_path
satk = satk_path.load()
from importlib import import_module
E = import_module("self-attention-tacotron_b200.engine")
hp = satk.load_hparams(os.path.join(ROOT, "examples", "ljspeech_self-attention-tacotron.json"))
d = satk.dims_from_hparams(hp)
ps = satk.ParamStore(d).init(5, "glorot")
f, l = satk.synthetic_batch(hp, 32, 148, 800, seed=77, device="cuda")
masks = satk.make_masks(d, 32, 148, 400, seed=3, device="cuda")
grads = {}
for mode in ("0", "1"):
    os.environ["SATK_WGRAD_STREAM"] = mode
    eng = E.TacotronEngine(hp, "cuda", params=ps)
    runs = []
    for rep in range(4):
        eng.forward(f, l, True, masks)
        eng.backward()
        torch.cuda.synchronize()
        runs.append(eng.ps.grad.clone())
    grads[mode] = runs
ref = grads["0"][0]
scale = ref.abs().max().item()
for mode, runs in grads.items():
    for i, g in enumerate(runs):
        err = (g - ref).abs().max().item()
        rel = ((g - ref).double().norm() / ref.double().norm()).item()
        print(f"streams={mode} run {i}: max|dg| {err:.3e} (scale {scale:.3e}), rel L2 {rel:.3e}, finite {bool(torch.isfinite(g).all())}", flush=True)
        assert rel < 1e-5 and torch.isfinite(g).all()
os.environ["SATK_WGRAD_STREAM"] = "1"
eng = E.TacotronEngine(hp, "cuda", seed=2)
losses = []
for step in range(40):
    out = eng.train_step(f, l)
    losses.append(out["losses"][2].item())
print("soak losses:", [round(x, 4) for x in losses[::5]], flush=True)
assert all(x == x for x in losses) and losses[-1] < losses[0]
print("OK")
