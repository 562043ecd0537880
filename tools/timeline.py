"""Kernel timeline of one training step at config 2 (B=32, Tt=148, Tm=800) through the CUPTI activity records of torch.profiler:
every kernel with its stream, start and duration -> gpurun_out/timeline_<tag>.json.  (No nsys in the image.)"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import satk_path
satk = satk_path.load()
from importlib import import_module
E = import_module("self-attention-tacotron_b200.engine")
from torch.profiler import profile, ProfilerActivity
tag = sys.argv[1] if len(sys.argv) > 1 else "step"
hp = satk.load_hparams(os.path.join(ROOT, "examples", "ljspeech_self-attention-tacotron.json"))
eng = E.TacotronEngine(hp, "cuda", seed=2)
f, l = satk.synthetic_batch(hp, 32, 148, 800, seed=77, device="cuda")
for _ in range(5):
    eng.train_step(f, l)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        eng.train_step(f, l)
    torch.cuda.synchronize()
ev = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        ev.append({"name": e.name[:80], "start_us": e.time_range.start, "dur_us": e.time_range.elapsed_us(), "stream": getattr(e, "device_resource_id", None) if hasattr(e, "device_resource_id") else None})
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
prof.export_chrome_trace(os.path.join(ROOT, "gpurun_out", f"trace_{tag}.json"))
print(len(ev), "device events")
