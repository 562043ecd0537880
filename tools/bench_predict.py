"""BASELINE.json configs[4]: free-running inference (predict_mel.py path), B=16, T_text=148, <= 1000 mel frames (500 decoder
steps), stop token disabled for timing.  Prints one JSON line: mel-frames/s through engine.predict (encoder + 500 graph-replayed
decoder steps), CUDA-event timed, plus the per-step time."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import satk_path
satk = satk_path.load()
from importlib import import_module
E = import_module("self-attention-tacotron_b200.engine")
O = import_module("self-attention-tacotron_b200.ops")
B, TT, T = int(os.environ.get("B", 16)), 148, int(os.environ.get("T", 500))
hp = satk.load_hparams(os.path.join(ROOT, "examples", "ljspeech_self-attention-tacotron.json"))
eng = E.TacotronEngine(hp, "cuda", seed=1)
d = eng.d
batches = [satk.synthetic_batch(hp, B, TT, 8 * d.r, seed=50 + i, device="cuda")[0] for i in range(3)]
res = {}
for mode, use_graph in (("graph", True), ("eager", False)):
    for i in range(2):
        eng.predict(batches[i % 3], max_iters=T, use_stop_token=False, use_graph=use_graph)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3
    l0 = O.launches()
    t0 = time.perf_counter()
    e0.record()
    for i in range(n):
        out = eng.predict(batches[i % 3], max_iters=T, use_stop_token=False, use_graph=use_graph)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    res[mode] = {"ms_per_utterance_batch": ms, "us_per_decoder_step": ms * 1e3 / T, "mel_frames_per_s": B * T * d.r / (ms * 1e-3),
                 "wall_ms": (time.perf_counter() - t0) * 1e3 / n, "host_launch_calls": (O.launches() - l0) // n}
print(json.dumps({"metric": "free_running_mel_frames_per_sec", "config": f"B={B} T_text={TT} steps={T} (r={d.r}, {T * d.r} frames), stop token off",
                  "finite": bool(torch.isfinite(out["mel"]).all()), **res}))
