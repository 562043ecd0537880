"""Device-time micro-benchmark of the recurrent kernels at the BASELINE config-2 shapes (CUDA events)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import satk_path  # noqa: E402

satk = satk_path.load()
from importlib import import_module  # noqa: E402

O = import_module("self-attention-tacotron_b200.ops")
E = import_module("self-attention-tacotron_b200.engine")


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)


def lstm(H, T, B):
    dev = "cuda"
    xg = torch.randn(T * B, 4 * H, device=dev) * 0.3
    Wh = torch.randn(H, 4 * H, device=dev) * 0.05
    out = torch.empty(T, B, H, device=dev)
    gates, cp, hp = torch.empty(T * B, 4 * H, device=dev), torch.empty(T * B, H, device=dev), torch.empty(T * B, H, device=dev)
    mc = (torch.rand(T, B, H, device=dev) < 0.9).to(torch.uint8)
    mh = (torch.rand(T, B, H, device=dev) < 0.9).to(torch.uint8)
    f = lambda: O.lstm_seq_fwd(xg, Wh, out, T, B, H, mask_c=mc, mask_h=mh, gates=gates, c_prev=cp, h_prev=hp)
    mn, av = timeit(f)
    print(f"lstm_fwd H={H} T={T} B={B}: {mn:.3f} ms  ({1e3 * mn / T:.2f} us/step)", flush=True)
    if H == 256:
        ref = out.clone()
        os.environ["SATK_LSTM_GEN"] = "1"
        mn, av = timeit(f)
        os.environ.pop("SATK_LSTM_GEN")
        print(f"lstm_fwd (first generation) H={H} T={T} B={B}: {mn:.3f} ms; max |diff| {(out - ref).abs().max().item():.2e}", flush=True)
        f()
    dout, dg = torch.randn(T, B, H, device=dev), torch.empty(T * B, 4 * H, device=dev)
    f = lambda: O.lstm_seq_bwd(Wh, gates, cp, dout, dg, T, B, H, mask_c=mc, mask_h=mh)
    mn, av = timeit(f)
    print(f"lstm_bwd H={H} T={T} B={B}: {mn:.3f} ms  ({1e3 * mn / T:.2f} us/step)", flush=True)
    if H == 256:
        ref = dg.clone()
        os.environ["SATK_LSTM_GEN"] = "1"
        mn, av = timeit(f)
        os.environ.pop("SATK_LSTM_GEN")
        print(f"lstm_bwd (first generation) H={H} T={T} B={B}: {mn:.3f} ms; max |diff| {(dg - ref).abs().max().item():.2e} of {ref.abs().max().item():.2e}", flush=True)


def attn(B=32, Tt=148, Tm=800):
    hp = satk.load_hparams(os.path.join(ROOT, "examples", "ljspeech_self-attention-tacotron.json"))
    eng = E.TacotronEngine(hp, "cuda", seed=1)
    f, l = satk.synthetic_batch(hp, B, Tt, Tm, seed=3, device="cuda")
    eng.timers = {}
    for _ in range(4):
        eng.forward(f, l, True)
        eng.backward()
    torch.cuda.synchronize()
    Td = Tm // 2
    for k, evs in eng.timers.items():
        ts = [a.elapsed_time(b) for a, b in evs][1:]
        print(f"{k} B={B}: {min(ts):.3f} ms  ({1e3 * min(ts) / Td:.2f} us/step)", flush=True)


if __name__ == "__main__":
    print(O.L.device_info())
    lstm(256, 400, 32)
    lstm(256, 400, 28)
    lstm(128, 148, 32)
    if "noattn" not in sys.argv:
        attn()
        attn(B=28)
