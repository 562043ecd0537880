#!/usr/bin/env python
"""Benchmark of the teacher-forced training hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this implementation (CUDA, sm_100a)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle restatement of the
                                                           # reference's TF1 graph on the host cores

A "step" is one full train step (forward, backward, gradient all-reduce when N>1, clip + Adam) on a
synthetic LJSpeech-shaped batch: configs[1] of BASELINE.json = examples/ljspeech self-attention
Tacotron, B=32 per GPU, T_text=148, T_mel=800, 80 mel channels (weak scaling: MirroredStrategy
semantics, per-replica batch = batch_size).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "teacher_forced_mel_frames_per_sec"
UNIT = "mel-frames/s"
# BASELINE.json configs: name -> (example json, hparams overrides, per-GPU batch, T_text, T_mel, mode)
WORKLOADS = {
    "ljspeech": ("ljspeech_self-attention-tacotron.json", None, 32, 148, 800, "train"),                      # configs[1]: headline
    "vctk": ("vctk_self-attention-tacotron.json", None, 64, 148, 800, "train"),                              # configs[2]: per-replica 64
    "location_sensitive": ("ljspeech_self-attention-tacotron.json", "attention=location_sensitive", 32, 148, 800, "train"),   # configs[3]
    "transition_agent": ("ljspeech_self-attention-tacotron.json", "use_forward_attention_transition_agent=True", 32, 148, 800, "train"),
    "predict": ("ljspeech_self-attention-tacotron.json", None, 16, 148, 1000, "predict"),                    # configs[4]: free-running
    "postnet_v2": ("ljspeech_self-attention-tacotron.json", "use_postnet_v2=True", 32, 148, 800, "train"),   # SURVEY f3 option (off in the public configs)
}
CFG, OVR, B, TT, TM, MODE = WORKLOADS["ljspeech"]


def select_workload(name):
    global CFG, OVR, B, TT, TM, MODE
    CFG, OVR, B, TT, TM, MODE = WORKLOADS[name]


def workload_name():
    return f"{CFG}{' ' + OVR if OVR else ''} B={B}/GPU T_text={TT} T_mel={TM} n_mels=80"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def oracle_sample(hp, satk, steps, warmup, sample_b):
    """CPU arm: full train step (fwd, autograd bwd, clip, Adam) of the oracle on a bounded sample."""
    import torch
    from oracle import model as OR
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    d = satk.dims_from_hparams(hp)
    ps = satk.ParamStore(d).init(1234, "glorot")
    tr = OR.OracleTrainer(d, hp, ps.as_dict())
    f, l = satk.synthetic_batch(hp, sample_b, TT, TM, seed=1234)
    masks = satk.make_masks(d, sample_b, TT, TM // d.r, seed=99)
    for _ in range(warmup):
        tr.train_step(f, l, masks)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.train_step(f, l, masks)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return sample_b * TM / dt, dt, cores


def run_reference(args, rank):
    """CPU arm (rank 0 only; ONE replica however many GPUs the other arm uses): the oracle's full train step on the SAME
    configuration and batch size as the GPU arm, with the requested warm-up."""
    if rank != 0:
        return
    import satk_path
    satk = satk_path.load()
    hp = satk.load_hparams(os.path.join(ROOT, "examples", CFG), OVR)
    if MODE != "train":
        print(json.dumps({"impl": "reference", "unavailable": "the CPU arm times the teacher-forced train step only"}), flush=True)
        return
    value, dt, cores = oracle_sample(hp, satk, args.steps, args.warmup, B)
    sample = (f"the full per-replica batch ({B} utterances, T_text={TT}, T_mel={TM}), full train step per step, {args.warmup} warm-up steps; "
              "torch-CPU fp32 restatement of the TF1 graph (TF1 itself cannot run here); one CPU replica regardless of --gpus")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "n_cpu_replicas": 1,
        "config": {"workload": f"{workload_name()} teacher-forced train step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def run_predict(args, rank, world, local_rank):
    """BASELINE.json configs[4]: free-running inference (predict_mel.py path), B=16, <= 1000 mel frames (500 decoder steps), stop
    token disabled for timing; every rank decodes its own batch (replicas, no collective)."""
    import torch
    import satk_path
    satk = satk_path.load()
    from importlib import import_module
    E = import_module("self-attention-tacotron_b200.engine")
    O = import_module("self-attention-tacotron_b200.ops")
    torch.cuda.set_device(local_rank)
    hp = satk.load_hparams(os.path.join(ROOT, "examples", CFG), OVR)
    eng = E.TacotronEngine(hp, f"cuda:{local_rank}", seed=1)
    d = eng.d
    T = TM // d.r
    host = [satk.synthetic_batch(hp, B, TT, 8 * d.r, seed=50 + i)[0] for i in range(3)]
    host = [satk.SourceData(*[x.pin_memory() if torch.is_tensor(x) else x for x in f]) for f in host]
    devb = [satk.SourceData(*[x.cuda() if torch.is_tensor(x) else x for x in f]) for f in host]
    for i in range(max(args.warmup, 3)):
        eng.predict(devb[i % 3], max_iters=T, use_stop_token=False)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    n = max(1, min(args.steps, 5))
    l0 = O.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        out = eng.predict(devb[i % 3], max_iters=T, use_stop_token=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    launches = O.launches() - l0
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    d2h = 0
    for i in range(n):
        f = satk.SourceData(*[x.cuda(non_blocking=True) if torch.is_tensor(x) else x for x in host[i % 3]])
        mel = eng.predict(f, max_iters=T, use_stop_token=False)["mel"].cpu()       # the mel frames are the result
        d2h = mel.numel() * 4
    e3.record()
    torch.cuda.synchronize()
    ms_e2e = e2.elapsed_time(e3) / n
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        return
    h2d = sum(x.numel() * x.element_size() for x in host[0] if torch.is_tensor(x))
    print(json.dumps({
        "metric": "free_running_mel_frames_per_sec", "value": world * B * T * d.r / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": n,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": f"{workload_name()} free-running decode, {T} decoder steps (r={d.r}), stop token off; "
                                                    "a step = one batch of utterances decoded", "parallelism": f"replicas x{world}"},
        "e2e": {"value": world * B * T * d.r / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e},
        "gpu_launches": launches, "clocks": clocks, "us_per_decoder_step": ms * 1e3 / T, "finite": bool(torch.isfinite(out["mel"]).all()),
        "roofline": None,
    }), flush=True)


def run_satk(args, rank, world, local_rank):
    import torch
    import satk_path
    satk = satk_path.load()
    from importlib import import_module
    M = import_module("self-attention-tacotron_b200.models")
    O = import_module("self-attention-tacotron_b200.ops")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    hp = satk.load_hparams(os.path.join(ROOT, "examples", CFG), OVR)

    def allreduce(flat, async_op=False):
        # ncclAllReduce(sum) over (a bucket of) the flat fp32 gradient buffer; the engine sends the decoder bucket while the
        # encoder's backward pass runs (engine.backward)
        return dist.all_reduce(flat, async_op=async_op)
    allreduce.supports_async = True

    model = M.tacotron_model_factory(hp, None, None, device=str(dev), allreduce=allreduce if world > 1 else None, world_size=world)
    eng = model.engine
    nb = 4   # distinct batches rotated through the timed region
    host, devb = [], []
    for i in range(nb):
        f, l = satk.synthetic_batch(hp, B, TT, TM, seed=1234 + 100 * rank + i)
        fp = satk.SourceData(*[x.pin_memory() if torch.is_tensor(x) else x for x in f])
        lp = satk.MelData(*[x.pin_memory() if torch.is_tensor(x) else x for x in l])
        host.append((fp, lp))
        devb.append((satk.SourceData(*[x.to(dev) if torch.is_tensor(x) else x for x in f]),
                     satk.MelData(*[x.to(dev) if torch.is_tensor(x) else x for x in l])))
    h2d = sum(x.numel() * x.element_size() for nt in host[0] for x in nt if torch.is_tensor(x))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        f, l = devb[i % nb]
        eng.train_step(f, l, None, allreduce=allreduce if world > 1 else None, world_size=world)
    barrier()

    # ---- timed region 1: inputs resident in HBM
    # ---- per-kernel device times (CUDA events around the launches) for the roofline: a short eager pass, because the timed
    # regions below replay the step from a CUDA graph (events cannot be read back from inside a graph)
    eng.timers = {}
    for i in range(4):
        f, l = devb[i % nb]
        eng.train_step(f, l, None, allreduce=allreduce if world > 1 else None, world_size=world)
    barrier()
    timers = eng.timers
    eng.timers = None
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = O.launches()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        f, l = devb[i % nb]
        eng.train_step(f, l, None, allreduce=allreduce if world > 1 else None, world_size=world)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = O.launches() - l0
    # ---- timed region 2: end to end through the Estimator surface, host (pinned) buffers in, loss out
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    loss = 0.0
    # device -> host read of EVERY step's loss (4 bytes) inside the timed region: an asynchronous copy into pinned memory right behind
    # the step, consumed once the next step has been enqueued (a blocking read per step would leave the GPU idle while the host
    # enqueues the ~100 short encoder launches of the next step); SATK_E2E_BLOCKING=1 restores the blocking read
    blocking = os.environ.get("SATK_E2E_BLOCKING", "0") == "1"
    pin = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    evs = [torch.cuda.Event(), torch.cuda.Event()]
    pending = None
    for i in range(args.steps):
        f, l = host[i % nb]
        spec = model.model_fn(f, l, M.ModeKeys.TRAIN, hp)
        if blocking:
            loss = float(spec.loss)
            continue
        pin[i % 2].copy_(spec.loss.reshape(1), non_blocking=True)
        evs[i % 2].record()
        if pending is not None:
            evs[pending].synchronize()
            loss = float(pin[pending])
        pending = i % 2
    if pending is not None:
        evs[pending].synchronize()
        loss = float(pin[pending])
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    clocks = sampler.stop() if sampler else None
    # ---- timed region 3: the same step with EVERY target at full length (all Td decoder steps carry a loss: nothing for the
    # backward walk to skip): the length distribution of the synthetic batch does not flatter `value`
    nfl = max(1, min(args.steps, 5))
    full = [tuple(satk.synthetic_batch(hp, B, TT, TM, seed=4321 + 100 * rank + i, device=dev, full_length=True)) for i in range(2)]
    for i in range(2):
        eng.train_step(full[i][0], full[i][1], None, allreduce=allreduce if world > 1 else None, world_size=world)
    barrier()
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record()
    for i in range(nfl):
        eng.train_step(full[i % 2][0], full[i % 2][1], None, allreduce=allreduce if world > 1 else None, world_size=world)
    e5.record()
    barrier()
    ms_full = e4.elapsed_time(e5) / nfl
    t = torch.tensor([ms, ms_e2e, ms_full], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_full = t.tolist()
    def shutdown():
        # captured graphs hold NCCL resources: release them before the process group goes away, and never let a teardown
        # problem keep the ranks alive (the result line is already out)
        if world > 1:
            killer = threading.Timer(20.0, lambda: os._exit(0))
            killer.daemon = True
            killer.start()
            eng.release_graphs()
            dist.destroy_process_group()
            killer.cancel()

    if rank != 0:
        shutdown()
        return
    frames = world * B * TM * args.steps
    value = frames / (ms * 1e-3)
    e2e = frames / (ms_e2e * 1e-3)
    # ---- roofline of the dominant kernel (attention-RNN forward / backward), CUDA-event timed in region 1
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    d = eng.d
    td = TM // d.r
    w_seq = (d.ctx + d.att_rnn) * 4 * d.att_rnn + d.att_rnn * d.att1 + d.att_filters * d.att1 + d.att_kernel * d.att_filters \
        + d.att_rnn * d.att2
    bytes_step = 2 * (w_seq + B * TT * (d.att1 + d.mem1 + d.att2 + d.mem2))      # SURVEY §8(d), LSTM-2/3 rows excluded
    kt = {}
    for name, evs in (timers or {}).items():
        torch.cuda.synchronize()
        kt[name] = statistics.median(a.elapsed_time(b) for a, b in evs)    # (a first use inside this eager pass can carry a one-time cost)
    roof = None
    # fraction of the (utterance, decoder step) pairs that carry a loss: the backward walk skips the others (engine.forward: step_end)
    loss_frac = statistics.mean(float((l_.binary_loss_mask != 0).sum()) / (B * td) for _, l_ in host)
    frames_with_loss = world * args.steps * statistics.mean(float((l_.spec_loss_mask != 0).sum()) for _, l_ in host)
    if kt:
        sections = {k[4:]: v for k, v in kt.items() if k.startswith("sec.")}
        kt = {k: v for k, v in kt.items() if not k.startswith("sec.")}
        kv = B * TT * (d.att1 + d.mem1 + d.att2 + d.mem2)

        def kernel_roof(name):
            # algorithmic bytes per launch: forward = bytes_step * Td; backward = 2x that, its key / value
            # term counted only for the pairs that are walked (the weights are algorithmic work for all Td steps)
            if name.endswith("bwd"):
                lf = loss_frac if getattr(eng, "skip_masked_steps", False) else 1.0
                alg = 2 * td * 2 * (w_seq + lf * kv)
            else:
                alg = bytes_step * td
            ach = alg / (kt[name] * 1e-3) / 1e9
            return {"achieved": ach, "frac": ach / hbm, "algorithmic_bytes_per_launch": alg, "avg_launch_ms": kt[name],
                    "us_per_decoder_step": kt[name] * 1e3 / td}
        rk = {n: kernel_roof(n) for n in ("attn_rnn_fwd", "attn_rnn_bwd") if n in kt}
        dom = max(rk, key=lambda n: kt[n]) if rk else max(kt, key=kt.get)
        traffic = None
        try:   # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (same shapes only)
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(dom)
        except Exception:
            pass
        r0 = rk.get(dom, {"achieved": None, "frac": None})
        roof = {"kernel": dom, "bound": "hbm", "achieved": r0["achieved"], "peak": hbm, "unit": "GB/s", "frac": r0["frac"], "traffic": traffic,
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                "algorithmic_bytes_per_launch": r0.get("algorithmic_bytes_per_launch"), "steps_with_loss_frac": loss_frac,
                "avg_launch_ms": kt[dom], "us_per_decoder_step": kt[dom] * 1e3 / td,
                "kernels": rk, "kernel_ms": kt, "section_ms": sections}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{workload_name()} teacher-forced train step (fwd+bwd+allreduce+clip+Adam)", "parallelism": f"dp{world}",
                   "l2": "working set per step (>1 GB of activations) exceeds the 126 MB L2; 4 distinct batches rotated",
                   "cuda_graph": bool(getattr(eng, "use_graph", False) and any(g.get("graph") is not None for g in eng._graphs.values()))},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                "last_loss": loss,
                "loss_readback": "blocking float(loss) per step" if blocking else
                "every step's loss copied to pinned memory behind the step and read after the next step is enqueued"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof,
        # the same step with all targets at full length (no masked decoder steps to skip), and the padded-frame-free variant of `value`
        "value_full_length": world * B * TM / (ms_full * 1e-3), "ms_per_step_full_length": ms_full,
        "frames_with_loss_per_sec": frames_with_loss / (ms * 1e-3),
    }
    if world == 1 and not args.no_cpu_baseline:
        v, dt, cores = oracle_sample(hp, satk, 1, 1, B)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"1 timed train step (after 1 warm-up) on the full batch ({B} utterances, T_text={TT}, T_mel={TM}); "
                                         "oracle = torch-CPU fp32 restatement of the TF1 graph"}
    print(json.dumps(out), flush=True)
    shutdown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="satk", choices=["satk", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="ljspeech", choices=sorted(WORKLOADS),
                    help="BASELINE.json workload: ljspeech = configs[1] (headline, default), vctk = configs[2], location_sensitive / "
                         "transition_agent = configs[3] variants, predict = configs[4]")
    args = ap.parse_args()
    select_workload(args.config)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
    elif MODE == "predict":
        run_predict(args, rank, world, local_rank)
    else:
        run_satk(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
