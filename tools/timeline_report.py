"""Reads a chrome trace written by tools/timeline.py and prints, for the last full training step, the per-stream busy time, the
gaps on the main stream and the longest kernels."""
import gzip, json, sys, collections
p = sys.argv[1]
tr = json.load(gzip.open(p) if p.endswith(".gz") else open(p))
ev = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
ev.sort(key=lambda e: e["ts"])
starts = [i for i, e in enumerate(ev) if "embedding_fwd" in e["name"]]
a = starts[-1]
sub = ev[a:]
# the step ends at adam_clip
end = max(i for i, e in enumerate(sub) if "adam_clip" in e["name"])
sub = sub[:end + 1]
t0 = sub[0]["ts"]; t1 = sub[-1]["ts"] + sub[-1]["dur"]
print(f"step: {len(sub)} device activities, {(t1 - t0) / 1e3:.3f} ms")
by = collections.defaultdict(list)
for e in sub: by[e["args"].get("stream")].append(e)
for s, es in by.items():
    busy = sum(e["dur"] for e in es)
    print(f" stream {s}: {len(es)} activities, busy {busy / 1e3:.3f} ms, first {(es[0]['ts'] - t0) / 1e3:.3f} last end {(es[-1]['ts'] + es[-1]['dur'] - t0) / 1e3:.3f}")
if "-v" in sys.argv:
    for e in sub:
        print(f"{(e['ts'] - t0) / 1e3:8.3f} {e['dur']:8.1f} s{e['args'].get('stream')} {e['name'][:70]} g{e['args'].get('grid')}")
# union busy time (any stream) and idle gaps
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in sub)
cur_s, cur_e = iv[0]; idle = 0.0; gaps = []
for s, e in iv[1:]:
    if s > cur_e:
        idle += s - cur_e; gaps.append((s - cur_e, cur_e - t0)); cur_e = e
    else: cur_e = max(cur_e, e)
print(f"device idle (no kernel on any stream): {idle / 1e3:.3f} ms in {len(gaps)} gaps; largest: {[ (round(g,1), round(t/1e3,3)) for g, t in sorted(gaps, reverse=True)[:8]]}")
