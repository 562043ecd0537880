"""A/B of one TF32 pass against 3xTF32 per GEMM group (satk_gemm_desc.precision).

For every group of dense products of the train step (engine.py: tf32_push / tf32_group marks) the group alone, and all groups
together, run with a single TF32 pass on the raw fp32 operands; everything else stays 3xTF32.  Measured against the oracle on
one medium TRAIN batch with injected masks: worst output error relative to the tensor's scale (mel, stop logits, alignments,
losses; budget 1e-3) and the whole-gradient relative L2 error (budget 1e-3); and the step time at config 2.
`python tools/ab_tf32.py > gpurun_out/ab_tf32.json`"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import satk_path  # noqa: E402

satk = satk_path.load()
from importlib import import_module  # noqa: E402

E = import_module("self-attention-tacotron_b200.engine")
O = import_module("self-attention-tacotron_b200.ops")
from oracle import model as OR  # noqa: E402

GROUPS = ["enc", "prenet", "lstmx", "mem", "sa", "proj"]


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def parity(hp, d, ps, f, l, masks, ref, rg, names):
    eng = E.TacotronEngine(hp, "cuda", params=ps)
    eng.sort_batches = False
    fd = satk.SourceData(*[x.cuda() if torch.is_tensor(x) else x for x in f])
    ld = satk.MelData(*[x.cuda() if torch.is_tensor(x) else x for x in l])
    md = {k: v.cuda() for k, v in masks.items()}
    out = eng.forward(fd, ld, True, md)
    B, Tm = l.mel.shape[0], l.mel.shape[1]
    Td = Tm // d.r
    errs = {
        "mel": rel(out["mel_tm"].view(Td, B, d.r, d.n_mels).permute(1, 0, 2, 3).reshape(B, Tm, d.n_mels), ref["mel"]),
        "stop": rel(out["stop_tm"].view(Td, B).t(), ref["stop"].squeeze(-1)),
        "alignment": rel(out["align1_tm"].permute(1, 2, 0), ref["alignment"]),
        "alignment2": rel(out["align2_tm"].permute(1, 2, 0), ref["alignment2"]),
        "loss": abs(out["losses"][2].item() - ref["loss"].item()) / abs(ref["loss"].item()),
    }
    eng.backward()
    num = den = 0.0
    worst = 0.0
    for n in names:
        a, b = eng.ps.g[n].detach().float().cpu(), rg[n].detach().float()
        num += ((a - b).double() ** 2).sum().item()
        den += (b.double() ** 2).sum().item()
        worst = max(worst, ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item())
    return {"worst_output_rel": max(errs.values()), "outputs": errs, "grad_rel_l2": (num / den) ** 0.5, "worst_grad_tensor_rel": worst}


def step_ms(hp, groups):
    O.set_tf32_1x(groups)
    eng = E.TacotronEngine(hp, "cuda", seed=1)
    batches = [satk.synthetic_batch(hp, 32, 148, 800, seed=1234 + i, device="cuda") for i in range(3)]
    for i in range(3):
        eng.train_step(*batches[i % 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(6):
        eng.train_step(*batches[i % 3])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 6


def main():
    hp = satk.load_hparams(os.path.join(ROOT, "examples", "ljspeech_self-attention-tacotron.json"))
    d = satk.dims_from_hparams(hp)
    ps = satk.ParamStore(d).init(7, "random")
    f, l = satk.synthetic_batch(hp, 8, 70, 120, seed=11)
    masks = satk.make_masks(d, 8, 70, 60, seed=8)
    tr = OR.OracleTrainer(d, hp, ps.as_dict())
    ref, rg, _ = tr.loss_and_grads(f, l, masks, True)
    res = {}
    for name, groups in [("3xTF32 everywhere", [])] + [(g, [g]) for g in GROUPS] + [("all groups 1x", ["all"])]:
        O.set_tf32_1x(groups)
        r = parity(hp, d, ps, f, l, masks, ref, rg, tr.names)
        r["ms_per_step_config2"] = step_ms(hp, groups)
        res[name] = r
        print(name, json.dumps(r), file=sys.stderr, flush=True)
    O.set_tf32_1x([])
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
