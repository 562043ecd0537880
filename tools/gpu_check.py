"""Developer check run on the GPU box: component-by-component comparison of the CUDA path with the
oracle, printing max errors (not a pytest; the pytest -m gpu suite asserts the same things)."""
import json
import os
import sys
import time
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import satk_path  # noqa: E402

satk = satk_path.load()
from importlib import import_module  # noqa: E402

O = import_module("self-attention-tacotron_b200.ops")
L = import_module("self-attention-tacotron_b200.lib")
E = import_module("self-attention-tacotron_b200.engine")
from oracle import model as OR  # noqa: E402

dev = "cuda"
RES = {}


def err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    d = (a - b).abs().max().item()
    s = b.abs().max().item()
    return d, d / max(s, 1e-12)


def report(name, a, b, tol=1e-3):
    d, r = err(a, b)
    ok = r <= tol or d <= tol * 1e-2
    RES[name] = dict(abs=d, rel=r, ok=bool(ok))
    print(f"{'OK ' if ok else 'BAD'} {name:48s} abs={d:.3e} rel={r:.3e}", flush=True)
    return ok


def section(fn):
    print(f"\n==== {fn.__name__}", flush=True)
    try:
        fn()
        torch.cuda.synchronize()
    except Exception:
        traceback.print_exc()
        RES[fn.__name__ + ".exception"] = dict(ok=False)
        try:
            torch.cuda.synchronize()
        except Exception:
            traceback.print_exc()


def t_info():
    print(L.device_info(), flush=True)
    out = (L.C.c_int * 5)()
    L.load().satk_struct_sizes(out)
    import ctypes
    exp = [ctypes.sizeof(x) for x in (L.GemmDesc, L.LstmFwdDesc, L.LstmBwdDesc, L.AttnRnnFwdDesc, L.AttnRnnBwdDesc)]
    print("struct sizes C:", list(out), "py:", exp, flush=True)
    assert list(out) == exp


def t_gemm():
    g = torch.Generator(device="cpu").manual_seed(0)
    M, N, K = 150, 70, 45
    A = torch.randn(M, K, generator=g).to(dev)
    Bm = torch.randn(K, N, generator=g).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    C = torch.empty(M, N, device=dev)
    O.gemm(A, Bm, C, M, N, K, lda=K, ldb=N, ldc=N, bias=bias, act="relu", engine=1)
    report("gemm.plain+bias+relu", C, torch.relu(A @ Bm + bias), 1e-5)
    At = A.t().contiguous()
    O.gemm(At, Bm, C, M, N, K, lda=M, ldb=N, ldc=N, transA=True, engine=1)
    report("gemm.transA", C, A @ Bm, 1e-5)
    Bt = Bm.t().contiguous()
    O.gemm(A, Bt, C, M, N, K, lda=K, ldb=K, ldc=N, transB=True, alpha=0.5, engine=1)
    report("gemm.transB.alpha", C, 0.5 * (A @ Bm), 1e-5)
    C2 = torch.randn(M, N, generator=g).to(dev)
    ref = C2 + A @ Bm
    O.gemm(A, Bm, C2, M, N, K, lda=K, ldb=N, ldc=N, split_k=4, engine=1)
    report("gemm.splitk.acc", C2, ref, 1e-5)
    # time-major conv: x [T,B,Cin], W [k,Cin,Cout]
    T, Bb, Cin, Cout, k = 13, 3, 20, 24, 4
    x = torch.randn(T, Bb, Cin, generator=g).to(dev)
    W = torch.randn(k, Cin, Cout, generator=g).to(dev)
    y = torch.empty(T * Bb, Cout, device=dev)
    pl = (k - 1) // 2
    O.gemm(x, W, y, T * Bb, Cout, Cin, lda=Cin, ldb=Cout, ldc=Cout, taps=k, shift0=-pl * Bb, tap_dir=Bb, sBtap=Cin * Cout, engine=1)
    ref = OR.conv1d_same(x.transpose(0, 1).cpu(), W.cpu()).transpose(0, 1).reshape(T * Bb, Cout)
    report("gemm.conv_fwd", y, ref, 1e-5)
    dy = torch.randn(T * Bb, Cout, generator=g).to(dev)
    xr = x.cpu().transpose(0, 1).clone().requires_grad_(True)
    Wr = W.cpu().clone().requires_grad_(True)
    (OR.conv1d_same(xr, Wr).transpose(0, 1).reshape(T * Bb, Cout) * dy.cpu()).sum().backward()
    dW = torch.zeros_like(W)
    O.gemm(x, dy, dW, Cin, Cout, T * Bb, lda=Cin, ldb=Cout, ldc=Cout, transA=True, batch1=k, sC=(Cin * Cout, 0),
           shift0=-pl * Bb, shift_per_batch1=Bb, split_k=2, beta=1.0, engine=1)
    report("gemm.conv_dW", dW, Wr.grad, 1e-5)
    dx = torch.empty(T * Bb, Cin, device=dev)
    O.gemm(dy, W, dx, T * Bb, Cin, Cout, lda=Cout, ldb=Cout, ldc=Cin, transB=True, taps=k, shift0=pl * Bb, tap_dir=-Bb,
           sBtap=Cin * Cout, engine=1)
    report("gemm.conv_dx", dx, xr.grad.transpose(0, 1).reshape(T * Bb, Cin), 1e-5)


def t_bn():
    g = torch.Generator(device="cpu").manual_seed(1)
    T, Bb, Cc = 11, 3, 40
    R = T * Bb
    x = (torch.randn(R, Cc, generator=g) * 2 + 1).to(dev)
    gamma, beta = torch.rand(Cc, generator=g).to(dev) + 0.5, torch.randn(Cc, generator=g).to(dev)
    mean, var = torch.empty(Cc, device=dev), torch.empty(Cc, device=dev)
    mm, mv = torch.zeros(Cc, device=dev), torch.ones(Cc, device=dev)
    O.bn_stats(x, R, Cc, mean, var, mov_mean=mm, mov_var=mv)
    report("bn.mean", mean, x.mean(0), 1e-5)
    report("bn.var", var, x.var(0, unbiased=False), 1e-5)
    report("bn.mov_var", mv, 0.99 + 0.01 * x.var(0, unbiased=True), 1e-5)
    y = torch.empty(R, Cc, device=dev)
    O.bn_apply(x, R, Cc, mean, var, gamma, beta, y, act="relu", maxpool_seq_len=T, pos_stride=Bb)
    xr = x.cpu().clone().requires_grad_(True)
    gr, br = gamma.cpu().clone().requires_grad_(True), beta.cpu().clone().requires_grad_(True)
    z = OR.batch_norm(xr.view(T, Bb, Cc).transpose(0, 1), gr, br, None, None, True)
    yr = OR.maxpool2_same(torch.relu(z)).transpose(0, 1).reshape(R, Cc)
    report("bn.apply.relu.maxpool", y, yr, 1e-5)
    dy = torch.randn(R, Cc, generator=g)
    (yr * dy).sum().backward()
    dx = torch.empty(R, Cc, device=dev)
    dg, db = torch.zeros(Cc, device=dev), torch.zeros(Cc, device=dev)
    O.bn_bwd(x, R, Cc, mean, var, gamma, beta, dy.to(dev), dx, dg, db, torch.empty(2 * Cc, device=dev), act="relu",
             maxpool_seq_len=T, pos_stride=Bb)
    report("bn.bwd.dx", dx, xr.grad, 1e-4)
    report("bn.bwd.dgamma", dg, gr.grad, 1e-4)
    report("bn.bwd.dbeta", db, br.grad, 1e-4)


def t_softmax_loss_adam():
    g = torch.Generator(device="cpu").manual_seed(2)
    nm, T = 6, 37
    S = torch.randn(nm, T, T, generator=g)
    mask = (torch.rand(nm, T, T, generator=g) < 0.9).to(torch.uint8)
    Sd = S.clone().to(dev)
    Pd = torch.empty(nm, T, T, device=dev)
    O.softmax_fwd(Sd, nm, T, True, mask.to(dev), 1 / 0.9, Pd)
    Sr = S.clone().requires_grad_(True)
    tri = torch.ones(T, T, dtype=torch.bool).tril()
    Pr = torch.softmax(torch.where(tri, Sr, torch.full_like(Sr, -float("inf"))), -1)
    Pdr = Pr * mask / 0.9
    report("softmax.P", Sd, Pr, 1e-5)
    report("softmax.Pd", Pd, Pdr, 1e-5)
    dP = torch.randn(nm, T, T, generator=g)
    (Pdr * dP).sum().backward()
    dS = torch.empty(nm, T, T, device=dev)
    O.softmax_bwd(Sd, dP.to(dev), nm, T, True, dS, mask.to(dev), 1 / 0.9)
    report("softmax.bwd", dS, Sr.grad, 1e-5)
    # adam
    n = 1003
    p, gr = torch.randn(n, generator=g), torch.randn(n, generator=g) * 3
    m, v = torch.zeros(n), torch.zeros(n)
    pd, gd, md, vd = p.to(dev), gr.to(dev), m.to(dev), v.to(dev)
    ss = torch.zeros(O.SUMSQ_SCRATCH, device=dev)
    O.grad_sumsq(gd, ss)
    O.adam_clip(pd, gd, md, vd, ss, 0.5, 1.0, 1e-3, 0.9, 0.999, 1e-8, 3)
    cl, norm = OR.clip_by_global_norm([gr * 0.5], 1.0)
    OR.adam_update(p, cl[0], m, v, 1e-3, 3, 0.9, 0.999, 1e-8)
    report("adam.p", pd, p, 1e-5)
    report("adam.v", vd, v, 1e-5)


def t_lstm():
    g = torch.Generator(device="cpu").manual_seed(3)
    for H, Bb, T, rev in ((128, 5, 9, False), (128, 5, 9, True), (256, 3, 7, False)):
        W = torch.randn(2 * H, 4 * H, generator=g) * 0.08
        b = torch.randn(4 * H, generator=g) * 0.1
        x = torch.randn(Bb, T, H, generator=g)
        lens = torch.randint(2, T + 1, (Bb,), generator=g)
        lens[0] = T
        mc = (torch.rand(T, Bb, H, generator=g) < 0.9).to(torch.uint8)
        mh = (torch.rand(T, Bb, H, generator=g) < 0.9).to(torch.uint8)
        Wr, br, xr = W.clone().requires_grad_(True), b.clone().requires_grad_(True), x.clone().requires_grad_(True)
        yr = OR.zoneout_lstm_sequence(xr, lens, Wr, br, mc, mh, 0.1, 0.1, True, reverse=rev)      # [B,T,H]
        dy = torch.randn(Bb, T, H, generator=g)
        (yr * dy).sum().backward()
        # CUDA: xg time-major
        Wd = W.to(dev)
        x_tm = x.transpose(0, 1).contiguous().to(dev)
        xg = torch.empty(T * Bb, 4 * H, device=dev)
        O.linear(x_tm, Wd[:H], xg, bias=b.to(dev))
        out = torch.full((T, Bb, H), 7.0, device=dev)
        gates, cp, hp = torch.empty(T * Bb, 4 * H, device=dev), torch.empty(T * Bb, H, device=dev), torch.empty(T * Bb, H, device=dev)
        O.lstm_seq_fwd(xg, Wd[H:], out, T, Bb, H, reverse=rev, lengths=lens.to(dev), mask_c=mc.to(dev), mask_h=mh.to(dev),
                       gates=gates, c_prev=cp, h_prev=hp)
        tag = f"lstm.H{H}.rev{int(rev)}"
        report(tag + ".out", out.transpose(0, 1), yr, 1e-4)
        dg = torch.empty(T * Bb, 4 * H, device=dev)
        O.lstm_seq_bwd(Wd[H:], gates, cp, dy.transpose(0, 1).contiguous().to(dev), dg, T, Bb, H, reverse=rev, lengths=lens.to(dev),
                       mask_c=mc.to(dev), mask_h=mh.to(dev))
        dW = torch.zeros(2 * H, 4 * H, device=dev)
        O.linear_dw(x_tm, dg, dW, T * Bb, H, 4 * H)
        O.linear_dw(hp, dg, dW, T * Bb, H, 4 * H, w_off=H * 4 * H)
        report(tag + ".dW", dW, Wr.grad, 1e-3)
        report(tag + ".db", dg.sum(0), br.grad, 1e-3)
        dx = torch.empty(T * Bb, H, device=dev)
        O.linear_dx(dg, Wd[:H], dx, T * Bb)
        report(tag + ".dx", dx.view(T, Bb, H).transpose(0, 1), xr.grad, 1e-3)


def model_case(cfg, B, Tt, Tm, training, tag, check_grads=True, overrides=None):
    hp = satk.load_hparams(os.path.join(ROOT, "examples", cfg), overrides)
    d = satk.dims_from_hparams(hp)
    ps = satk.ParamStore(d).init(7, "random")
    f, l = satk.synthetic_batch(hp, B, Tt, Tm, seed=11)
    masks = satk.make_masks(d, B, Tt, Tm // d.r, seed=5) if training else None
    P = {k: v.clone() for k, v in ps.as_dict().items()}
    tr = OR.OracleTrainer(d, hp, P)
    ref, grads, stats = tr.loss_and_grads(f, l, masks, training)
    eng = E.TacotronEngine(hp, dev, params=ps)
    fd = satk.SourceData(*[x.to(dev) if torch.is_tensor(x) else x for x in f])
    ld = satk.MelData(*[x.to(dev) if torch.is_tensor(x) else x for x in l])
    md = {k: v.to(dev) for k, v in masks.items()} if masks else None
    out = eng.forward(fd, ld, training, md)
    torch.cuda.synchronize()
    Td = Tm // d.r
    report(tag + ".memory1", out["memory1_tm"].view(Tt, B, -1).transpose(0, 1), ref["memory1"])
    if d.dual:
        report(tag + ".memory2", out["memory2_tm"].view(Tt, B, -1).transpose(0, 1), ref["memory2"])
        report(tag + ".enc_self_align0", out["enc_self_P"][0].transpose(1, 2), ref["enc_self_alignments"][0])
    report(tag + ".alignment", out["align1_tm"].permute(1, 2, 0), ref["alignment"])
    if d.dual:
        report(tag + ".alignment2", out["align2_tm"].permute(1, 2, 0), ref["alignment2"])
    report(tag + ".mel", out["mel_tm"].view(Td, B, d.r, d.n_mels).permute(1, 0, 2, 3).reshape(B, Tm, d.n_mels), ref["mel"])
    report(tag + ".stop", out["stop_tm"].view(Td, B).t(), ref["stop"].squeeze(-1))
    report(tag + ".loss", out["losses"], torch.stack([ref["mel_loss"], ref["done_loss"], ref["loss"]]))
    if check_grads:
        eng.backward()
        torch.cuda.synchronize()
        bad = 0
        worst = []
        for n in tr.names:
            dd, rr = err(eng.ps.g[n], grads[n])
            gs = grads[n].abs().max().item()
            ok = rr <= 2e-3 or dd <= 2e-6
            RES[f"{tag}.grad.{n}"] = dict(abs=dd, rel=rr, ok=bool(ok))
            worst.append((rr, n, dd, gs))
            bad += (not ok)
        worst.sort(reverse=True)
        print(f"  grads: {len(tr.names) - bad}/{len(tr.names)} ok; worst:", flush=True)
        for rr, n, dd, gs in worst[:12]:
            print(f"     {n:32s} rel={rr:.3e} abs={dd:.3e} |g|max={gs:.3e}", flush=True)
    return eng, fd, ld, md


def t_model_dual_eval_small():
    model_case("ljspeech_self-attention-tacotron.json", 3, 20, 24, False, "dual.eval.small")


def t_model_dual_train_small():
    model_case("ljspeech_self-attention-tacotron.json", 5, 23, 28, True, "dual.train.small")


def t_model_single_train_small():
    model_case("ljspeech_tacotron.json", 2, 21, 20, True, "single.train.small")


def t_model_dual_train_medium():
    model_case("ljspeech_self-attention-tacotron.json", 8, 70, 120, True, "dual.train.medium")


def t_model_vctk_small():
    model_case("vctk_self-attention-tacotron.json", 4, 18, 16, True, "vctk.train.small")


def t_variants():
    model_case("ljspeech_self-attention-tacotron.json", 3, 20, 24, True, "dual.locsens", overrides="attention=location_sensitive")
    model_case("ljspeech_tacotron.json", 3, 20, 24, True, "single.additive", overrides="attention=additive")


def t_full_size_timing():
    hp = satk.load_hparams(os.path.join(ROOT, "examples", "ljspeech_self-attention-tacotron.json"))
    eng = E.TacotronEngine(hp, dev, seed=1)
    f, l = satk.synthetic_batch(hp, 32, 148, 800, seed=3, device=dev)
    for i in range(4):
        torch.cuda.synchronize()
        t0 = time.time()
        out = eng.forward(f, l, True)
        torch.cuda.synchronize()
        t1 = time.time()
        eng.backward()
        torch.cuda.synchronize()
        t2 = time.time()
        eng.optimizer_step()
        torch.cuda.synchronize()
        t3 = time.time()
        print(f"  step {i}: fwd {1e3 * (t1 - t0):.2f} ms  bwd {1e3 * (t2 - t1):.2f} ms  opt {1e3 * (t3 - t2):.2f} ms  "
              f"loss {out['losses'].tolist()}", flush=True)
    RES["full.step_ms"] = dict(ok=True, fwd=1e3 * (t1 - t0), bwd=1e3 * (t2 - t1), opt=1e3 * (t3 - t2))


if __name__ == "__main__":
    which = sys.argv[1:] or None
    allt = [t_info, t_gemm, t_bn, t_softmax_loss_adam, t_lstm, t_model_dual_eval_small, t_model_dual_train_small,
            t_model_single_train_small, t_model_vctk_small, t_variants, t_model_dual_train_medium, t_full_size_timing]
    for fn in allt:
        if which is None or fn.__name__ in which:
            section(fn)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(RES, open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w"), indent=1)
    nbad = sum(1 for v in RES.values() if not v["ok"])
    print(f"\nSUMMARY: {len(RES) - nbad}/{len(RES)} ok, {nbad} bad", flush=True)
