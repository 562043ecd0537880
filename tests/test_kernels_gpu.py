"""GPU unit tests of individual C-ABI entry points against plain torch / oracle primitives."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import model as OR  # noqa: E402


def _O():
    from importlib import import_module
    return import_module("self-attention-tacotron_b200.ops")


def _close(a, b, tol, what):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    d, s = (a - b).abs().max().item(), b.abs().max().item()
    assert d <= tol * max(s, 1e-6), f"{what}: {d:.3e} vs {s:.3e}"


@pytest.mark.parametrize("engine", [1, 0])
def test_gemm_variants(engine):
    O = _O()
    g = torch.Generator().manual_seed(0)
    for (M, N, K) in ((150, 70, 45), (256, 256, 128), (1, 1, 7), (65, 129, 257)):
        A, B = torch.randn(M, K, generator=g).cuda(), torch.randn(K, N, generator=g).cuda()
        bias = torch.randn(N, generator=g).cuda()
        C = torch.empty(M, N, device="cuda")
        O.gemm(A, B, C, M, N, K, lda=K, ldb=N, ldc=N, bias=bias, act="tanh", engine=engine)
        _close(C, torch.tanh(A @ B + bias), 2e-5, f"gemm {M}x{N}x{K}")
        O.gemm(A.t().contiguous(), B.t().contiguous(), C, M, N, K, lda=M, ldb=K, ldc=N, transA=True, transB=True, alpha=0.25, engine=engine)
        _close(C, 0.25 * (A @ B), 2e-5, "gemm transA transB")
        C2 = torch.randn(M, N, generator=g).cuda()
        ref = C2 + A @ B
        O.gemm(A, B, C2, M, N, K, lda=K, ldb=N, ldc=N, split_k=3, engine=engine)
        _close(C2, ref, 2e-5, "gemm split-k accumulate")


def test_gemm_tensor_core_tile():
    """tcgen05 3xTF32 tile (engine 2) against fp64: plain, epilogue, split-K and tapped (conv) products.
    Tolerance 1e-7 * K relative: the tensor core accumulates with truncation (measured ~6e-8 * K), not IEEE fp32."""
    O = _O()
    g = torch.Generator().manual_seed(7)
    for (M, N, K) in ((128, 128, 32), (300, 200, 160), (4736, 224, 256), (1000, 1024, 544), (130, 48, 36)):
        A, Bt = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
        bias, res = torch.randn(N, generator=g), torch.randn(M, N, generator=g)
        ref = (torch.tanh(A.double() @ Bt.double().t() + bias.double()) + res.double()).float()
        C = torch.empty(M, N, device="cuda")
        O.gemm(A.cuda(), Bt.cuda(), C, M, N, K, lda=K, ldb=K, ldc=N, transB=True, bias=bias.cuda(), act="tanh", residual=res.cuda(),
               ldres=N, engine=2)
        _close(C, ref, max(1e-5, 1e-7 * K), f"tc gemm {M}x{N}x{K}")
        C2 = torch.randn(M, N, generator=g).cuda()
        ref2 = (C2.cpu().double() + 0.5 * (A.double() @ Bt.double().t())).float()
        O.gemm(A.cuda(), Bt.cuda(), C2, M, N, K, lda=K, ldb=K, ldc=N, transB=True, alpha=0.5, split_k=2, engine=2)
        _close(C2, ref2, max(1e-5, 1e-7 * K), f"tc gemm split-k {M}x{N}x{K}")
    # time-major conv: x [T,B,Cin], transposed taps Wt [k,Cout,Cin]
    T, Bb, Cin, Cout, k = 40, 4, 128, 128, 5
    x, W = torch.randn(T, Bb, Cin, generator=g), torch.randn(k, Cin, Cout, generator=g) * 0.1
    ref = OR.conv1d_same(x.transpose(0, 1).double(), W.double()).transpose(0, 1).reshape(T * Bb, Cout).float()
    Wt = W.transpose(1, 2).contiguous().cuda()
    y = torch.empty(T * Bb, Cout, device="cuda")
    pl = (k - 1) // 2
    O.gemm(x.cuda(), Wt, y, T * Bb, Cout, Cin, lda=Cin, ldb=Cin, ldc=Cout, transB=True, taps=k, shift0=-pl * Bb, tap_dir=Bb,
           sBtap=Cin * Cout, engine=2)
    _close(y, ref, 1e-7 * Cin * k, "tc conv")


def test_weight_gradient_products_tensor_core_path():
    """dW += X^T dY through the transposed-operand tcgen05 path (rows >= DW_TC_MIN_ROWS): plain, strided with a delayed X
    (shift0 < 0, the LSTM-1 context rows), shared transposed dY, ragged row count; against fp64."""
    O = _O()
    g = torch.Generator().manual_seed(11)
    for (rows, K, N, ldx, x_off, shift0) in ((2304, 96, 80, None, 0, 0), (4096, 288, 1024, 544, 256, -32), (2051, 64, 48, None, 0, -3)):
        x = torch.randn(rows, ldx or K, generator=g)
        dy = torch.randn(rows, N, generator=g)
        dW0 = torch.randn(K, N, generator=g)
        xs = x[:, x_off:x_off + K].double()
        if shift0 < 0:
            xs = torch.cat([torch.zeros(-shift0, K, dtype=torch.float64), xs[:shift0]], 0)
        ref = (dW0.double() + xs.t() @ dy.double()).float()
        dW = dW0.clone().cuda()
        O.linear_dw(x.cuda(), dy.cuda(), dW, rows, K, N, ldx=ldx, x_off=x_off, shift0=shift0)
        _close(dW, ref, 1e-7 * rows, f"dW tc {rows}x{K}x{N}")
        dW = dW0.clone().cuda()
        yT = O.transposed_rows(dy.cuda(), rows, N)
        O.linear_dw(x.cuda(), dy.cuda(), dW, rows, K, N, ldx=ldx, x_off=x_off, shift0=shift0, yT=yT)
        _close(dW, ref, 1e-7 * rows, f"dW tc shared yT {rows}x{K}x{N}")


def test_conv_weight_gradient_all_taps_one_tensor_core_launch():
    """dW[tap] += x_shifted(tap)^T draw for every tap of a SAME conv as z-batches of one tcgen05 launch (K-shifted TMA
    coordinate, TMA reduce-add epilogue) against fp64 autograd of the oracle conv; shared transposes for a bank slice."""
    O = _O()
    g = torch.Generator().manual_seed(5)
    T, Bb = 80, 32
    R = T * Bb
    for (k, cin, cout, ld_draw, off) in ((1, 128, 128, 128, 0), (4, 128, 128, 384, 256), (3, 256, 128, 128, 0), (16, 128, 64, 64, 0)):
        x = torch.randn(T, Bb, cin, generator=g)
        draw_full = torch.randn(R, ld_draw, generator=g)
        draw = draw_full[:, off:off + cout]
        W = torch.randn(k, cin, cout, generator=g).double().requires_grad_(True)
        y = OR.conv1d_same(x.transpose(0, 1).double(), W).transpose(0, 1).reshape(R, cout)
        (y * draw.double()).sum().backward()
        dW0 = torch.randn(k, cin, cout, generator=g)
        dW = dW0.clone().cuda()
        xT = O.transposed_rows(x.cuda(), R, cin)
        drawT = O.transposed_rows(draw_full.cuda(), R, ld_draw)
        O.conv_dw_tc(xT, drawT, dW, R, cin, cout, k, Bb, drawT_row0=off)
        _close(dW, (dW0.double() + W.grad).float(), 1e-7 * R, f"conv dW tc k={k} cin={cin} cout={cout}")
        # the same straight from x and draw (MN-major operands, no transposed copies)
        dW = dW0.clone().cuda()
        O.conv_dw_mn(x.cuda(), draw_full.cuda(), dW, R, cin, cout, k, Bb, draw_ld=ld_draw, draw_off=off)
        _close(dW, (dW0.double() + W.grad).float(), 1e-7 * R, f"conv dW tc (MN-major) k={k} cin={cin} cout={cout}")


def test_weight_gradient_products_mn_major_operands():
    """dW += X^T dY with both operands read row-major as they are (the reduction index is the row): MN-major shared-memory
    descriptors of the tcgen05 tile.  Ragged M / N / row counts, strided operands with offsets, a delayed and an advanced X; against
    fp64, and against the transposed-operand path."""
    O = _O()
    assert O.DW_MN
    g = torch.Generator().manual_seed(12)
    for (rows, K, N, ldx, x_off, ldy, y_off, shift0) in ((2304, 96, 80, None, 0, None, 0, 0), (4096, 288, 1024, 544, 256, None, 0, -32),
                                                         (2051, 64, 48, None, 0, 112, 60, -3), (2500, 200, 130, 260, 8, None, 0, 5),
                                                         (12800, 544, 1024, None, 0, None, 0, 0)):
        x = torch.randn(rows, ldx or K, generator=g)
        dy = torch.randn(rows, ldy or N, generator=g)
        dW0 = torch.randn(K, N, generator=g)
        xs = x[:, x_off:x_off + K].double()
        if shift0 < 0:
            xs = torch.cat([torch.zeros(-shift0, K, dtype=torch.float64), xs[:shift0]], 0)
        elif shift0 > 0:
            xs = torch.cat([xs[shift0:], torch.zeros(shift0, K, dtype=torch.float64)], 0)
        ref = (dW0.double() + xs.t() @ dy[:, y_off:y_off + N].double()).float()
        dW = dW0.clone().cuda()
        O.linear_dw(x.cuda(), dy.cuda(), dW, rows, K, N, ldx=ldx, x_off=x_off, ldy=ldy, y_off=y_off, shift0=shift0)
        _close(dW, ref, 1e-7 * rows, f"dW MN-major {rows}x{K}x{N}")


def test_conv_bank_one_launch_forward_and_input_gradient():
    """All widths of a SAME-padded conv bank as z-batches of one tcgen05 launch (satk_gemm_desc.bank_widths): forward into the
    concatenated output and the summed input gradient (TMA reduce-add onto the residual), against fp64 convs."""
    O = _O()
    g = torch.Generator().manual_seed(9)
    T, Bb, cin, C, W = 37, 8, 128, 128, 6
    R = T * Bb
    x = torch.randn(T, Bb, cin, generator=g)
    Ws = [torch.randn(k, cin, C, generator=g) / (k * cin) ** 0.5 for k in range(1, W + 1)]
    # forward: kernels stored back to back in the K-contiguous layout [tap][C][cin]
    Wt_all = torch.cat([w.transpose(1, 2).contiguous().reshape(-1) for w in Ws]).cuda()
    raw = torch.full((R, W * C + 4), 7.0, device="cuda")
    O.gemm(x.cuda(), Wt_all, raw, R, C, cin, lda=cin, ldb=cin, ldc=W * C + 4, transB=True, tap_dir=Bb, sBtap=cin * C, bank_widths=W,
           bank_c_nstep=C, engine=2)
    for k in range(1, W + 1):
        ref = OR.conv1d_same(x.transpose(0, 1).double(), Ws[k - 1].double()).transpose(0, 1).reshape(R, C).float()
        _close(raw[:, (k - 1) * C:k * C], ref, 1e-6 * cin * k, f"bank forward width {k}")
    assert (raw[:, W * C:] == 7.0).all()
    # input gradient: dx = res + sum_k conv_k^T(draw_k)
    draw = torch.randn(R, W * C, generator=g)
    res = torch.randn(R, cin, generator=g)
    xr = x.transpose(0, 1).double().clone().requires_grad_(True)
    tot = sum((OR.conv1d_same(xr, Ws[k - 1].double()).transpose(0, 1).reshape(R, C) * draw[:, (k - 1) * C:k * C].double()).sum()
              for k in range(1, W + 1))
    tot.backward()
    W_all = torch.cat([w.reshape(-1) for w in Ws]).cuda()
    dx = res.clone().cuda()
    O.gemm(draw.cuda(), W_all, dx, R, cin, C, lda=W * C, ldb=C, ldc=cin, transB=True, tap_dir=-Bb, sBtap=cin * C, beta=1.0,
           bank_widths=W, bank_a_kstep=C, engine=2)
    _close(dx, (res.double() + xr.grad.transpose(0, 1).reshape(R, cin)).float(), 2e-6 * C * W, "bank input gradient")


def test_gemm_time_major_conv_and_grads():
    O = _O()
    g = torch.Generator().manual_seed(1)
    for k in (1, 2, 3, 10, 16):
        T, Bb, Cin, Cout = 19, 3, 20, 24
        x = torch.randn(T, Bb, Cin, generator=g)
        W = torch.randn(k, Cin, Cout, generator=g)
        xr, Wr = x.transpose(0, 1).clone().requires_grad_(True), W.clone().requires_grad_(True)
        yr = OR.conv1d_same(xr, Wr).transpose(0, 1).reshape(T * Bb, Cout)
        dy = torch.randn(T * Bb, Cout, generator=g)
        (yr * dy).sum().backward()
        pl = (k - 1) // 2
        xd, Wd, dyd = x.cuda(), W.cuda(), dy.cuda()
        y = torch.empty(T * Bb, Cout, device="cuda")
        O.gemm(xd, Wd, y, T * Bb, Cout, Cin, lda=Cin, ldb=Cout, ldc=Cout, taps=k, shift0=-pl * Bb, tap_dir=Bb, sBtap=Cin * Cout)
        _close(y, yr, 1e-5, f"conv k={k}")
        dW = torch.zeros_like(Wd)
        O.gemm(xd, dyd, dW, Cin, Cout, T * Bb, lda=Cin, ldb=Cout, ldc=Cout, transA=True, batch1=k, sC=(Cin * Cout, 0),
               shift0=-pl * Bb, shift_per_batch1=Bb, split_k=2, beta=1.0)
        _close(dW, Wr.grad, 1e-5, f"conv dW k={k}")
        dx = torch.empty(T * Bb, Cin, device="cuda")
        O.gemm(dyd, Wd, dx, T * Bb, Cin, Cout, lda=Cout, ldb=Cout, ldc=Cin, transB=True, taps=k, shift0=pl * Bb, tap_dir=-Bb,
               sBtap=Cin * Cout)
        _close(dx, xr.grad.transpose(0, 1).reshape(T * Bb, Cin), 1e-5, f"conv dx k={k}")


def test_batch_norm_relu_maxpool_fwd_bwd():
    O = _O()
    g = torch.Generator().manual_seed(2)
    T, Bb, Cc = 11, 3, 40
    R = T * Bb
    x = (torch.randn(R, Cc, generator=g) * 2 + 1)
    gamma, beta = torch.rand(Cc, generator=g) + 0.5, torch.randn(Cc, generator=g)
    xd = x.cuda()
    mean, var = torch.empty(Cc, device="cuda"), torch.empty(Cc, device="cuda")
    mm, mv = torch.zeros(Cc, device="cuda"), torch.ones(Cc, device="cuda")
    O.bn_stats(xd, R, Cc, mean, var, mov_mean=mm, mov_var=mv)
    _close(mean, x.mean(0), 1e-5, "mean")
    _close(var, x.var(0, unbiased=False), 1e-5, "var")
    _close(mv, 0.99 + 0.01 * x.var(0, unbiased=True), 1e-5, "moving var (Bessel)")
    y = torch.empty(R, Cc, device="cuda")
    O.bn_apply(xd, R, Cc, mean, var, gamma.cuda(), beta.cuda(), y, act="relu", maxpool_seq_len=T, pos_stride=Bb)
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    z = OR.batch_norm(xr.view(T, Bb, Cc).transpose(0, 1), gr, br, None, None, True)
    yr = OR.maxpool2_same(torch.relu(z)).transpose(0, 1).reshape(R, Cc)
    _close(y, yr, 1e-5, "bn+relu+maxpool")
    dy = torch.randn(R, Cc, generator=g)
    (yr * dy).sum().backward()
    dx, dg, db = torch.empty(R, Cc, device="cuda"), torch.zeros(Cc, device="cuda"), torch.zeros(Cc, device="cuda")
    O.bn_bwd(xd, R, Cc, mean, var, gamma.cuda(), beta.cuda(), dy.cuda(), dx, dg, db, torch.empty(2 * Cc, device="cuda"), act="relu",
             maxpool_seq_len=T, pos_stride=Bb)
    _close(dx, xr.grad, 1e-4, "bn dx")
    _close(dg, gr.grad, 1e-4, "bn dgamma")
    _close(db, br.grad, 1e-4, "bn dbeta")


def test_softmax_causal_dropout_fwd_bwd():
    O = _O()
    g = torch.Generator().manual_seed(3)
    nm, T = 6, 37
    S = torch.randn(nm, T, T, generator=g)
    mask = (torch.rand(nm, T, T, generator=g) < 0.9).to(torch.uint8)
    Sd, Pd = S.clone().cuda(), torch.empty(nm, T, T, device="cuda")
    O.softmax_fwd(Sd, nm, T, True, mask.cuda(), 1 / 0.9, Pd)
    Sr = S.clone().requires_grad_(True)
    tri = torch.ones(T, T, dtype=torch.bool).tril()
    Pr = torch.softmax(torch.where(tri, Sr, torch.full_like(Sr, -float("inf"))), -1)
    Pdr = Pr * mask / 0.9
    _close(Sd, Pr, 1e-5, "P")
    _close(Pd, Pdr, 1e-5, "Pd")
    dP = torch.randn(nm, T, T, generator=g)
    (Pdr * dP).sum().backward()
    dS = torch.empty(nm, T, T, device="cuda")
    O.softmax_bwd(Sd, dP.cuda(), nm, T, True, dS, mask.cuda(), 1 / 0.9)
    _close(dS, Sr.grad, 1e-5, "dS")


@pytest.mark.parametrize("H,Bb,T,rev,ragged", [(128, 5, 9, False, True), (128, 5, 9, True, True), (256, 3, 7, False, True),
                                               (256, 9, 12, True, True),
                                               # no per-row lengths: the decoder layers' 5-rows-per-cluster kernels (lstm_seq5.cu),
                                               # one partial cluster / two clusters with a ragged last one / 7 clusters
                                               (256, 3, 7, False, False), (256, 9, 12, False, False), (256, 32, 21, False, False)])
def test_zoneout_lstm_sequence(H, Bb, T, rev, ragged):
    O = _O()
    g = torch.Generator().manual_seed(4)
    W, b = torch.randn(2 * H, 4 * H, generator=g) * 0.08, torch.randn(4 * H, generator=g) * 0.1
    x = torch.randn(Bb, T, H, generator=g)
    lens = torch.randint(1, T + 1, (Bb,), generator=g)
    lens[0] = T
    if not ragged:
        lens[:] = T
    dev_lens = lens.cuda() if ragged else None
    mc = (torch.rand(T, Bb, H, generator=g) < 0.9).to(torch.uint8)
    mh = (torch.rand(T, Bb, H, generator=g) < 0.9).to(torch.uint8)
    Wr, br, xr = W.clone().requires_grad_(True), b.clone().requires_grad_(True), x.clone().requires_grad_(True)
    yr = OR.zoneout_lstm_sequence(xr, lens, Wr, br, mc, mh, 0.1, 0.1, True, reverse=rev)
    dy = torch.randn(Bb, T, H, generator=g)
    (yr * dy).sum().backward()
    Wd = W.cuda()
    x_tm = x.transpose(0, 1).contiguous().cuda()
    xg = torch.empty(T * Bb, 4 * H, device="cuda")
    O.linear(x_tm, Wd[:H], xg, bias=b.cuda())
    out = torch.full((T, Bb, H), 7.0, device="cuda")
    gates, cp, hp = (torch.empty(T * Bb, 4 * H, device="cuda"), torch.empty(T * Bb, H, device="cuda"), torch.empty(T * Bb, H, device="cuda"))
    O.lstm_seq_fwd(xg, Wd[H:], out, T, Bb, H, reverse=rev, lengths=dev_lens, mask_c=mc.cuda(), mask_h=mh.cuda(), gates=gates,
                   c_prev=cp, h_prev=hp)
    _close(out.transpose(0, 1), yr, 1e-4, "lstm out")
    dg = torch.empty(T * Bb, 4 * H, device="cuda")
    O.lstm_seq_bwd(Wd[H:], gates, cp, dy.transpose(0, 1).contiguous().cuda(), dg, T, Bb, H, reverse=rev, lengths=dev_lens,
                   mask_c=mc.cuda(), mask_h=mh.cuda())
    dW = torch.zeros(2 * H, 4 * H, device="cuda")
    O.linear_dw(x_tm, dg, dW, T * Bb, H, 4 * H)
    O.linear_dw(hp, dg, dW, T * Bb, H, 4 * H, w_off=H * 4 * H)
    _close(dW, Wr.grad, 1e-3, "lstm dW")
    _close(dg.sum(0), br.grad, 1e-3, "lstm db")
    dx = torch.empty(T * Bb, H, device="cuda")
    O.linear_dx(dg, Wd[:H], dx, T * Bb)
    _close(dx.view(T, Bb, H).transpose(0, 1), xr.grad, 1e-3, "lstm dx")
    # eval-mode interpolation (no masks)
    yr2 = OR.zoneout_lstm_sequence(x, lens, W, b, None, None, 0.1, 0.1, False, reverse=rev)
    O.lstm_seq_fwd(xg, Wd[H:], out, T, Bb, H, reverse=rev, lengths=dev_lens, zc=0.1, zh=0.1)
    _close(out.transpose(0, 1), yr2, 1e-4, "lstm eval out")


def test_adam_clip_and_losses():
    O = _O()
    g = torch.Generator().manual_seed(5)
    n = 1003
    p, gr = torch.randn(n, generator=g), torch.randn(n, generator=g) * 3
    m, v = torch.zeros(n), torch.zeros(n)
    pd, gd, md, vd, ss = p.cuda(), gr.cuda(), m.cuda(), v.cuda(), torch.zeros(O.SUMSQ_SCRATCH, device="cuda")
    O.grad_sumsq(gd, ss)
    O.adam_clip(pd, gd, md, vd, ss, 0.5, 1.0, 1e-3, 0.9, 0.999, 1e-8, 3)
    cl, _ = OR.clip_by_global_norm([gr * 0.5], 1.0)
    OR.adam_update(p, cl[0], m, v, 1e-3, 3, 0.9, 0.999, 1e-8)
    _close(pd, p, 1e-6, "adam p")
    _close(md, m, 1e-5, "adam m")
    # losses: time-major predictions vs batch-major targets
    B, Tm, nm, r = 3, 12, 8, 2
    Td = Tm // r
    pred = torch.randn(Td, B, r * nm, generator=g)
    stop = torch.randn(Td, B, generator=g)
    mel = torch.randn(B, Tm, nm, generator=g)
    lens = torch.tensor([12, 8, 6])
    sm = (torch.arange(Tm)[None] < lens[:, None]).float()
    bm = (torch.arange(Td)[None] < (lens // r)[:, None]).float()
    done = (torch.arange(Td)[None] >= (lens // r)[:, None] - 1).float()
    pr, sr = pred.clone().requires_grad_(True), stop.clone().requires_grad_(True)
    melp = pr.view(Td, B, r, nm).permute(1, 0, 2, 3).reshape(B, Tm, nm)
    l1 = OR.spec_loss_l1(melp, mel, sm)
    l2 = OR.binary_loss(sr.t().unsqueeze(-1), done, bm)
    (l1 + l2).backward()
    out3, dp, ds = torch.empty(3, device="cuda"), torch.empty(Td, B, r * nm, device="cuda"), torch.empty(Td, B, device="cuda")
    O.losses(pred.cuda(), stop.cuda(), mel.cuda(), done.cuda(), sm.cuda(), bm.cuda(), B, Tm, nm, r, out3, dp, ds, torch.empty(4, device="cuda"))
    _close(out3, torch.stack([l1, l2, l1 + l2]), 1e-5, "losses")
    _close(dp, pr.grad, 1e-5, "dpred")
    _close(ds, sr.grad, 1e-5, "dstop")


def test_bernoulli_mask_rate():
    O = _O()
    m = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    O.bernoulli_mask(m, 0.9, 12345)
    assert abs(m.float().mean().item() - 0.9) < 2e-3
    m2 = torch.empty_like(m)
    O.bernoulli_mask(m2, 0.9, 12346)
    assert (m != m2).float().mean().item() > 0.1


@pytest.mark.parametrize("T,nz,dh,causal", [(400, 6, 128, True), (200, 4, 64, False), (128, 3, 32, True), (148, 2, 128, False)])
def test_self_attention_products_tensor_core_zbatches(T, nz, dh, causal):
    """QK^T, P.V and the four gradient products of the self-attention block (self_attention.py:45-65) as z-batches of the
    tcgen05 tile over shared 2-D operand views (satk_gemm_desc.zcoord): heads as k-shifts of the time-major activations, stacked
    score matrices as row shifts, operands reduced over their row index read through MN-major descriptors, ragged last tiles
    clipped by the TMA stores, causal tile /
    k-range skipping.  Against fp64; entries the causal softmax never reads are not compared."""
    O = _O()
    g = torch.Generator().manual_seed(T + dh)
    W = nz * dh
    X, Y = torch.randn(T, W, generator=g), torch.randn(T, W, generator=g)
    P = torch.randn(nz, T, T, generator=g)
    if causal:
        P = torch.tril(P)      # what the causal softmax leaves: zeros above the diagonal
    Xz = X.double().view(T, nz, dh).transpose(0, 1)     # [nz, T, dh]
    Yz = Y.double().view(T, nz, dh).transpose(0, 1)
    tol = 3e-6 * max(T, dh) ** 0.5

    S = torch.full((nz, T, T), float("nan"), device="cuda")
    O.attn_scores_tc(X.cuda(), Y.cuda(), S, T, nz, dh, alpha=0.25, causal=causal)
    ref = 0.25 * Xz @ Yz.transpose(1, 2)
    S = S.cpu()
    if causal:
        keep = torch.tril(torch.ones(T, T, dtype=torch.bool))
        S, ref = torch.where(keep, S, torch.zeros(())), torch.where(keep, ref, torch.zeros((), dtype=torch.float64))
    _close(S, ref.float(), tol, "scores")

    Out = torch.full((T, W), float("nan"), device="cuda")
    O.attn_apply_tc(P.cuda(), Y.cuda(), Out, T, nz, dh, alpha=0.5, causal=causal)
    ref = 0.5 * (P.double() @ Yz).transpose(0, 1).reshape(T, W)
    _close(Out, ref.float(), tol, "apply")

    Out = torch.full((T, W), float("nan"), device="cuda")
    O.attn_apply_t_tc(P.cuda(), Y.cuda(), Out, T, nz, dh, alpha=2.0, causal=causal)
    ref = 2.0 * (P.double().transpose(1, 2) @ Yz).transpose(0, 1).reshape(T, W)
    _close(Out, ref.float(), tol, "apply_t")


def test_self_attention_products_zcoord_refused_on_simt_tile():
    O = _O()
    x = torch.zeros(128, 128, device="cuda")
    s = torch.zeros(1, 128, 128, device="cuda")
    with pytest.raises(RuntimeError):
        O.gemm(x, x, s, 128, 128, 128, lda=128, ldb=128, ldc=128, transB=True, batch1=1, sC=(128 * 128, 0), engine=1,
               zcoord=dict(za_k=128, zb_k=128, a_rows=128, a_cols=128, b_rows=128, b_cols=128))
