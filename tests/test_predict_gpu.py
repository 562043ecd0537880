"""GPU parity of the free-running decode (PREDICT mode, SURVEY §8 a20 / f1, BASELINE configs[4]) against the oracle.

The oracle (oracle/model.py: decoder_free_running) follows the reference's inference branch literally: it re-runs the decoder
self-attention over the WHOLE history every step (rnn_wrappers.py:111-124); the CUDA path serves row t from a key/value cache.
Tolerance: max|cuda - oracle| <= 1e-3 * max|oracle| per tensor (BASELINE.json north_star), alignments with 1e-6 absolute slack.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import model as OR  # noqa: E402

RTOL = 1e-3


def _mods():
    from importlib import import_module
    return (import_module("self-attention-tacotron_b200.engine"), import_module("self-attention-tacotron_b200.ops"),
            import_module("self-attention-tacotron_b200.lib"), import_module("self-attention-tacotron_b200.models"))


def _close(a, b, rtol, what, atol=0.0):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    d = (a - b).abs().max().item()
    s = b.abs().max().item()
    assert d <= rtol * s + atol, f"{what}: max abs err {d:.3e} vs scale {s:.3e}"


def _setup(satk, root, cfg, B, Tt, overrides=None, seed=3):
    E, O, L, M = _mods()
    hp = satk.load_hparams(os.path.join(root, "examples", cfg), overrides)
    d = satk.dims_from_hparams(hp)
    ps = satk.ParamStore(d).init(seed, "random")
    f, _ = satk.synthetic_batch(hp, B, Tt, 8 * d.r, seed=seed + 1)
    fd = satk.SourceData(*[x.cuda() if torch.is_tensor(x) else x for x in f])
    eng = E.TacotronEngine(hp, "cuda", params=ps)
    return hp, d, ps, f, fd, eng


def _compare(out, ref, d, T):
    assert out["steps"] == T
    _close(out["mel"], ref["mel"], RTOL, "mel")
    _close(out["stop"], ref["stop"], RTOL, "stop logits", 1e-5)
    _close(out["alignment"], ref["alignment"], RTOL, "alignment", 1e-6)
    if d.dual:
        _close(out["alignment2"], ref["alignment2"], RTOL, "alignment2", 1e-6)


@pytest.mark.parametrize("cfg,overrides,B,Tt,T", [
    ("ljspeech_self-attention-tacotron.json", None, 3, 20, 14),
    ("ljspeech_self-attention-tacotron.json", "attention=location_sensitive,cumulative_weights=True", 2, 17, 9),
    ("ljspeech_self-attention-tacotron.json", "use_forward_attention_transition_agent=True", 5, 23, 12),
    ("ljspeech_self-attention-tacotron.json", "decoder_self_attention_num_hop=2", 2, 19, 10),
    ("ljspeech_tacotron.json", None, 2, 21, 11),
    ("vctk_self-attention-tacotron.json", None, 4, 18, 10),
])
def test_free_running_matches_oracle(satk, root, cfg, overrides, B, Tt, T):
    hp, d, ps, f, fd, eng = _setup(satk, root, cfg, B, Tt, overrides)
    with torch.no_grad():
        ref = OR.model_predict(ps.as_dict(), d, f, max_iters=T, use_stop_token=False)
    for use_graph in (False, True, True):                  # eager, graph capture, cached-graph replay
        out = eng.predict(fd, max_iters=T, use_stop_token=False, use_graph=use_graph)
        torch.cuda.synchronize()
        _compare(out, ref, d, T)
    probs = [p_.clone() for p_ in out["dec_self_P"]]
    eng.fused_decode_tail = False                          # per-layer launches (rowgemm + sa_step) instead of the fused tail kernel
    out = eng.predict(fd, max_iters=T, use_stop_token=False, use_graph=False)
    _compare(out, ref, d, T)
    for a, b_ in zip(probs, out["dec_self_P"]):            # decoder self-attention alignments of both paths agree
        _close(a, b_, RTOL, "decoder self-attention alignment", 1e-6)


def test_stop_token_terminates_like_the_helper(satk, root):
    """StopTokenBasedInferenceHelper: the loop ends after the first step t > min_iters at which every utterance's
    sigmoid(stop) > 0.5.  A large stop bias makes that t = min_iters + 1, i.e. min_iters + 2 executed steps."""
    hp, d, ps, f, fd, eng = _setup(satk, root, "ljspeech_self-attention-tacotron.json", 3, 16)
    eng.ps.p["dec.stop_proj.b"].fill_(50.0)
    P = {k: v.clone() for k, v in ps.as_dict().items()}
    P["dec.stop_proj.b"] = torch.full_like(P["dec.stop_proj.b"], 50.0)
    with torch.no_grad():
        ref = OR.model_predict(P, d, f, max_iters=40, min_iters=4, use_stop_token=True)
    assert ref["stop"].shape[1] == 6
    out = eng.predict(fd, max_iters=40, min_iters=4, use_stop_token=True, check_every=3)
    _compare(out, ref, d, 6)
    # never finishing: runs to max_iters
    eng.ps.p["dec.stop_proj.b"].fill_(-50.0)
    out = eng.predict(fd, max_iters=15, min_iters=4, use_stop_token=True)
    assert out["steps"] == 15 and out["mel"].shape == (3, 15 * d.r, d.n_mels)


def test_free_running_equals_teacher_forcing_on_its_own_output_full_size(satk, root):
    """Size-independent property at BASELINE configs[4] size (B=16, T_text=148, 500 steps = 1000 frames): feeding the
    free-running prediction back as the teacher-forcing target (eval mode) must reproduce it — step kernels + KV cache on one
    side, the persistent cluster kernels + dense causal attention on the other."""
    E, O, L, M = _mods()
    hp = satk.load_hparams(os.path.join(root, "examples", "ljspeech_self-attention-tacotron.json"))
    d = satk.dims_from_hparams(hp)
    B, Tt, T = 16, 148, 500
    eng = E.TacotronEngine(hp, "cuda", seed=5)
    f, l = satk.synthetic_batch(hp, B, Tt, T * d.r, seed=9, device="cuda")
    out = eng.predict(f, max_iters=T, use_stop_token=False)
    mel = out["mel"].clone()
    al1, stop = out["alignment"].clone(), out["stop"].clone()
    assert mel.shape == (B, T * d.r, d.n_mels) and torch.isfinite(mel).all()
    s = al1.sum(1)
    assert (s - 1).abs().max().item() < 1e-4                                    # alignments are distributions over the text
    lab = l._replace(mel=mel)
    tf = eng.forward(f, lab, False)
    mel_tf = tf["mel_tm"].view(T, B, d.r, d.n_mels).permute(1, 0, 2, 3).reshape(B, T * d.r, d.n_mels)
    _close(mel_tf, mel, RTOL, "teacher-forced mel on free-running output")
    _close(tf["align1_tm"].permute(1, 2, 0), al1, RTOL, "alignment", 1e-5)
    _close(tf["stop_tm"].view(T, B).t(), stop, RTOL, "stop", 1e-4)


def test_rowgemm_unit(satk):
    E, O, L, M = _mods()
    g = torch.Generator().manual_seed(0)
    for Mr, K, Ns in ((16, 672, (1024,)), (5, 160, (256,)), (33, 256, (256, 256, 256)), (16, 256, (160, 1))):
        A = torch.randn(3, Mr, K + 7, generator=g).cuda()
        tdev = torch.tensor([2], dtype=torch.int32, device="cuda")
        mats, refs = [], []
        for i, N in enumerate(Ns):
            W = (torch.randn(K, N, generator=g) / K ** 0.5).cuda()
            b = torch.randn(N, generator=g).cuda()
            res = torch.randn(Mr, N, generator=g).cuda()
            Cc = torch.zeros(4, Mr, N + 3, device="cuda")
            act = (None, "relu", "tanh")[i % 3]
            mats.append(dict(W=W, bias=b, act=act, residual=res if i == 0 else None, C=Cc, ldc=N + 3, c_off=Mr * (N + 3),
                             c_tstride=Mr * (N + 3)))
            y = A[2, :, 5:5 + K].double() @ W.double() + b.double()
            y = {None: y, "relu": torch.relu(y), "tanh": torch.tanh(y)}[act]
            refs.append((y + res.double()) if i == 0 else y)
        dsc = O.rowgemm_desc(A, Mr, K, mats, lda=K + 7, a_off=5, a_tstride=Mr * (K + 7), t_ptr=tdev)
        O.rowgemm(dsc)
        torch.cuda.synchronize()
        for m, r in zip(mats, refs):
            N = r.shape[1]
            _close(m["C"][3, :, :N], r, 1e-5, f"rowgemm M={Mr} K={K} N={N}")
            assert m["C"][:3].abs().max().item() == 0 and m["C"][3, :, N:].abs().max().item() == 0


def test_rowgemm_lstm_epilogue_unit(satk):
    """Skinny GEMM with the ZoneoutLSTMCell epilogue (inference interpolation) against the oracle cell, parity buffers included."""
    E, O, L, M = _mods()
    g = torch.Generator().manual_seed(1)
    for Mr, Kx, H in ((16, 416, 256), (5, 40, 64), (20, 256, 128)):
        K = Kx + H
        W = (torch.randn(K, 4 * H, generator=g) / K ** 0.5)
        bias = torch.randn(4 * H, generator=g)
        x = torch.randn(Mr, Kx, generator=g)
        c0, h0 = torch.randn(Mr, H, generator=g), torch.randn(Mr, H, generator=g) * 0.5
        zc, zh = 0.1, 0.2
        out_ref, c_ref, h_ref = OR.zoneout_lstm_step(x, c0, h0, W, bias, None, None, zc, zh, False)
        for t in (0, 1):
            rows = torch.zeros(2, Mr, K)
            rows[t & 1, :, :Kx] = x
            rows[t & 1, :, Kx:] = h0
            rows = rows.cuda()
            outb = torch.zeros(2, Mr, H + 3, device="cuda")
            c, h = c0.cuda().clone(), h0.cuda().clone()
            tdev = torch.tensor([t], dtype=torch.int32, device="cuda")
            dsc = O.rowgemm_desc(rows, Mr, K, [dict(W=W.cuda(), bias=bias.cuda())], a_pstride=Mr * K, t_ptr=tdev,
                                 lstm=dict(H=H, c=c, h=h, zc=zc, zh=zh, forget_bias=1.0, out=outb, ld_out=H + 3, out_pstride=Mr * (H + 3),
                                           hdst=rows, ld_hdst=K, hdst_off=Kx, hdst_pstride=Mr * K))
            O.rowgemm(dsc)
            torch.cuda.synchronize()
            _close(outb[t & 1, :, :H], out_ref, 1e-5, "lstm out")
            _close(c, c_ref, 1e-5, "lstm c")
            _close(h, h_ref, 1e-5, "lstm h")
            _close(rows[(t + 1) & 1, :, Kx:], h_ref, 1e-5, "next-step h rows")
            assert outb[(t + 1) & 1].abs().max().item() == 0


def test_tfrecord_files_to_prediction_files(satk, root, tmp_path):
    """The callers either side of the path (train.py:40-88, predict_mel.py:36-74) on the reference's own file formats: TFRecord
    source / target files -> input_fn -> train two steps -> free-running predict -> <key>.mfbsp + <key>.tfrecord outputs."""
    import numpy as np
    E, O, L, M = _mods()
    TF = satk.tfrecord
    hp = satk.load_hparams(os.path.join(root, "examples", "ljspeech_self-attention-tacotron.json"), "batch_size=3,max_iters=9")
    rs = np.random.RandomState(0)
    srcs, mels = [], []
    for i in range(6):
        n, Lm = 9 + i, 11 + 3 * i
        srcs.append(TF.PreprocessedSourceData(i, f"LJ{i:03d}", np.concatenate([[0], rs.randint(1, 68, n - 2), [0]]).astype(np.int64), n,
                                              f"text {i}", None, None, None))
        mels.append(TF.PreprocessedMelData(i, f"LJ{i:03d}", (rs.randn(Lm, hp.num_mels) * 8 - 40).astype(np.float32), hp.num_mels, Lm))
    sp, mp = str(tmp_path / "source.tfrecord"), str(tmp_path / "target.tfrecord")
    TF.write_records(sp, [TF.encode_source_record(s) for s in srcs])
    TF.write_records(mp, [TF.encode_mel_record(m) for m in mels])
    model = M.tacotron_model_factory(hp, str(tmp_path / "ckpt"), None)
    # (max_iters = 9 only keeps the free-running decode short: the training pipeline's max-output-length filter is switched off)
    spec = model.train(satk.tfrecord_input_fn([sp], [mp], hp, filter_max_output_length=False), steps=2)
    assert spec.train_op == 2 and torch.isfinite(spec.loss)
    out_dir = str(tmp_path / "pred")
    keys = TF.write_predictions(model.predict(satk.tfrecord_input_fn([sp], [mp], hp, batch_size=1, for_prediction=True)), out_dir)
    assert keys == [f"LJ{i:03d}" for i in range(6)]
    for i, k in enumerate(keys):
        mel = np.fromfile(os.path.join(out_dir, k + ".mfbsp"), dtype="<f4").reshape(-1, hp.num_mels)
        assert mel.shape[0] == 9 * hp.outputs_per_step and np.isfinite(mel).all()
        ex = TF.decode_example(next(TF.read_records(os.path.join(out_dir, k + ".tfrecord"), verify_data_crc=True)))
        assert ex["key"] == [k.encode()] and ex["mel_length"].tolist() == [mel.shape[0]] and len(ex["alignment"]) >= 2
        assert np.array_equal(np.frombuffer(ex["source"][0], "<i8"), srcs[i].source)
        gt = np.frombuffer(ex["ground_truth_mel"][0], "<f4").reshape(-1, hp.num_mels)
        assert gt.shape[0] == int(ex["ground_truth_mel_length"][0]) == (mels[i].target_length + 2 * hp.outputs_per_step + 1) // 2 * 2


@pytest.mark.parametrize("cfg", ["ljspeech_self-attention-tacotron.json", "ljspeech_tacotron.json"])
def test_forced_alignment_mode_matches_oracle(satk, root, cfg):
    """use_forced_alignment_mode outside training (models/models.py:384-427, modules/teacher_forcing_attention.py:13-78,
    models/attention_factories.py:40-66): a teacher-forced pass gives the alignments, the second decode feeds back its own output
    while the attention mechanisms replay them.  EVAL metrics / predictions and PREDICT (from a SourceDataForPrediction record)."""
    from importlib import import_module
    M = import_module("self-attention-tacotron_b200.models")
    hp = satk.load_hparams(os.path.join(root, "examples", cfg), "use_forced_alignment_mode=True")
    d = satk.dims_from_hparams(hp)
    assert d.forced_alignment
    ps = satk.ParamStore(d).init(5, "random")
    B, Tt, Tm = 3, 21, 24
    f, l = satk.synthetic_batch(hp, B, Tt, Tm, seed=9)
    ref = OR.model_forced_alignment(ps.as_dict(), d, f, l)
    model = M.tacotron_model_factory(hp, None, None)
    model.engine.ps.flat.copy_(ps.flat.cuda())
    model.engine.ps.bn_mean_flat.copy_(ps.bn_mean_flat.cuda())
    model.engine.ps.bn_var_flat.copy_(ps.bn_var_flat.cuda())
    model.engine.refresh_transposed()
    spec = model.model_fn(f, l, M.ModeKeys.EVAL, hp)
    scale = ref["mel"].abs().max().item()
    assert (spec.predictions["mel"].cpu() - ref["mel"]).abs().max().item() <= 1e-3 * scale
    assert (spec.predictions["alignment"].cpu() - ref["alignment"]).abs().max().item() <= 1e-5
    want = OR.spec_loss_l1(ref["mel"], l.mel, l.spec_loss_mask)
    assert abs(float(spec.scalars["mel_loss"]) - float(want)) <= 1e-3 * float(want)
    # the replayed alignments are those of the teacher-forced pass, the mel is not
    tf = OR.model_forward(ps.as_dict(), d, f, l, False)
    assert torch.allclose(ref["alignment"], tf["alignment"]) and (ref["mel"] - tf["mel"]).abs().max().item() > 1e-4
    rec = satk.SourceDataForPrediction(f.id, f.key, f.source, f.source_length, f.text, f.speaker_id, l.mel, l.mel_width, l.target_length)
    pred = model.model_fn(rec, None, M.ModeKeys.PREDICT, hp).predictions
    assert (pred["mel"].cpu() - ref["mel"]).abs().max().item() <= 1e-3 * scale
    with pytest.raises(NotImplementedError, match="TRAIN"):
        model.model_fn(f, l, M.ModeKeys.TRAIN, hp)
