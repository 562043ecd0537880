"""CPU tests of the oracle restatement: invariants, the reference's one property test
(modules/transformer_test.py:40-82) re-expressed, numpy-vs-torch twins, and the golden fixtures."""
import os

import numpy as np
import pytest
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import model as OR
from oracle import np_primitives as NP


def _small(satk, cfg="ljspeech_self-attention-tacotron.json", B=3, Tt=14, Tm=16, seed=3, overrides=None, root=None):
    hp = satk.load_hparams(os.path.join(root, "examples", cfg), overrides)
    d = satk.dims_from_hparams(hp)
    ps = satk.ParamStore(d).init(seed, "random")
    f, l = satk.synthetic_batch(hp, B, Tt, Tm, seed=seed)
    masks = satk.make_masks(d, B, Tt, Tm // d.r, seed=seed)
    return hp, d, ps, f, l, masks


def test_param_count_matches_survey(satk, root):
    hp = satk.load_hparams(os.path.join(root, "examples", "ljspeech_self-attention-tacotron.json"))
    assert satk.num_trainable(satk.dims_from_hparams(hp)) == 6246104       # SURVEY §8(a) total / BASELINE.md §4


def test_invariants(satk, root):
    hp, d, ps, f, l, masks = _small(satk, root=root)
    out = OR.model_forward(ps.as_dict(), d, f, l, True, masks)
    al = out["alignment"]                                   # (B, Tt, Td)
    assert torch.allclose(al.sum(1), torch.ones_like(al.sum(1)), atol=1e-5)
    for b in range(al.shape[0]):
        n = int(f.source_length[b])
        if n < al.shape[1]:
            assert al[b, n:].abs().max() == 0                            # zero past source_length
            assert out["memory1"][b, n:].abs().max() == 0                # BiLSTM output zero past length
    assert (f.source_length < al.shape[1]).any()
    assert torch.allclose(out["alignment2"].sum(1), torch.ones_like(al.sum(1)), atol=1e-5)
    P = out["dec_self_alignments"][0].transpose(1, 2)       # back to (B, Tq, Tk)
    assert P.triu(1).abs().max() == 0                       # causal
    assert torch.isfinite(out["loss"])


def test_first_step_alpha_is_one_hot(satk, root):
    """alpha_0 = one-hot(0) (forward_attention.py:131-133): after one step alpha concentrates on j in {0,1}... the
    recursion can only move one position per step."""
    hp, d, ps, f, l, masks = _small(satk, root=root)
    out = OR.model_forward(ps.as_dict(), d, f, l, False, None)
    al = out["alignment"]
    for t in range(min(4, al.shape[2])):
        assert al[:, t + 2:, t].max() < 1e-4                # mass cannot be further than t+1 after t+1 steps


@settings(max_examples=12, deadline=None)
@given(bs=st.integers(1, 3), r=st.integers(1, 2), tf=st.integers(2, 8), dim=st.integers(1, 10).map(lambda x: 2 * x),
       seed=st.integers(0, 1000))
def test_training_branch_equals_inference_branch(bs, r, tf, dim, seed):
    """modules/transformer_test.py:40-82: RNNTransformer training branch (batched causal self-attention after the
    loop) == inference branch (step-wise re-attention over the growing history); drop_rate 0, 2 heads, 1 hop."""
    g = torch.Generator().manual_seed(seed)
    D = dim * r
    T = tf
    x = torch.randint(-1, 2, (bs, T, D), generator=g).float()
    P = {}
    for nm in ("key", "value", "query", "output", "transform"):
        P[f"sa.{nm}.W"] = torch.randn(D, D, generator=g) * 0.3
        P[f"sa.{nm}.b"] = torch.randn(D, generator=g) * 0.1
    full, _ = OR.self_attention_transformer(x, P, "sa", 2, True, None, 1.0)
    rows = []
    for t in range(T):
        h, _ = OR.self_attention_transformer(x[:, :t + 1], P, "sa", 2, True, None, 1.0)
        rows.append(h[:, -1])
    step = torch.stack(rows, 1)
    assert torch.allclose(full, step, rtol=1e-5, atol=1e-5)


def test_decoder_inference_branch_matches(satk, root):
    hp, d, ps, f, l, masks = _small(satk, B=2, Tt=10, Tm=12, root=root)
    P = ps.as_dict()
    m1, m2, _ = OR.encoder_forward(P, d, f.source, f.source_length, False)
    a = OR.decoder_forward(P, d, m1, m2, f.source_length, l.mel, None, False)
    b = OR.decoder_forward(P, d, m1, m2, f.source_length, l.mel, None, False, inference_branch=True)
    assert torch.allclose(a[0], b[0], atol=2e-5) and torch.allclose(a[1], b[1], atol=2e-5)


def test_numpy_twins():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 9, 5))
    for k in (1, 2, 3, 4, 10):
        W = rng.standard_normal((k, 5, 4))
        ref = NP.conv1d_same(x, W)
        got = OR.conv1d_same(torch.tensor(x), torch.tensor(W)).numpy()
        assert np.allclose(ref, got, atol=1e-10), k
    assert np.allclose(NP.maxpool2_same(x), OR.maxpool2_same(torch.tensor(x)).numpy())
    H = 6
    W = rng.standard_normal((5 + H, 4 * H)); b = rng.standard_normal(4 * H)
    c = rng.standard_normal((2, H)); h = rng.standard_normal((2, H)); xi = rng.standard_normal((2, 5))
    cn, hn = NP.lstm_cell(xi, c, h, W, b)
    cn2, hn2 = OR.lstm_cell(torch.tensor(xi), torch.tensor(c), torch.tensor(h), torch.tensor(W), torch.tensor(b))
    assert np.allclose(cn, cn2.numpy()) and np.allclose(hn, hn2.numpy())


def test_forward_attention_twin(satk, root):
    hp, d, ps, f, l, masks = _small(satk, root=root)
    P = {k: v.double() for k, v in ps.as_dict().items()}
    rng = np.random.default_rng(1)
    T, B = 14, 2
    keys = torch.tensor(rng.standard_normal((B, T, d.att1)))
    values = torch.tensor(rng.standard_normal((B, T, d.mem1)))
    q = torch.tensor(rng.standard_normal((B, d.att_rnn)))
    lens = torch.tensor([T, 9])
    pa = torch.tensor(rng.random((B, T))); pa[1, 9:] = 0; pa = pa / pa.sum(1, keepdim=True)
    pal = torch.tensor(rng.random((B, T))); pal[1, 9:] = 0; pal = pal / pal.sum(1, keepdim=True)
    alpha, st_ = OR.attention1_step(P, d, q, (pa, pal, torch.full((B, 1), 0.5, dtype=torch.float64)), keys, values, lens)
    for b in range(B):
        a_np, al_np = NP.forward_attention_step(q[b].numpy(), pa[b].numpy(), pal[b].numpy(), keys[b].numpy(), int(lens[b]),
                                                P["att1.query.W"].numpy(), P["att1.loc_conv.W"].numpy(), P["att1.loc_conv.b"].numpy(),
                                                P["att1.loc_layer.W"].numpy(), P["att1.v"].numpy(), P["att1.b"].numpy())
        assert np.allclose(st_[0][b].numpy(), a_np, atol=1e-10)
        assert np.allclose(alpha[b].numpy(), al_np, atol=1e-10)


def test_fp32_vs_fp64_gradients(satk, root):
    """The fp32 oracle's autograd gradients agree with the fp64 run (the gradient oracle of SURVEY §7 step 1)."""
    hp, d, ps, f, l, masks = _small(satk, B=2, Tt=10, Tm=12, root=root)
    t32 = OR.OracleTrainer(d, hp, ps.as_dict())
    t64 = OR.OracleTrainer(d, hp, ps.as_dict(), dtype=torch.float64)
    _, g32, _ = t32.loss_and_grads(f, l, masks)
    _, g64, _ = t64.loss_and_grads(f, l, masks)
    for n in t32.names:
        s = g64[n].abs().max().item()
        assert (g32[n].double() - g64[n]).abs().max().item() <= 2e-4 * max(s, 1e-6) + 1e-9, n


def test_golden_fixture(satk, root):
    """tests/golden/oracle_small.npz was produced by oracle/make_golden.py (self-generated: the reference ships no
    vectors, SURVEY F6 — parity is UNPINNED; the fixture only freezes the oracle against accidental drift)."""
    z = np.load(os.path.join(root, "tests", "golden", "oracle_small.npz"))
    from oracle.make_golden import golden_case
    out, grads = golden_case(satk, root)
    assert np.allclose(out["mel"].detach().numpy(), z["mel"], atol=2e-5)
    assert np.allclose(out["alignment"].detach().numpy(), z["alignment"], atol=2e-6)
    assert abs(float(out["loss"]) - float(z["loss"])) < 1e-5
    assert np.allclose(grads["dec.lstm1.W"].numpy()[100:164, :64], z["grad_dec_lstm1_W_slice"], atol=1e-6)
    assert np.allclose(grads["att1.v"].numpy(), z["grad_att1_v"], atol=1e-6)


def test_golden_fixture_predict_and_variants(satk, root):
    """tests/golden/oracle_predict.npz (oracle/make_golden.py): the free-running branch and the transition-agent + cumulative-weights
    variant of the oracle, frozen against drift (self-generated, like oracle_small.npz)."""
    z = np.load(os.path.join(root, "tests", "golden", "oracle_predict.npz"))
    from oracle.make_golden import golden_predict_case
    res = golden_predict_case(satk, root)
    assert set(res) == set(z.files)
    for k, v in res.items():
        assert np.allclose(np.asarray(v), z[k], atol=2e-5, rtol=1e-5), k


def test_golden_fixture_l2_regularization(satk, root):
    """tests/golden/oracle_l2.npz (oracle/make_golden.py): loss with the l2 term, the term itself and gradients of a regularised, a
    black-listed and an attention tensor of the single-attention model, frozen against drift (self-generated)."""
    z = np.load(os.path.join(root, "tests", "golden", "oracle_l2.npz"))
    from oracle.make_golden import golden_l2_case
    res = golden_l2_case(satk, root)
    assert set(res) == set(z.files)
    for k, v in res.items():
        assert np.allclose(np.asarray(v), z[k], atol=2e-6, rtol=1e-5), k
    assert 0.05 < float(z["regularization_loss"]) / float(z["loss"]) < 0.5


def test_free_running_oracle_is_consistent_with_teacher_forcing(satk, root):
    """PREDICT-branch restatement (oracle.decoder_free_running, module.py:762-778) vs the teacher-forced restatement: feeding
    the free-running output back as the target reproduces it (eval mode), for the dual and the single-attention model."""
    for cfg in ("ljspeech_self-attention-tacotron.json", "ljspeech_tacotron.json"):
        hp = satk.load_hparams(os.path.join(root, "examples", cfg))
        d = satk.dims_from_hparams(hp)
        P = satk.ParamStore(d).init(11, "random").as_dict()
        f, l = satk.synthetic_batch(hp, 2, 12, 8 * d.r, seed=4)
        with torch.no_grad():
            pr = OR.model_predict(P, d, f, max_iters=8, use_stop_token=False)
            m1, m2, _ = OR.encoder_forward(P, d, f.source, f.source_length, False, None, None)
            mel, stop, al1, al2, _ = OR.decoder_forward(P, d, m1, m2, f.source_length, pr["mel"], None, False)
        assert pr["mel"].shape == (2, 8 * d.r, d.n_mels)
        assert torch.allclose(mel, pr["mel"], atol=2e-5) and torch.allclose(stop.squeeze(-1), pr["stop"], atol=2e-5)
        assert torch.allclose(al1, pr["alignment"], atol=1e-6)


def test_free_running_oracle_stop_token(satk, root):
    hp = satk.load_hparams(os.path.join(root, "examples", "ljspeech_self-attention-tacotron.json"))
    d = satk.dims_from_hparams(hp)
    P = satk.ParamStore(d).init(11, "random").as_dict()
    P["dec.stop_proj.b"] = torch.full_like(P["dec.stop_proj.b"], 50.0)
    f, _ = satk.synthetic_batch(hp, 2, 10, 8 * d.r, seed=4)
    with torch.no_grad():
        pr = OR.model_predict(P, d, f, max_iters=30, min_iters=3, use_stop_token=True)
    assert pr["stop"].shape == (2, 5)          # first t > min_iters is t = 4 -> 5 executed steps


def test_l2_regularization_selection_and_value(satk, root):
    """models/models.py:470-478 + regularizers.py:11-18: the oracle applies the reference's literal black-list to (recalled) TF
    variable names, the parameter store selects by its own naming rules; both must pick the same tensors, and the oracle's loss
    must be scale * sum(w^2) / 2 over them.  Embeddings, biases (attention_bias included), BN parameters, LSTM kernels and the stop
    projection are never regularised; the mel projection only in the transformer decoder ("out_projection/kernel")."""
    import os
    from importlib import import_module
    P = import_module("self-attention-tacotron_b200.params")
    for cfg, ov in (("ljspeech_self-attention-tacotron.json", "use_forward_attention_transition_agent=True"),
                    ("ljspeech_tacotron.json", "attention=additive"), ("vctk_self-attention-tacotron.json", None)):
        ov = (ov + "," if ov else "") + "use_l2_regularization=True,l2_regularization_weight=1e-3"
        hp = satk.load_hparams(os.path.join(root, "examples", cfg), ov)
        d = satk.dims_from_hparams(hp)
        assert d.l2_weight == 1e-3
        ps = satk.ParamStore(d).init(3, "random")
        mine = {n for n in ps.p if P.l2_regularized(d, n)}
        theirs = {n for n in ps.p if not any(b in OR.tf_variable_name(n, d) for b in OR.L2_BLACKLIST)}
        assert mine == theirs and len(mine) > 30
        for n in ("embedding", "att1.b", "cbhg.proj1.gamma", "dec.lstm1.W", "cbhg.lstm_fw.W", "dec.stop_proj.W", "enc.prenet0.b"):
            assert n not in mine
        assert ("dec.out_proj.W" in mine) == d.dual and "att1.v" in mine and "cbhg.bank3.W" in mine
        want = 1e-3 * sum(0.5 * float((ps.p[n].double() ** 2).sum()) for n in mine)
        got = float(OR.l2_regularization_loss(ps.as_dict(), d))
        assert abs(got - want) <= 1e-6 * want
        assert float(ps.l2_mask().sum()) == sum(ps.p[n].numel() for n in mine)
    d0 = satk.dims_from_hparams(satk.load_hparams(os.path.join(root, "examples", "ljspeech_tacotron.json")))
    assert d0.l2_weight == 0.0


def test_postnet_v2_restatement(satk, root):
    """PostNetV2 (models/models.py:92-100,116-118,210,440-462): residual structure, loss term, train / eval batch-norm modes, dropout
    scaling, gradients reach every post-net tensor; the numpy conv twin agrees with the torch conv used inside it."""
    ov = "use_postnet_v2=True,postnet_v2_out_channels=16,num_postnet_v2_layers=3"
    hp, d, ps, f, l, masks = _small(satk, root=root, overrides=ov)
    assert d.postnet_v2 and {"postnet.conv0", "postnet.conv1", "postnet.conv2"} <= set(masks)
    assert masks["postnet.conv1"].shape == (l.mel.shape[1], l.mel.shape[0], 16)            # frame-major [T_mel, B, channels]
    P = {k: v.clone() for k, v in ps.as_dict().items()}
    import dataclasses
    base = OR.model_forward(P, dataclasses.replace(d, postnet_v2=False), f, l, True, masks)
    out = OR.model_forward(P, d, f, l, True, masks)
    assert torch.allclose(out["mel"], base["mel"])                                          # the post-net does not touch the decoder
    assert torch.allclose(out["loss"], base["loss"] + out["postnet_v2_mel_loss"], atol=1e-6)  # models.py:118 / 482
    assert torch.allclose(out["postnet_v2_mel_loss"], OR.spec_loss_l1(out["mel_postnet"], l.mel, l.spec_loss_mask))
    # zero projection -> the residual connection alone
    P0 = dict(P, **{"postnet.proj.W": torch.zeros_like(P["postnet.proj.W"]), "postnet.proj.b": torch.zeros_like(P["postnet.proj.b"])})
    assert torch.allclose(OR.postnet_v2(P0, d, out["mel"], False), out["mel"])
    # eval mode: moving statistics, no dropout -> deterministic and different from the training pass
    ev1, ev2 = OR.postnet_v2(P, d, out["mel"], False), OR.postnet_v2(P, d, out["mel"], False)
    assert torch.equal(ev1, ev2) and not torch.allclose(ev1, out["mel_postnet"])
    # dropout: all-ones masks scale every layer by 1 / keep
    ones = {k: torch.ones_like(v) for k, v in masks.items()}
    stats = {}
    OR.postnet_v2(P, d, out["mel"], True, ones, stats)
    assert set(stats) == {"postnet.conv0", "postnet.conv1", "postnet.conv2"}
    # gradients
    tr = OR.OracleTrainer(d, hp, ps.as_dict())
    _, grads, _ = tr.loss_and_grads(f, l, masks, True)
    for n in ("postnet.conv0.W", "postnet.conv2.gamma", "postnet.conv1.beta", "postnet.proj.W", "postnet.proj.b", "dec.out_proj.W"):
        assert float(grads[n].abs().max()) > 0, n
    x = torch.randn(2, 9, 5)
    W = torch.randn(5, 5, 4)
    assert np.allclose(NP.conv1d_same(x.numpy(), W.numpy()), OR.conv1d_same(x, W).numpy(), atol=1e-5)
