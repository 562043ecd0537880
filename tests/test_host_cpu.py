"""CPU tests of the host side: config contract, batch contract, C-ABI surface, engine host logic (dry run),
and the data-parallel gradient averaging on gloo with world_size 2."""
import ctypes
import json
import os
import re
import subprocess
import sys

import pytest
import torch


def test_hparams_defaults_and_overrides(satk, root):
    hp = satk.default_hparams()
    assert hp.batch_size == 32 and hp.outputs_per_step == 2 and hp.attention_kernel == 31 and hp.zoneout_factor_cell == 0.1
    hp = satk.load_hparams(os.path.join(root, "examples", "ljspeech_self-attention-tacotron.json"), "batch_size=16,attention=forward")
    assert hp.tacotron_model == "DualSourceSelfAttentionTacotronModel" and hp.batch_size == 16
    assert hp.attention_kernel == 10 and hp.attention_filters == 5 and len(hp.average_mel_level_db) == 80
    hp.parse("encoder_prenet_out_units=[256,128],decay_learning_rate=false,initial_learning_rate=0.001")
    assert hp.encoder_prenet_out_units == [256, 128] and hp.decay_learning_rate is False
    with pytest.raises(ValueError):
        hp.parse("no_such_key=1")
    with pytest.raises(ValueError):
        satk.default_hparams().parse_json(json.dumps({"nope": 1}))
    assert "Hyperparameters:" in satk.hparams_debug_string(hp)


def test_every_example_config_loads(satk, root):
    for f in os.listdir(os.path.join(root, "examples")):
        hp = satk.load_hparams(os.path.join(root, "examples", f))
        d = satk.dims_from_hparams(hp)
        assert d.att_kernel == 10 and d.attention == "forward"
        assert d.use_speaker == f.startswith("vctk")


def test_unknown_model_rejected(satk, root):
    hp = satk.load_hparams(os.path.join(root, "examples", "ljspeech_tacotron.json"), "tacotron_model=Nope")
    with pytest.raises(ValueError, match="Unknown Tacotron model"):
        satk.dims_from_hparams(hp)


def test_batch_contract(satk, root):
    hp = satk.load_hparams(os.path.join(root, "examples", "vctk_self-attention-tacotron.json"))
    f, l = satk.synthetic_batch(hp, 5, 30, 40, seed=1)
    r = hp.outputs_per_step
    assert f.source.dtype == torch.int64 and f.source.shape == (5, 30) and l.mel.shape == (5, 40, 80)
    assert l.done.shape == (5, 20) and l.spec_loss_mask.shape == (5, 40) and l.binary_loss_mask.shape == (5, 20)
    for b in range(5):
        n, t = int(f.source_length[b]), int(l.target_length[b])
        assert (f.source[b, n:] == 0).all() and t % r == 0
        assert (l.mel[b, t:] == hp.silence_mel_level_db).all() and (l.mel[b, :r] == hp.silence_mel_level_db).all()
        assert l.spec_loss_mask[b, :t].all() and not l.spec_loss_mask[b, t:].any()
        assert l.done[b, t // r - 1] == 1 and (l.done[b, :t // r - 1] == 0).all() and (l.done[b, t // r:] == 1).all()
    assert ((f.speaker_id >= 225) & (f.speaker_id < 377)).all()


def test_abi_symbols_and_struct_sizes(satk, root):
    """The C-ABI library loads without a GPU and exports every symbol include/satk.h declares."""
    from importlib import import_module
    L = import_module("self-attention-tacotron_b200.lib")
    lib = L.load()
    hdr = open(os.path.join(root, "include", "satk.h")).read()
    declared = set(re.findall(r"\b(satk_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(L.SYMBOLS), declared ^ set(L.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    out = (ctypes.c_int * 5)()
    assert lib.satk_struct_sizes(out) == 0
    assert list(out) == [ctypes.sizeof(x) for x in (L.GemmDesc, L.LstmFwdDesc, L.LstmBwdDesc, L.AttnRnnFwdDesc, L.AttnRnnBwdDesc)]
    out5 = (ctypes.c_int * 5)()
    assert lib.satk_struct_sizes_decode(out5) == 0
    assert list(out5) == [ctypes.sizeof(x) for x in (L.RowGemmDesc, L.AttnStepDesc, L.SaStepDesc, L.SaTailDesc, L.MlpChainDesc)]
    assert lib.satk_version() >= 100


def test_workspace_size_helpers_follow_the_header(root):
    """ops.de_ws_floats / eg_sync_ints / SUMSQ_SCRATCH restate macros of include/satk.h: evaluate the macros and compare."""
    import re
    from importlib import import_module
    O = import_module("self-attention-tacotron_b200.ops")
    src = open(os.path.join(root, "include", "satk.h")).read()
    de_row = re.search(r"#define SATK_DE_ROW\(Tt\) (.+)", src).group(1)
    sync = re.search(r"#define SATK_EG_SYNC_INTS\(B\) (.+)", src).group(1)
    sumsq = int(re.search(r"#define SATK_SUMSQ_SCRATCH (\d+)", src).group(1))
    assert O.SUMSQ_SCRATCH == sumsq
    for Tt in (1, 31, 32, 33, 148, 192):
        row = eval(de_row.replace("/", "//"), {"Tt": Tt})
        assert row % 32 == 0 and row >= Tt and row - Tt < 32
        for Td, B in ((1, 1), (400, 32), (7, 5)):
            assert O.de_ws_floats(Td, B, Tt) == Td * B * (2 * row + 8 * Tt)
    for B in (1, 5, 32, 64):
        assert O.eg_sync_ints(B) == eval(sync, {"B": B})
    assert (O.EG_FEATURES, O.EG_GRADIENTS, O.EG_PREPARED) == tuple(
        int(re.search(rf"#define SATK_EG_{n} (\d+)", src).group(1)) for n in ("FEATURES", "GRADIENTS", "PREPARED"))


def test_engine_refuses_cpu(satk, root):
    from importlib import import_module
    E = import_module("self-attention-tacotron_b200.engine")
    L = import_module("self-attention-tacotron_b200.lib")
    hp = satk.load_hparams(os.path.join(root, "examples", "ljspeech_tacotron.json"))
    with pytest.raises(L.SatkError, match="no CPU fallback"):
        E.TacotronEngine(hp, "cpu")


def test_engine_host_logic_dry_run(root):
    """Every GEMM the engine issues stays inside its tensors (extent-checking stubs instead of kernels)."""
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "dryrun.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("dry run ok") == 5


def test_noam_lr(satk):
    from importlib import import_module
    E = import_module("self-attention-tacotron_b200.engine")
    from oracle import model as OR
    for step in (0, 1, 3999, 4000, 100000):
        assert abs(E.noam_lr(0.0005, step, 1) - OR.noam_lr(0.0005, step, 1)) < 1e-15
    assert abs(E.noam_lr(0.002, 3999, 1) - 0.002) < 1e-9       # peak at the end of warm-up


_DP_SCRIPT = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["SATK_ROOT"])
import satk_path; satk = satk_path.load()
from oracle import model as OR
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
hp = satk.load_hparams(os.path.join(os.environ["SATK_ROOT"], "examples", "ljspeech_tacotron.json"))
d = satk.dims_from_hparams(hp)
ps = satk.ParamStore(d).init(5, "random")
B, Tt, Tm = 2, 10, 12
shards = [satk.synthetic_batch(hp, B, Tt, Tm, seed=40 + r, full_length=True) for r in range(world)]
f, l = shards[rank]
tr = OR.OracleTrainer(d, hp, ps.as_dict(), dtype=torch.float64)
_, grads, _ = tr.loss_and_grads(f, l, None, training=False)
flat = torch.cat([grads[n].reshape(-1) for n in tr.names])
dist.all_reduce(flat)                      # ONE all-reduce(sum) of the flat gradient (SURVEY 8e)
flat /= world                              # mean of per-replica gradients (MirroredStrategy semantics)
if rank == 0:
    # single-process reference: mean over replicas of each replica's masked-mean loss
    tot = None
    for (ff, ll) in shards:
        t2 = OR.OracleTrainer(d, hp, ps.as_dict(), dtype=torch.float64)
        _, g2, _ = t2.loss_and_grads(ff, ll, None, training=False)
        v = torch.cat([g2[n].reshape(-1) for n in t2.names])
        tot = v if tot is None else tot + v
    tot /= world
    err = (flat - tot).abs().max().item()
    print("DP_ERR", err)
    assert err < 1e-12
dist.destroy_process_group()
'''


def test_data_parallel_gradient_mean_gloo(root, tmp_path):
    script = tmp_path / "dp.py"
    script.write_text(_DP_SCRIPT)
    env = dict(os.environ, SATK_ROOT=root)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", str(script)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DP_ERR" in r.stdout


def test_unbuilt_model_switches_are_refused(satk, root):
    """Options of the reference's model_fn that are not built must fail loudly, never be silently ignored."""
    cfg = os.path.join(root, "examples", "ljspeech_self-attention-tacotron.json")
    for flag in ("use_external_speaker_embedding", "use_language_embedding", "use_accent_type",
                 "speaker_embedd_to_decoder", "apply_dropout_on_inference", "language_embedd_to_input", "language_embedd_to_decoder"):
        with pytest.raises(NotImplementedError, match=flag):
            satk.dims_from_hparams(satk.load_hparams(cfg, f"{flag}=True"))
    for ov in ("speaker_for_synthesis=3", "attention_filters=32", "attention_kernel=40", "attention_out_units=128"):
        with pytest.raises(NotImplementedError):
            satk.dims_from_hparams(satk.load_hparams(cfg, ov))
    with pytest.raises(NotImplementedError, match="speaker_embedd_to_prenet"):
        satk.dims_from_hparams(satk.load_hparams(os.path.join(root, "examples", "vctk_self-attention-tacotron.json"), "speaker_embedd_to_prenet=False"))
    # forced-alignment mode is built for EVAL / PREDICT (tests/test_predict_gpu.py); TRAIN refuses at the call
    assert satk.dims_from_hparams(satk.load_hparams(cfg, "use_forced_alignment_mode=True")).forced_alignment
    # PostNetV2 is built (tests/test_parity_gpu.py::test_postnet_v2); its speaker / channel conditioned variants are not
    dp = satk.dims_from_hparams(satk.load_hparams(cfg, "use_postnet_v2=True"))
    assert dp.postnet_v2 and (dp.postnet_layers, dp.postnet_kernel, dp.postnet_ch, dp.postnet_drop) == (5, 5, 512, 0.5)
    names = [n for n, _, _ in satk.param_specs(dp)]
    assert "postnet.conv4.gamma" in names and names[-1] == "postnet.proj.b"
    assert len({satk.tf_variable_name(n, dp) for n in names}) == len(names)
    for flag in ("speaker_embedd_to_postnet", "channel_id_to_postnet"):
        with pytest.raises(NotImplementedError, match=flag):
            satk.dims_from_hparams(satk.load_hparams(cfg, f"use_postnet_v2=True,{flag}=True"))
    satk.dims_from_hparams(satk.load_hparams(cfg, "cumulative_weights=True,use_forward_attention_transition_agent=True,use_l2_regularization=True"))


def test_loss_step_end_and_train_batch_order(satk, root):
    """Host logic of the step_end contract (satk_attn_rnn_bwd_desc / satk_lstm_bwd_desc): 1 + the last decoder step with a non-zero
    loss mask, and the TRAIN batch order built from it — a permutation; the 8 shortest targets form the last two clusters; head and
    tail are each ordered by source length (two-level), or the whole batch by (target, source) when one wave of clusters suffices."""
    from importlib import import_module
    E = import_module("self-attention-tacotron_b200.engine")
    hp = satk.load_hparams(os.path.join(root, "examples", "ljspeech_self-attention-tacotron.json"))
    f, l = satk.synthetic_batch(hp, 32, 148, 800, seed=5)
    se = E.loss_step_end(l, 400, 2)
    assert se.dtype == torch.int32 and se.shape == (32,)
    want = torch.tensor([max(1, -(-int(t) // 2)) for t in l.target_length], dtype=torch.int32)     # ceil(target_length / r)
    assert torch.equal(se, want) and int(se.max()) == 400 and int(se.min()) < 320
    # a mask with a hole and an all-zero mask: last non-zero step counts, never below 1
    l2 = l._replace(binary_loss_mask=torch.zeros_like(l.binary_loss_mask), spec_loss_mask=torch.zeros_like(l.spec_loss_mask))
    l2.spec_loss_mask[0, 10] = 1.0
    l2.spec_loss_mask[0, 101] = 1.0
    se2 = E.loss_step_end(l2, 400, 2)
    assert int(se2[0]) == 51 and int(se2[1]) == 1
    src = f.source_length
    perm = E.train_batch_order(se, src, 148, True)
    assert sorted(perm.tolist()) == list(range(32))
    tail, head = perm[24:], perm[:24]
    assert int(se[tail].max()) <= int(se[head].min())                      # the 8 shortest targets are last
    assert (src[tail][1:] <= src[tail][:-1]).all() and (src[head][1:] <= src[head][:-1]).all()
    flat = E.train_batch_order(se, src, 148, False)
    key = se.to(torch.int64) * 149 + src
    assert (key[flat][1:] <= key[flat][:-1]).all() and set(flat[24:].tolist()) == set(tail.tolist())
    small = E.train_batch_order(se[:28], src[:28], 148, True)                # one wave of clusters: plain (target, source) order
    assert (key[:28][small][1:] <= key[:28][small][:-1]).all()
    assert (src[E.train_batch_order(None, src, 148)][1:] <= src[E.train_batch_order(None, src, 148)][:-1]).all()


def test_header_is_plain_c(root, tmp_path):
    """include/satk.h is the drop-in boundary: it must compile as C99 (no C++ types in the signatures), alone."""
    src = tmp_path / "abi.c"
    src.write_text('#include "satk.h"\nint main(void) { satk_gemm_desc g; satk_attn_rnn_bwd_desc a; satk_lstm_bwd_desc l; '
                   '(void)g; (void)a; (void)l; return (int)sizeof(g) == 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"), "-fsyntax-only", str(src)],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr


def test_warm_start_selection_by_variable_name_regex(satk, root):
    """train.py:76-78 / hparams.py:200-202: vars_to_warm_start regular expressions select trainable tensors by (TF-style or own) name."""
    hp = satk.load_hparams(os.path.join(root, "examples", "ljspeech_self-attention-tacotron.json"))
    d = satk.dims_from_hparams(hp)
    names = list(satk.ParamStore(d, "cpu").offsets.keys())
    assert hp.vars_to_warm_start == [".*"] and satk.select_warm_start(names, d, hp.vars_to_warm_start) == names
    enc = satk.select_warm_start(names, d, "encoder/")
    assert enc and all(n.startswith(("enc.", "cbhg.")) for n in enc) and "enc.prenet0.W" in enc and "dec.lstm2.W" not in enc
    both = satk.select_warm_start(names, d, ["embedding", r"decoder/.*lstm_cell"])
    assert set(both) == {"embedding", "dec.lstm1.W", "dec.lstm1.b", "dec.lstm2.W", "dec.lstm2.b", "dec.lstm3.W", "dec.lstm3.b"}
    assert satk.select_warm_start(names, d, r"att1\.") == [n for n in names if n.startswith("att1.")]     # the store's own names match too
    # in-tree anchored names (forward_attention.py:17-26,73,78; module.py:717-723)
    assert satk.tf_variable_name("att1.loc_conv.W", d) == "decoder/ForwardAttention/location_features_convolution/kernel"
    assert satk.tf_variable_name("att1.v", d) == "decoder/ForwardAttention/attention_variable"
    assert satk.tf_variable_name("dec.stop_proj.W", d) == "decoder/stop_token_projection/kernel"
    assert len({satk.tf_variable_name(n, d) for n in names}) == len(names)                                 # the map is injective
