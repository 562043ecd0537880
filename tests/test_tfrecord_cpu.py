"""TFRecord / tf.train.Example compatibility without TensorFlow: known-answer vectors for CRC-32C, byte-exact record
framing, the protobuf encoding cross-checked against the real `protobuf` runtime with the Example schema built at run time, and the
reference's input pipeline semantics (datasets/ljspeech/dataset.py:127-167,225-322; utils/tfrecord.py:82-104,135-152)."""
import os
import struct

import numpy as np
import pytest
import torch


def _TF():
    from importlib import import_module
    return import_module("self-attention-tacotron_b200.tfrecord")


def test_crc32c_known_answers():
    TF = _TF()
    assert TF.crc32c(b"123456789") == 0xE3069283                   # the CRC-32C check value (RFC 3720 appendix B.4)
    assert TF.crc32c(b"\x00" * 32) == 0x8A9136AA                    # RFC 3720: 32 bytes of zeros
    assert TF.crc32c(b"\xff" * 32) == 0x62A8AB43                    # RFC 3720: 32 bytes of ones
    assert TF.crc32c(bytes(range(32))) == 0x46DD794E                # RFC 3720: incrementing bytes
    assert TF.masked_crc32c(b"") == 0xA282EAD8                      # crc 0 -> the mask delta itself


def test_record_framing_is_byte_exact(tmp_path):
    TF = _TF()
    path = str(tmp_path / "x.tfrecord")
    TF.write_records(path, [b"abc", b""])
    raw = open(path, "rb").read()
    head = struct.pack("<Q", 3)
    exp = head + struct.pack("<I", TF.masked_crc32c(head)) + b"abc" + struct.pack("<I", TF.masked_crc32c(b"abc"))
    head0 = struct.pack("<Q", 0)
    exp += head0 + struct.pack("<I", TF.masked_crc32c(head0)) + struct.pack("<I", TF.masked_crc32c(b""))
    assert raw == exp
    assert list(TF.read_records(path, verify_data_crc=True)) == [b"abc", b""]
    bad = bytearray(raw)
    bad[13] ^= 1                                                     # flip a payload bit
    open(path, "wb").write(bytes(bad))
    with pytest.raises(IOError, match="payload"):
        list(TF.read_records(path, verify_data_crc=True))
    bad = bytearray(raw)
    bad[0] ^= 1                                                      # flip a length bit
    open(path, "wb").write(bytes(bad))
    with pytest.raises(IOError, match="length"):
        list(TF.read_records(path))


def _example_classes():
    """tensorflow/core/example/{example,feature}.proto rebuilt with descriptor_pb2 (field numbers are the public schema)."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="satk_test_example.proto", package="satk_test", syntax="proto3")
    T = descriptor_pb2.FieldDescriptorProto

    def msg(name):
        m = fd.message_type.add()
        m.name = name
        return m
    m = msg("BytesList"); m.field.add(name="value", number=1, type=T.TYPE_BYTES, label=T.LABEL_REPEATED)
    m = msg("FloatList"); m.field.add(name="value", number=1, type=T.TYPE_FLOAT, label=T.LABEL_REPEATED)
    m = msg("Int64List"); m.field.add(name="value", number=1, type=T.TYPE_INT64, label=T.LABEL_REPEATED)
    m = msg("Feature")
    m.oneof_decl.add(name="kind")
    m.field.add(name="bytes_list", number=1, type=T.TYPE_MESSAGE, type_name=".satk_test.BytesList", label=T.LABEL_OPTIONAL, oneof_index=0)
    m.field.add(name="float_list", number=2, type=T.TYPE_MESSAGE, type_name=".satk_test.FloatList", label=T.LABEL_OPTIONAL, oneof_index=0)
    m.field.add(name="int64_list", number=3, type=T.TYPE_MESSAGE, type_name=".satk_test.Int64List", label=T.LABEL_OPTIONAL, oneof_index=0)
    m = msg("Features")
    e = m.nested_type.add(name="FeatureEntry")
    e.options.map_entry = True
    e.field.add(name="key", number=1, type=T.TYPE_STRING, label=T.LABEL_OPTIONAL)
    e.field.add(name="value", number=2, type=T.TYPE_MESSAGE, type_name=".satk_test.Feature", label=T.LABEL_OPTIONAL)
    m.field.add(name="feature", number=1, type=T.TYPE_MESSAGE, type_name=".satk_test.Features.FeatureEntry", label=T.LABEL_REPEATED)
    m = msg("Example")
    m.field.add(name="features", number=1, type=T.TYPE_MESSAGE, type_name=".satk_test.Features", label=T.LABEL_OPTIONAL)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("satk_test.Example"))


def test_example_encoding_against_the_protobuf_runtime():
    TF = _TF()
    Example = _example_classes()
    mel = np.random.RandomState(0).randn(7, 5).astype("<f4")
    feats = {"id": np.array([-3], np.int64), "key": [b"LJ001-0001"], "mel": [mel.tobytes()], "mel_width": np.array([5], np.int64),
             "target_length": np.array([7], np.int64), "floats": np.array([0.5, -2.25, 3e7], np.float32), "empty": [],
             "many": np.array([0, 1, 127, 128, 2 ** 40, -2 ** 62], np.int64), "two": [b"a", b"\x00\xff"]}
    buf = TF.encode_example(feats)
    ex = Example()
    ex.ParseFromString(buf)                                        # ours -> real runtime
    f = ex.features.feature
    assert list(f["id"].int64_list.value) == [-3] and list(f["key"].bytes_list.value) == [b"LJ001-0001"]
    assert f["mel"].bytes_list.value[0] == mel.tobytes() and list(f["many"].int64_list.value) == feats["many"].tolist()
    assert list(f["floats"].float_list.value) == [0.5, -2.25, 3e7] and list(f["two"].bytes_list.value) == [b"a", b"\x00\xff"]
    back = TF.decode_example(ex.SerializeToString())               # real runtime -> ours
    for k, v in feats.items():
        if isinstance(v, list):
            assert back[k] == v, k
        else:
            assert np.array_equal(back[k], v) and back[k].dtype == v.dtype, k


def test_source_and_mel_records_round_trip(tmp_path):
    TF = _TF()
    src = TF.PreprocessedSourceData(5, "p225_001", np.array([0, 12, 33, 7, 0], np.int64), 5, "hello", 225, 23, 1)
    mel = TF.PreprocessedMelData(5, "p225_001", np.arange(12, dtype=np.float32).reshape(4, 3), 3, 4)
    sp, mp = str(tmp_path / "s.tfrecord"), str(tmp_path / "m.tfrecord")
    TF.write_records(sp, [TF.encode_source_record(src)])
    TF.write_records(mp, [TF.encode_mel_record(mel)])
    s2, m2 = next(TF.read_source_file(sp)), next(TF.read_mel_file(mp))
    assert s2.key == "p225_001" and s2.speaker_id == 225 and np.array_equal(s2.source, src.source) and s2.text == "hello"
    assert s2.phone is None and s2.phone_length is None
    assert m2.target_length == 4 and m2.mel_width == 3 and np.array_equal(m2.mel, mel.mel)
    # VCTK records also carry a phone sequence (datasets/vctk/dataset.py:74-76)
    src_p = src._replace(phone=np.array([0, 40, 41, 0], np.int64), phone_length=4, phone_txt="HH AH")
    TF.write_records(sp, [TF.encode_source_record(src_p)])
    s3 = next(TF.read_source_file(sp))
    assert np.array_equal(s3.phone, src_p.phone) and s3.phone_length == 4 and s3.phone_txt == "HH AH" and np.array_equal(s3.source, src.source)


def test_prepare_target_and_batching_follow_the_reference(satk, root, tmp_path):
    TF = _TF()
    hp = satk.load_hparams(os.path.join(root, "examples", "ljspeech_self-attention-tacotron.json"),
                           "batch_size=2,approx_min_target_length=4,batch_bucket_width=3,batch_num_buckets=5")
    r, nm = hp.outputs_per_step, hp.num_mels
    avg, std = np.asarray(hp.average_mel_level_db, np.float32), np.asarray(hp.stddev_mel_level_db, np.float32)
    rs = np.random.RandomState(1)
    lengths = [5, 8, 6, 3, 9]                                       # odd and even target lengths (r = 2)
    srcs, mels = [], []
    for i, L in enumerate(lengths):
        n = 4 + i
        srcs.append(TF.PreprocessedSourceData(i, f"utt{i}", np.concatenate([[0], rs.randint(1, 68, n - 2), [0]]).astype(np.int64), n,
                                              f"text {i}", None, None, None))
        mels.append(TF.PreprocessedMelData(i, f"utt{i}", rs.randn(L, nm).astype(np.float32) * 10 - 30, nm, L))
    t = satk.prepare_target(mels[0], hp)                            # L = 5 -> 5 + 2r = 9 -> padded to 10
    assert t["target_length"] == 10 and t["mel"].shape == (10, nm) and t["done"].tolist() == [0, 0, 0, 0, 1]
    assert np.allclose(t["mel"][r:r + 5], (mels[0].mel - avg) / std) and np.all(t["mel"][:r] == hp.silence_mel_level_db)
    assert np.all(t["mel"][r + 5:] == hp.silence_mel_level_db) and t["spec_loss_mask"].sum() == 10 and t["binary_loss_mask"].sum() == 5
    t = satk.prepare_target(mels[1], hp)                            # L = 8 -> 12, already a multiple of r: no extra padding
    assert t["target_length"] == 12
    sp, mp = str(tmp_path / "s.tfrecord"), str(tmp_path / "m.tfrecord")
    TF.write_records(sp, [TF.encode_source_record(s) for s in srcs])
    TF.write_records(mp, [TF.encode_mel_record(m) for m in mels])
    batches = list(satk.tfrecord_input_fn([sp], [mp], hp)())
    assert [b[0].source.shape[0] for b in batches] == [2, 2, 1]     # one bucket (the reference's key is <= 0 for every length)
    f, l = batches[0]
    assert f.key == ["utt0", "utt1"] and f.source.shape == (2, 5) and f.source[0, 4].item() == 0 and f.source_length.tolist() == [4, 5]
    assert l.mel.shape == (2, 12, nm) and l.target_length.tolist() == [10, 12]
    assert torch.all(l.mel[0, 10:] == hp.silence_mel_level_db) and l.done[0].tolist() == [0, 0, 0, 0, 1, 1]
    assert l.spec_loss_mask[0].tolist() == [1.0] * 10 + [0.0] * 2 and l.binary_loss_mask[0].tolist() == [1.0] * 5 + [0.0]
    fp, lp = next(iter(satk.tfrecord_input_fn([sp], [mp], hp, batch_size=1, for_prediction=True)()))
    assert type(fp).__name__ == "SourceDataForPrediction" and fp.mel.shape == (1, 10, nm) and fp.target_length.tolist() == [10]
    # the stages of train.py:53-54: filter_by_max_output_length (max_iters * r frames), repeat, shuffle buffer
    hp5 = satk.load_hparams(os.path.join(root, "examples", "ljspeech_self-attention-tacotron.json"),
                            "batch_size=2,approx_min_target_length=4,batch_bucket_width=3,batch_num_buckets=5,max_iters=5")
    kept = [k for b in satk.tfrecord_input_fn([sp], [mp], hp5)() for k in b[0].key]
    assert kept == ["utt0", "utt2", "utt3"]                        # padded targets 10, 12, 10, 8, 14 frames against 5 * 2
    assert len([k for b in satk.tfrecord_input_fn([sp], [mp], hp5, filter_max_output_length=False)() for k in b[0].key]) == 5
    assert len([k for b in satk.tfrecord_input_fn([sp], [mp], hp5, batch_size=1, for_prediction=True)() for k in b[0].key]) == 5
    it = satk.tfrecord_input_fn([sp], [mp], hp, repeat=True, shuffle_buffer_size=3, seed=4)()
    first = [k for _ in range(5) for k in next(it)[0].key]
    again = [k for _, b in zip(range(5), satk.tfrecord_input_fn([sp], [mp], hp, repeat=True, shuffle_buffer_size=3, seed=4)()) for k in b[0].key]
    assert len(first) == 10 and first == again and set(first) == {f"utt{i}" for i in range(5)} and first[:5] != [f"utt{i}" for i in range(5)]
    assert [k for b in satk.tfrecord_input_fn([sp], [mp], hp, max_source_length=6)() for k in b[0].key] == ["utt0", "utt1", "utt2"]
    # hparams.source == 'phone' (datasets/vctk/dataset.py:144-146): the phone sequence is the source; records without one are refused
    hpp = satk.load_hparams(os.path.join(root, "examples", "ljspeech_self-attention-tacotron.json"),
                            "batch_size=2,approx_min_target_length=4,batch_bucket_width=3,batch_num_buckets=5,source=phone")
    with pytest.raises(ValueError, match="phone"):
        next(iter(satk.tfrecord_input_fn([sp], [mp], hpp)()))
    srcs_p = [x._replace(phone=np.array([0, 50 + i, 0], np.int64), phone_length=3, phone_txt=f"P{i}") for i, x in enumerate(srcs)]
    TF.write_records(sp, [TF.encode_source_record(x) for x in srcs_p])
    f, l = next(iter(satk.tfrecord_input_fn([sp], [mp], hpp)()))
    assert f.source.tolist() == [[0, 50, 0], [0, 51, 0]] and f.source_length.tolist() == [3, 3] and f.text == ["P0", "P1"]


def test_prediction_outputs_round_trip(tmp_path):
    TF = _TF()
    mel = torch.arange(2 * 6 * 4, dtype=torch.float32).view(2, 6, 4)
    pred = {"id": torch.tensor([3, 4]), "key": ["a", "b"], "mel": mel, "ground_truth_mel": mel + 1, "source": torch.tensor([[0, 5, 0], [0, 6, 0]]),
            "text": ["ta", "tb"], "alignment": torch.rand(2, 3, 3), "alignment2": torch.rand(2, 3, 3)}
    keys = TF.write_predictions([pred], str(tmp_path))
    assert keys == ["a", "b"]
    raw = np.fromfile(str(tmp_path / "b.mfbsp"), dtype="<f4").reshape(-1, 4)       # predict_mel.py:61
    assert np.array_equal(raw, mel[1].numpy())
    ex = TF.decode_example(next(TF.read_records(str(tmp_path / "a.tfrecord"), verify_data_crc=True)))
    assert ex["key"] == [b"a"] and ex["mel_length"].tolist() == [6] and ex["mel_width"].tolist() == [4] and len(ex["alignment"]) == 2
    assert np.array_equal(np.frombuffer(ex["ground_truth_mel"][0], "<f4").reshape(6, 4), (mel[0] + 1).numpy())
    assert np.array_equal(np.frombuffer(ex["source"][0], "<i8"), np.array([0, 5, 0])) and ex["accent_type"] == []
