"""TFRecord / tf.train.Example compatibility without TensorFlow: the data formats either side of the hot path.

The reference reads its pre-processed corpora from TFRecord files of ``tf.train.Example`` protos
(/root/reference/utils/tfrecord.py:82-104 mel targets; /root/reference/datasets/ljspeech/dataset.py:50-72 and
/root/reference/datasets/vctk/dataset.py:64-98 sources) and ``predict_mel.py`` writes one TFRecord + one raw ``.mfbsp`` file per
utterance (/root/reference/predict_mel.py:56-74, /root/reference/utils/tfrecord.py:135-152).  TensorFlow cannot be installed here, so
this module restates the two public, stable formats it needs:

  * TFRecord framing (tensorflow/core/lib/io/record_writer.cc): ``uint64 length | uint32 masked_crc32c(length) | data |
    uint32 masked_crc32c(data)``, little endian, mask = rotr15(crc) + 0xa282ead8;
  * the protobuf wire encoding of ``Example{features=1}``, ``Features{map<string,Feature> feature=1}``,
    ``Feature{bytes_list=1 | float_list=2 | int64_list=3}``, lists with ``repeated value=1`` (packed for float / int64).

Host-side only (numpy); nothing here touches the GPU path.
"""
from __future__ import annotations

import os
import struct
from collections import namedtuple
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Union

import numpy as np

# ------------------------------------------------------------------------------------------------ CRC-32C (Castagnoli)
_POLY = 0x82F63B78
_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ (_POLY if _c & 1 else 0)
    _TABLE.append(_c)
_TABLE_NP = np.array(_TABLE, dtype=np.uint32)


def crc32c(data: bytes) -> int:
    crc = 0xFFFFFFFF
    tab = _TABLE
    for b in data:
        crc = tab[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    crc = crc32c(data)
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ record framing
def write_records(path: str, records: Iterable[bytes]) -> None:
    with open(path, "wb") as f:
        for rec in records:
            head = struct.pack("<Q", len(rec))
            f.write(head)
            f.write(struct.pack("<I", masked_crc32c(head)))
            f.write(rec)
            f.write(struct.pack("<I", masked_crc32c(rec)))


def read_records(path: str, verify_data_crc: bool = False) -> Iterator[bytes]:
    """Yields the payload of every record.  The 12-byte header CRC is always checked; the payload CRC on request (pure-Python
    CRC-32C costs ~1 s per MB)."""
    with open(path, "rb") as f:
        while True:
            head = f.read(8)
            if not head:
                return
            if len(head) != 8:
                raise IOError(f"{path}: truncated record header")
            (hcrc,) = struct.unpack("<I", f.read(4))
            if hcrc != masked_crc32c(head):
                raise IOError(f"{path}: corrupted record length")
            (n,) = struct.unpack("<Q", head)
            data = f.read(n)
            tail = f.read(4)
            if len(data) != n or len(tail) != 4:
                raise IOError(f"{path}: truncated record")
            if verify_data_crc and struct.unpack("<I", tail)[0] != masked_crc32c(data):
                raise IOError(f"{path}: corrupted record payload")
            yield data


# ------------------------------------------------------------------------------------------------ protobuf wire format
def _varint(n: int) -> bytes:
    n &= 0xFFFFFFFFFFFFFFFF          # int64 two's complement
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf: bytes, pos: int):
    shift = val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _ld(field: int, payload: bytes) -> bytes:
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def _fields(buf: bytes):
    """(field number, wire type, value) triples of one message; value is bytes for length-delimited fields."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _read_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 1:
            v, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            v, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            v, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield field, wt, v


FeatureValue = Union[List[bytes], np.ndarray]


def encode_example(features: Dict[str, FeatureValue]) -> bytes:
    """Serialised ``tf.train.Example``.  A list of ``bytes`` becomes a BytesList, an integer array an Int64List, a floating
    array a FloatList (feature order = dict order; TensorFlow sorts map keys on parse, so order is immaterial)."""
    entries = b""
    for name, val in features.items():
        if isinstance(val, (list, tuple)) and all(isinstance(x, (bytes, bytearray)) for x in val):
            lst = b"".join(_ld(1, bytes(x)) for x in val)
            feat = _ld(1, lst)
        else:
            arr = np.asarray(val)
            if arr.dtype.kind in "iu":
                packed = b"".join(_varint(int(x)) for x in arr.reshape(-1))
                feat = _ld(3, _ld(1, packed) if arr.size else b"")
            elif arr.dtype.kind == "f":
                feat = _ld(2, _ld(1, arr.astype("<f4").tobytes()) if arr.size else b"")
            else:
                raise TypeError(f"feature {name}: unsupported value type {type(val)} / dtype {arr.dtype}")
        entries += _ld(1, _ld(1, name.encode("utf-8")) + _ld(2, feat))
    return _ld(1, entries)


def decode_example(buf: bytes) -> Dict[str, FeatureValue]:
    """Inverse of ``encode_example``: BytesList -> list of bytes, Int64List -> int64 array, FloatList -> float32 array."""
    out: Dict[str, FeatureValue] = {}
    for f1, _, features in _fields(buf):
        if f1 != 1:
            continue
        for f2, _, entry in _fields(features):
            if f2 != 1:
                continue
            name, feat = None, b""
            for f3, _, v in _fields(entry):
                if f3 == 1:
                    name = v.decode("utf-8")
                elif f3 == 2:
                    feat = v
            value: FeatureValue = []
            for kind, _, lst in _fields(feat):
                if kind == 1:
                    value = [bytes(v) for f, _, v in _fields(lst) if f == 1]
                elif kind == 2:
                    vals = []
                    for f, wt, v in _fields(lst):
                        if f == 1:
                            vals.append(np.frombuffer(v, dtype="<f4"))       # packed block or a single fixed32
                    value = np.concatenate(vals).astype(np.float32) if vals else np.zeros(0, np.float32)
                elif kind == 3:
                    vals: List[int] = []
                    for f, wt, v in _fields(lst):
                        if f != 1:
                            continue
                        if wt == 0:
                            vals.append(v)
                        else:
                            pos = 0
                            while pos < len(v):
                                x, pos = _read_varint(v, pos)
                                vals.append(x)
                    value = np.array([x - (1 << 64) if x >= (1 << 63) else x for x in vals], dtype=np.int64)
            if name is not None:
                out[name] = value
    return out


# ------------------------------------------------------------------------------------------------ the reference's record types
PreprocessedSourceData = namedtuple("PreprocessedSourceData", ["id", "key", "source", "source_length", "text", "speaker_id",
                                                               "age", "gender", "phone", "phone_length", "phone_txt"])
PreprocessedSourceData.__new__.__defaults__ = (None, None, None)     # the phone fields exist in VCTK records only (vctk/dataset.py:74-76)
PreprocessedMelData = namedtuple("PreprocessedMelData", ["id", "key", "mel", "mel_width", "target_length"])


def _scalar(ex, name, default=None):
    v = ex.get(name)
    if v is None or len(v) == 0:
        if default is None:
            raise KeyError(f"feature '{name}' missing")
        return default
    return v[0]


def decode_source_record(buf: bytes) -> PreprocessedSourceData:
    """datasets/ljspeech/dataset.py:50-72 (+ speaker_id / age / gender / phone, phone_length, phone_txt of
    datasets/vctk/dataset.py:64-98 when present)."""
    ex = decode_example(buf)
    source = np.frombuffer(_scalar(ex, "source"), dtype="<i8").copy()
    spk = ex.get("speaker_id")
    return PreprocessedSourceData(
        id=int(_scalar(ex, "id")), key=_scalar(ex, "key").decode("utf-8"), source=source,
        source_length=int(_scalar(ex, "source_length")), text=_scalar(ex, "text", b"").decode("utf-8"),
        speaker_id=int(spk[0]) if spk is not None and len(spk) else None,
        age=int(ex["age"][0]) if "age" in ex and len(ex["age"]) else None,
        gender=int(ex["gender"][0]) if "gender" in ex and len(ex["gender"]) else None,
        phone=np.frombuffer(_scalar(ex, "phone"), dtype="<i8").copy() if "phone" in ex and len(ex["phone"]) else None,
        phone_length=int(ex["phone_length"][0]) if "phone_length" in ex and len(ex["phone_length"]) else None,
        phone_txt=_scalar(ex, "phone_txt", b"").decode("utf-8") if "phone_txt" in ex and len(ex["phone_txt"]) else None)


def encode_source_record(d: PreprocessedSourceData) -> bytes:
    feats: Dict[str, FeatureValue] = {
        "id": np.array([d.id], np.int64), "key": [d.key.encode("utf-8")],
        "source": [np.asarray(d.source, dtype="<i8").tobytes()], "source_length": np.array([d.source_length], np.int64),
        "text": [d.text.encode("utf-8")]}
    if d.speaker_id is not None:
        feats.update(speaker_id=np.array([d.speaker_id], np.int64), age=np.array([d.age or 0], np.int64),
                     gender=np.array([d.gender or 0], np.int64))
    if d.phone is not None:
        feats.update(phone=[np.asarray(d.phone, dtype="<i8").tobytes()], phone_length=np.array([d.phone_length], np.int64),
                     phone_txt=[(d.phone_txt or "").encode("utf-8")])
    return encode_example(feats)


def decode_mel_record(buf: bytes) -> PreprocessedMelData:
    """utils/tfrecord.py:82-104: mel is raw little-endian float32 [target_length, mel_width]."""
    ex = decode_example(buf)
    width, length = int(_scalar(ex, "mel_width")), int(_scalar(ex, "target_length"))
    mel = np.frombuffer(_scalar(ex, "mel"), dtype="<f4").reshape(length, width).copy()
    return PreprocessedMelData(int(_scalar(ex, "id")), _scalar(ex, "key").decode("utf-8"), mel, width, length)


def encode_mel_record(d: PreprocessedMelData) -> bytes:
    mel = np.asarray(d.mel, dtype="<f4")
    return encode_example({"id": np.array([d.id], np.int64), "key": [d.key.encode("utf-8")], "mel": [mel.tobytes()],
                           "mel_width": np.array([mel.shape[1]], np.int64), "target_length": np.array([mel.shape[0]], np.int64)})


def read_source_file(path: str) -> Iterator[PreprocessedSourceData]:
    for rec in read_records(path):
        yield decode_source_record(rec)


def read_mel_file(path: str) -> Iterator[PreprocessedMelData]:
    for rec in read_records(path):
        yield decode_mel_record(rec)


# ------------------------------------------------------------------------------------------------ predict_mel.py outputs
def write_prediction_result(id_: int, key: str, alignments: Sequence[np.ndarray], mel: np.ndarray, ground_truth_mel: np.ndarray,
                            text: str, source: np.ndarray, accent_type: Optional[np.ndarray], filename: str) -> None:
    """utils/tfrecord.py:135-152, field for field."""
    mel = np.asarray(mel, dtype="<f4")
    gt = np.asarray(ground_truth_mel, dtype="<f4")
    ex = encode_example({
        "id": np.array([id_], np.int64), "key": [key.encode("utf-8")], "mel": [mel.tobytes()],
        "mel_length": np.array([mel.shape[0]], np.int64), "mel_width": np.array([mel.shape[1]], np.int64),
        "ground_truth_mel": [gt.tobytes()], "ground_truth_mel_length": np.array([gt.shape[0]], np.int64),
        "alignment": [np.asarray(a, dtype="<f4").tobytes() for a in alignments], "text": [text.encode("utf-8")],
        "source": [np.asarray(source, dtype="<i8").tobytes()], "source_length": np.array([np.asarray(source).shape[0]], np.int64),
        "accent_type": [np.asarray(accent_type).tobytes()] if accent_type is not None else []})
    write_records(filename, [ex])


def write_mel_file(mel: np.ndarray, path: str) -> None:
    """``mel.tofile(path, format='<f4')`` of predict_mel.py:61 — raw little-endian float32 frames (the ``.mfbsp`` file)."""
    np.asarray(mel, dtype="<f4").tofile(path)


def write_predictions(predictions: Iterable[dict], output_dir: str, extension: str = "mfbsp") -> List[str]:
    """The output loop of predict_mel.py:56-74 (without the plots): for every utterance of every prediction batch write
    ``<key>.<extension>`` (raw mel) and ``<key>.tfrecord`` (PredictionResult)."""
    os.makedirs(output_dir, exist_ok=True)
    written = []
    for p in predictions:
        mel = p["mel"].detach().float().cpu().numpy() if hasattr(p["mel"], "detach") else np.asarray(p["mel"])
        B = mel.shape[0]
        to_np = lambda x: x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)   # noqa: E731
        aligns = [to_np(p[k]) for k in ("alignment", "alignment2", "alignment3", "alignment4", "alignment5", "alignment6")
                  if p.get(k) is not None]
        gt = to_np(p["ground_truth_mel"]) if p.get("ground_truth_mel") is not None else np.zeros((B, 0, mel.shape[2]), np.float32)
        src, ids = to_np(p["source"]), to_np(p["id"])
        for b in range(B):
            key = p["key"][b]
            key = key.decode("utf-8") if isinstance(key, bytes) else str(key)
            text = p["text"][b] if p.get("text") is not None else ""
            text = text.decode("utf-8") if isinstance(text, bytes) else str(text)
            write_mel_file(mel[b], os.path.join(output_dir, f"{key}.{extension}"))
            write_prediction_result(int(ids[b]), key, [a[b] for a in aligns], mel[b], gt[b], text, src[b], None,
                                    os.path.join(output_dir, f"{key}.tfrecord"))
            written.append(key)
    return written
