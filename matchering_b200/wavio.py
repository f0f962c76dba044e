"""Minimal RIFF/WAVE reader and writer (PCM 16/24/32, IEEE float 32/64) on numpy.

The reference delegates file I/O to libsndfile (matchering/loader.py:35, saver.py:32), which is
not in this image; file I/O is outside the accelerated hot path."""
from __future__ import annotations

import struct

import numpy as np

_PCM, _FLOAT, _EXTENSIBLE = 1, 3, 0xFFFE


def read(path: str):
    """-> (float64 array (frames, channels), sample_rate).  Raises RuntimeError on a bad file."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise RuntimeError("Format not recognised")
    pos, fmt, payload = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = body
        elif cid == b"data":
            payload = body
        pos += 8 + size + (size & 1)
    if fmt is None or payload is None or len(fmt) < 16:
        raise RuntimeError("Format not recognised")
    tag, channels, rate, _, _, bits = struct.unpack("<HHIIHH", fmt[:16])
    if tag == _EXTENSIBLE and len(fmt) >= 26:
        tag = struct.unpack("<H", fmt[24:26])[0]
    if channels < 1:
        raise RuntimeError("Format not recognised")
    if tag == _PCM and bits == 16:
        x = np.frombuffer(payload, dtype="<i2").astype(np.float64) / 32768.0
    elif tag == _PCM and bits == 24:
        raw = np.frombuffer(payload[: len(payload) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = raw[:, 0] | (raw[:, 1] << 8) | (raw[:, 2] << 16)
        v = np.where(v >= 1 << 23, v - (1 << 24), v)
        x = v.astype(np.float64) / 8388608.0
    elif tag == _PCM and bits == 32:
        x = np.frombuffer(payload, dtype="<i4").astype(np.float64) / 2147483648.0
    elif tag == _FLOAT and bits == 32:
        x = np.frombuffer(payload, dtype="<f4").astype(np.float64)
    elif tag == _FLOAT and bits == 64:
        x = np.frombuffer(payload, dtype="<f8").astype(np.float64)
    else:
        raise RuntimeError("unknown format")
    frames = len(x) // channels
    return np.ascontiguousarray(x[: frames * channels].reshape(frames, channels)), int(rate)


def write(path: str, array: np.ndarray, sample_rate: int, subtype: str) -> None:
    """Float samples in [-1, 1] -> WAV.  Integer subtypes scale by 2^(bits-1)-1, round to nearest
    (ties to even) and clip, which is libsndfile's float->int conversion."""
    a = np.asarray(array, dtype=np.float64)
    if a.ndim == 1:
        a = a[:, None]
    channels = a.shape[1]
    if subtype in ("PCM_16", "PCM_24", "PCM_32"):
        bits = int(subtype[4:])
        top = float((1 << (bits - 1)) - 1)
        q = np.clip(np.rint(a * top), -top - 1, top).astype(np.int64)
        if bits == 16:
            payload = q.astype("<i2").tobytes()
        elif bits == 32:
            payload = q.astype("<i4").tobytes()
        else:
            u = (q & 0xFFFFFF).astype(np.uint32).reshape(-1)
            payload = np.stack([u & 0xFF, (u >> 8) & 0xFF, (u >> 16) & 0xFF], axis=1).astype(np.uint8).tobytes()
        tag = _PCM
    elif subtype == "FLOAT":
        bits, tag, payload = 32, _FLOAT, a.astype("<f4").tobytes()
    elif subtype == "DOUBLE":
        bits, tag, payload = 64, _FLOAT, a.astype("<f8").tobytes()
    else:
        raise TypeError(f"WAV format does not have {subtype} subtype")
    block = channels * bits // 8
    fmt = struct.pack("<HHIIHH", tag, channels, int(sample_rate), int(sample_rate) * block, block, bits)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"data" + struct.pack("<I", len(payload)) + payload
    if len(payload) & 1:
        body += b"\x00"
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", len(body)) + body)


def write_pcm(path: str, pcm: np.ndarray, sample_rate: int, bits: int, channels: int = 2) -> None:
    """Already-quantised samples (int16 (frames, ch) or packed 24-bit uint8 (frames, 3*ch)) -> WAV."""
    payload = np.ascontiguousarray(pcm).tobytes()
    block = channels * bits // 8
    fmt = struct.pack("<HHIIHH", _PCM, channels, int(sample_rate), int(sample_rate) * block, block, bits)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"data" + struct.pack("<I", len(payload)) + payload
    if len(payload) & 1:
        body += b"\x00"
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", len(body)) + body)


def read_pcm(path: str):
    """Raw 16- or 24-bit PCM of a WAV file without converting it on the host:
    -> (samples, sample_rate, channels, bits) with samples int16 (frames, ch) or packed 24-bit uint8
    (frames, 3*ch); None when the file is not 16/24-bit PCM WAV."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        return None
    pos, fmt, payload = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        if cid == b"fmt ":
            fmt = data[pos + 8:pos + 8 + size]
        elif cid == b"data":
            payload = memoryview(data)[pos + 8:pos + 8 + size]
        pos += 8 + size + (size & 1)
    if fmt is None or payload is None or len(fmt) < 16:
        return None
    tag, channels, rate, _, _, bits = struct.unpack("<HHIIHH", fmt[:16])
    if tag == _EXTENSIBLE and len(fmt) >= 26:
        tag = struct.unpack("<H", fmt[24:26])[0]
    if tag != _PCM or bits not in (16, 24) or channels < 1:
        return None
    block = channels * bits // 8
    frames = len(payload) // block
    raw = np.frombuffer(payload[: frames * block], dtype=np.int16 if bits == 16 else np.uint8)
    return raw.reshape(frames, channels if bits == 16 else 3 * channels), int(rate), int(channels), int(bits)
