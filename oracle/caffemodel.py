"""Minimal protobuf-wire reader for BVLC Caffe ``*.caffemodel`` files (oracle side).

Test infrastructure -- see ``oracle/__init__.py``.

Format facts (BVLC ``caffe.proto``):
``NetParameter``: field 1 = name, field 100 = repeated ``LayerParameter``.
``LayerParameter``: 1 name, 2 type, 3 bottom, 4 top, 7 = repeated ``BlobProto``.
``BlobProto``: 7 = ``BlobShape{1: packed int64 dim}``, 5 = packed float32 data,
legacy 1..4 = num/channels/height/width.
The reference loads these files at ``retinaface/RetinaFace.cpp:276,311-312``.
"""
from __future__ import annotations

import struct
from typing import Dict, Iterator, List, Tuple

import numpy as np


def _varint(buf: memoryview, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not (b & 0x80):
            return result, pos
        shift += 7


def _fields(buf: memoryview) -> Iterator[Tuple[int, int, object]]:
    """Yield (field_number, wire_type, value) over one message body."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = bytes(buf[pos:pos + 8])
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            val = bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt} at {pos}")
        yield fno, wt, val


def _parse_blob(buf: memoryview) -> np.ndarray:
    dims: List[int] = []
    legacy = {}
    data = None
    loose: List[float] = []
    for fno, wt, val in _fields(buf):
        if fno == 7 and wt == 2:  # BlobShape
            for f2, w2, v2 in _fields(val):
                if f2 == 1 and w2 == 2:  # packed
                    p = 0
                    while p < len(v2):
                        d, p = _varint(v2, p)
                        dims.append(d)
                elif f2 == 1 and w2 == 0:
                    dims.append(int(v2))
        elif fno == 5 and wt == 2:  # packed float data
            data = np.frombuffer(bytes(val), dtype="<f4").copy()
        elif fno == 5 and wt == 5:  # unpacked float
            loose.append(struct.unpack("<f", val)[0])
        elif fno in (1, 2, 3, 4) and wt == 0:
            legacy[fno] = int(val)
    if data is None:
        data = np.asarray(loose, dtype=np.float32)
    if not dims and legacy:
        dims = [legacy.get(i, 1) for i in (1, 2, 3, 4)]
    if dims:
        data = data.reshape(dims)
    return data


def load_caffemodel(path: str) -> Dict[str, dict]:
    """Return {layer_name: {"type": str, "blobs": [ndarray, ...]}} in file order."""
    raw = memoryview(open(path, "rb").read())
    layers: Dict[str, dict] = {}
    for fno, wt, val in _fields(raw):
        if fno != 100 or wt != 2:
            continue
        name = typ = None
        blobs = []
        for f2, w2, v2 in _fields(val):
            if f2 == 1 and w2 == 2:
                name = bytes(v2).decode()
            elif f2 == 2 and w2 == 2:
                typ = bytes(v2).decode()
            elif f2 == 7 and w2 == 2:
                blobs.append(_parse_blob(v2))
        layers[name] = {"type": typ, "blobs": blobs}
    return layers
