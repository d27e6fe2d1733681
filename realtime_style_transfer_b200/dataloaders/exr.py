"""Minimal OpenEXR reader / writer for the Unreal HDR screenshot planes.

The reference reads its G-buffer planes with the third-party ``pyroexr`` module
(realtime_style_transfer/dataloaders/hdrScreenshots.py:14-29), which is not available offline.  This is a from-the-spec
implementation of the subset those files use: single-part scan-line images, channels of type HALF / FLOAT / UINT without
sub-sampling, compression NONE, ZIPS (1 scan line per block) or ZIP (16 scan lines per block).  PIZ / PXR24 / B44 / DWA and
tiled or multi-part files raise ``NotImplementedError``.

File layout (OpenEXR "technical introduction"): magic 20000630, version word, header = list of
(name\\0, type\\0, int32 size, value) terminated by an empty name, then one uint64 offset per block, then the blocks
(int32 y, int32 byte count, payload).  Inside a block the scan lines are stored one after the other, and inside a scan line
the channels (sorted by name) one after the other.  ZIP payloads are zlib streams of the block after two reversible
filters: the bytes are split into even / odd halves, then delta-encoded.
"""
from __future__ import annotations

import struct
import zlib
from pathlib import Path
from typing import Dict, Tuple

import numpy as np

MAGIC = 20000630
PIXEL_TYPES = {0: np.dtype("<u4"), 1: np.dtype("<f2"), 2: np.dtype("<f4")}
COMPRESSION_NAMES = {0: "NONE", 1: "RLE", 2: "ZIPS", 3: "ZIP", 4: "PIZ", 5: "PXR24", 6: "B44", 7: "B44A", 8: "DWAA", 9: "DWAB"}
LINES_PER_BLOCK = {0: 1, 2: 1, 3: 16}


class ExrImage:
    """Decoded image: ``channel(name)`` -> float32 (H, W) array, ``channels()`` -> dict in file (alphabetical) order."""

    def __init__(self, planes: Dict[str, np.ndarray], header: dict):
        self._planes = planes
        self.header = header

    def channel(self, name: str) -> np.ndarray:
        return self._planes[name]

    def channels(self) -> Dict[str, np.ndarray]:
        return dict(self._planes)

    @property
    def shape(self) -> Tuple[int, int]:
        first = next(iter(self._planes.values()))
        return first.shape


def _read_cstr(buf: bytes, pos: int) -> Tuple[str, int]:
    end = buf.index(b"\0", pos)
    return buf[pos:end].decode("latin-1"), end + 1


def _unzip_block(payload: bytes, expected: int) -> bytes:
    raw = np.frombuffer(zlib.decompress(payload), dtype=np.uint8)
    if raw.size != expected:
        raise ValueError(f"EXR block inflates to {raw.size} bytes, expected {expected}")
    # undo the delta predictor: t[i] = t[i-1] + enc[i] - 128 (mod 256)
    t = raw
    if raw.size:
        t = (np.cumsum(np.concatenate(([int(raw[0])], raw[1:].astype(np.int64) - 128))) & 0xFF).astype(np.uint8)
    # undo the even/odd interleave: first half holds the even bytes, second half the odd ones
    half = (expected + 1) // 2
    out = np.empty(expected, np.uint8)
    out[0::2] = t[:half]
    out[1::2] = t[half:]
    return out.tobytes()


def _zip_block(data: bytes) -> bytes:
    raw = np.frombuffer(data, dtype=np.uint8)
    t = np.concatenate((raw[0::2], raw[1::2]))
    d = t.astype(np.int64)
    enc = np.empty_like(d)
    if d.size:
        enc[0] = d[0]
        enc[1:] = (d[1:] - d[:-1] + 128 + 256) & 0xFF
    return zlib.compress(enc.astype(np.uint8).tobytes())


def load(path, keep_half: bool = False) -> ExrImage:
    """keep_half: HALF channels are returned as float16 planes instead of being widened to float32 (pyroexr's behaviour, and the
    default here): the values are identical, and a float16 G-buffer crosses PCIe in half the bytes (rst_transfer_*_typed)."""
    buf = Path(path).read_bytes()
    magic, version = struct.unpack_from("<iI", buf, 0)
    if magic != MAGIC:
        raise ValueError(f"{path}: not an OpenEXR file")
    if version & 0x200 or version & 0x800 or version & 0x1000:
        raise NotImplementedError(f"{path}: tiled / deep / multi-part EXR files are not supported")
    pos = 8
    header = {}
    while True:
        name, pos = _read_cstr(buf, pos)
        if not name:
            break
        typ, pos = _read_cstr(buf, pos)
        (size,) = struct.unpack_from("<i", buf, pos)
        pos += 4
        header[name] = (typ, buf[pos:pos + size])
        pos += size
    chans = []
    cbuf = header["channels"][1]
    cpos = 0
    while cbuf[cpos] != 0:
        cname, cpos = _read_cstr(cbuf, cpos)
        ptype, _plinear, xs, ys = struct.unpack_from("<iB3xii", cbuf, cpos)
        cpos += 16
        if xs != 1 or ys != 1:
            raise NotImplementedError(f"{path}: sub-sampled channel {cname}")
        chans.append((cname, PIXEL_TYPES[ptype]))
    compression = header["compression"][1][0]
    if compression not in LINES_PER_BLOCK:
        raise NotImplementedError(f"{path}: {COMPRESSION_NAMES.get(compression, compression)} compression is not supported "
                                  "(NONE, ZIPS and ZIP are)")
    x0, y0, x1, y1 = struct.unpack("<iiii", header["dataWindow"][1])
    width, height = x1 - x0 + 1, y1 - y0 + 1
    lines = LINES_PER_BLOCK[compression]
    nblocks = (height + lines - 1) // lines
    offsets = struct.unpack_from(f"<{nblocks}Q", buf, pos)
    line_bytes = sum(dt.itemsize for _, dt in chans) * width
    planes = {name: np.empty((height, width), np.float16 if (keep_half and dt == np.dtype("<f2")) else np.float32)
              for name, dt in chans}
    for off in offsets:
        y, nbytes = struct.unpack_from("<ii", buf, off)
        payload = buf[off + 8:off + 8 + nbytes]
        rows = min(lines, y1 - y + 1)
        expected = rows * line_bytes
        if compression != 0 and nbytes < expected:
            payload = _unzip_block(payload, expected)
        p = 0
        for r in range(rows):
            for name, dt in chans:
                n = width * dt.itemsize
                planes[name][y - y0 + r] = np.frombuffer(payload, dtype=dt, count=width, offset=p)
                p += n
    info = {"channels": [c for c, _ in chans], "compression": COMPRESSION_NAMES[compression], "dataWindow": (x0, y0, x1, y1)}
    return ExrImage(planes, info)


def save(path, planes: Dict[str, np.ndarray], compression: str = "ZIP", pixel_type: str = "HALF") -> None:
    """Writes a single-part scan-line file; used by the tests and to produce fixtures the reference's loader can read."""
    comp = {v: k for k, v in COMPRESSION_NAMES.items()}[compression]
    if comp not in LINES_PER_BLOCK:
        raise NotImplementedError(compression)
    ptype = {"UINT": 0, "HALF": 1, "FLOAT": 2}[pixel_type]
    dt = PIXEL_TYPES[ptype]
    names = sorted(planes)
    height, width = planes[names[0]].shape

    def attr(name, typ, value):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(value)) + value

    chlist = b"".join(n.encode() + b"\0" + struct.pack("<iB3xii", ptype, 0, 1, 1) for n in names) + b"\0"
    window = struct.pack("<iiii", 0, 0, width - 1, height - 1)
    head = struct.pack("<iI", MAGIC, 2)
    head += attr("channels", "chlist", chlist) + attr("compression", "compression", bytes([comp]))
    head += attr("dataWindow", "box2i", window) + attr("displayWindow", "box2i", window)
    head += attr("lineOrder", "lineOrder", b"\0") + attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
    head += attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0))
    head += attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + b"\0"
    lines = LINES_PER_BLOCK[comp]
    blocks = []
    for y in range(0, height, lines):
        raw = b"".join(np.ascontiguousarray(planes[n][yy], dtype=np.float32).astype(dt).tobytes()
                       for yy in range(y, min(y + lines, height)) for n in names)
        payload = raw
        if comp != 0:
            z = _zip_block(raw)
            payload = z if len(z) < len(raw) else raw          # the format stores the raw bytes when compression does not help
        blocks.append(struct.pack("<ii", y, len(payload)) + payload)
    table_pos = len(head)
    pos = table_pos + 8 * len(blocks)
    offsets = []
    for b in blocks:
        offsets.append(pos)
        pos += len(b)
    Path(path).write_bytes(head + struct.pack(f"<{len(offsets)}Q", *offsets) + b"".join(blocks))
