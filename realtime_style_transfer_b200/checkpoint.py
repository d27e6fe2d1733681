"""TF2 object-based checkpoint ingestion without TensorFlow.

The reference writes its weights with ``tf.train.Checkpoint`` / ``model.save_weights`` (tracing/checkpoint.py:18-37) and
restores them with ``load_weights(path).assert_nontrivial_match()`` (predict_using_checkpoint.py:84-85).  A TF2 checkpoint
is a *tensor bundle*:

  <prefix>.index                 LevelDB-style SSTable: key -> BundleEntryProto (dtype, shape, shard, offset, size, crc32c);
                                 the empty key holds the BundleHeaderProto
  <prefix>.data-00000-of-00001   the raw little-endian tensor bytes
  key ``_CHECKPOINTABLE_OBJECT_GRAPH``  a string tensor with the TrackableObjectGraph proto: per variable its ``full_name``
                                 (e.g. ``contract_start_conv/kernel``) and its ``checkpoint_key``
                                 (e.g. ``layer_with_weights-3/.../kernel/.ATTRIBUTES/VARIABLE_VALUE``)

Both the SSTable and the two protos are parsed by hand here (a few dozen lines each; TF's bundle writer never compresses).
``write_checkpoint`` produces the same container: every tensor carries its masked crc32c (``rst_host_crc32c``), the object
graph has children edges (root -> ``layer_with_weights-i`` -> ``v``) that spell the keys and the Keras ``full_name`` of every
variable.  What it can NOT reproduce without Keras is the reference model's own trackable object tree (nested functional
models), so a bundle written here is meant for this package's ``load_weights`` and for name-based readers
(``tf.train.load_checkpoint(prefix).get_tensor(key)``), not for Keras' object-based ``model.load_weights``.
No checkpoint of the reference is available offline: parity with a real TensorFlow-written file is unpinned;
the reader is additionally tested against a bundle assembled by an independent encoder (tests/test_checkpoint.py).
"""
from __future__ import annotations

import logging
import os
import re
import struct
from typing import Dict, List, Tuple

import numpy as np
log = logging.getLogger(__name__)


_TABLE_MAGIC = 0xDB4775248B80FB57
_DT_FLOAT, _DT_STRING = 1, 7
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_}
_OBJECT_GRAPH_KEY = "_CHECKPOINTABLE_OBJECT_GRAPH"
_VALUE_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"


# ---- varints / protobuf wire format -----------------------------------------------------------------------------
def _get_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_fields(buf: bytes) -> List[Tuple[int, int, object]]:
    """Flat protobuf decode: [(field_number, wire_type, value)]; length-delimited values stay bytes."""
    out, pos = [], 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        fn, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        out.append((fn, wt, v))
    return out


def _field(fn: int, wt: int, payload: bytes) -> bytes:
    return _put_varint((fn << 3) | wt) + payload


def _len_delimited(payload: bytes) -> bytes:
    return _put_varint(len(payload)) + payload


# ---- crc32c (Castagnoli), masked as LevelDB / TF do -----------------------------------------------------------------
_CRC_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ 0x82F63B78 if _c & 1 else _c >> 1
    _CRC_TABLE.append(_c)


def _crc32c_python(data: bytes) -> int:
    c = 0xFFFFFFFF
    for b in data:
        c = _CRC_TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def crc32c(data) -> int:
    """bytes / bytearray / memoryview / contiguous ndarray.  Large buffers go through the library's slicing-by-8 routine
    (rst_host_crc32c, host code: needs no GPU); the table loop above is the fallback and the cross-check."""
    if isinstance(data, np.ndarray):
        data = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    n = len(data)
    if n >= 4096:
        try:
            from ._native import load_library
            import ctypes as C
            lib = load_library()
            buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
            return int(lib.rst_host_crc32c(buf.ctypes.data_as(C.c_void_p), n))
        except Exception:                               # noqa: BLE001 - library not built: fall back to the Python loop
            pass
    return _crc32c_python(bytes(data))


def _mask(crc: int) -> int:
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---- SSTable ----------------------------------------------------------------------------------------------------
def _read_block(data: bytes, offset: int, size: int, verify: bool = True) -> List[Tuple[bytes, bytes]]:
    block = data[offset:offset + size]
    ctype = data[offset + size]
    if ctype != 0:
        raise ValueError("compressed SSTable block: TF tensor bundles are written uncompressed")
    if verify:
        stored = struct.unpack_from("<I", data, offset + size + 1)[0]
        if _mask(crc32c(block + bytes([ctype]))) != stored:
            raise ValueError("checkpoint index block fails its crc32c check")
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    entries, pos, key = [], 0, b""
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        entries.append((key, block[pos:pos + vlen]))
        pos += vlen
    return entries


def _read_table(path: str) -> Dict[bytes, bytes]:
    data = open(path, "rb").read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != _TABLE_MAGIC:
        raise ValueError(f"{path} is not a TensorFlow checkpoint index (bad SSTable magic)")
    footer = data[-48:]
    _mo, pos = _get_varint(footer, 0)
    _ms, pos = _get_varint(footer, pos)
    io, pos = _get_varint(footer, pos)
    isz, pos = _get_varint(footer, pos)
    table = {}
    for _sep, handle in _read_block(data, io, isz):
        bo, p2 = _get_varint(handle, 0)
        bs, _ = _get_varint(handle, p2)
        for k, v in _read_block(data, bo, bs):
            table[k] = v
    return table


def _parse_shape(buf: bytes) -> Tuple[int, ...]:
    dims = []
    for fn, _wt, v in _parse_fields(buf):
        if fn == 2:
            size = 0
            for f2, _w2, v2 in _parse_fields(v):
                if f2 == 1:
                    size = v2
            dims.append(int(size))
    return tuple(dims)


def read_index(prefix: str) -> Dict[str, dict]:
    """key -> {dtype, shape, shard, offset, size} for every tensor of the bundle."""
    table = _read_table(prefix + ".index")
    out = {}
    for k, v in table.items():
        if k == b"":
            continue                                    # BundleHeaderProto
        e = {"dtype": 0, "shape": (), "shard": 0, "offset": 0, "size": 0, "crc32c": 0}
        for fn, _wt, val in _parse_fields(v):
            if fn == 1:
                e["dtype"] = val
            elif fn == 2:
                e["shape"] = _parse_shape(val)
            elif fn == 3:
                e["shard"] = val
            elif fn == 4:
                e["offset"] = val
            elif fn == 5:
                e["size"] = val
            elif fn == 6:
                e["crc32c"] = val
        out[k.decode()] = e
    return out


def _parse_object_graph(buf: bytes) -> Dict[str, str]:
    """TrackableObjectGraph -> {checkpoint_key: full_name} for every serialized variable."""
    out = {}
    for fn, _wt, node in _parse_fields(buf):
        if fn != 1:
            continue
        for f2, _w2, attr in _parse_fields(node):
            if f2 != 2:
                continue
            name = full = key = ""
            for f3, _w3, v in _parse_fields(attr):
                if f3 == 1:
                    name = v.decode()
                elif f3 == 2:
                    full = v.decode()
                elif f3 == 3:
                    key = v.decode()
            if name == "VARIABLE_VALUE" and key:
                out[key] = full
    return out


def read_checkpoint_variables(prefix: str, verify_tensors: bool = False) -> Dict[str, dict]:
    """Reads every numeric variable of a TF2 checkpoint: checkpoint_key -> {'value': ndarray, 'full_name': str}.
    verify_tensors: also check every tensor against the masked crc32c of its BundleEntryProto (as TensorFlow's reader does)."""
    prefix = str(prefix)
    if prefix.endswith(".index"):
        prefix = prefix[:-6]
    if os.path.isdir(prefix) or os.path.basename(prefix) == "checkpoint":
        # a directory or the CheckpointManager state file: follow model_checkpoint_path
        state = prefix if os.path.basename(prefix) == "checkpoint" else os.path.join(prefix, "checkpoint")
        if not os.path.exists(state):
            raise FileNotFoundError(f"{prefix}: no checkpoint state file")
        m = re.search(r'model_checkpoint_path:\s*"([^"]+)"', open(state).read())
        if not m:
            raise ValueError(f"{state}: no model_checkpoint_path")
        prefix = os.path.join(os.path.dirname(state), m.group(1))
    if not os.path.exists(prefix + ".index"):
        raise FileNotFoundError(f"{prefix}.index not found")
    index = read_index(prefix)
    shards = {}

    def shard_bytes(i):
        if i not in shards:
            n = max(e["shard"] for e in index.values()) + 1
            shards[i] = np.memmap(f"{prefix}.data-{i:05d}-of-{n:05d}", dtype=np.uint8, mode="r")
        return shards[i]

    names = {}
    if _OBJECT_GRAPH_KEY in index:
        e = index[_OBJECT_GRAPH_KEY]
        raw = bytes(shard_bytes(e["shard"])[e["offset"]:e["offset"] + e["size"]])
        n, pos = _get_varint(raw, 0)                    # scalar string tensor: [varint length][4-byte crc of lengths][bytes]
        names = _parse_object_graph(raw[pos + 4:pos + 4 + n])
    out = {}
    for key, e in index.items():
        if e["dtype"] not in _DTYPES:
            continue
        dt = np.dtype(_DTYPES[e["dtype"]])
        buf = shard_bytes(e["shard"])[e["offset"]:e["offset"] + e["size"]]
        value = np.frombuffer(bytes(buf), dtype=dt).reshape(e["shape"]).copy()
        if verify_tensors and _mask(crc32c(value)) != e["crc32c"]:
            raise ValueError(f"checkpoint tensor {key} fails its crc32c check")
        out[key] = {"value": value, "full_name": names.get(key, "")}
    return out


# ---- name mapping -----------------------------------------------------------------------------------------------------
def _keras_to_ours(full_name: str):
    """Keras variable name (tf.Variable.name without ':0') -> our registry name, or None."""
    n = full_name
    m = re.fullmatch(r"contract_(\w+)_conv/(kernel|bias)", n)
    if m:
        return f"contract_{m.group(1)}/conv/{m.group(2)}"
    m = re.fullmatch(r"residual_block_(\d)_conv(\d)/(kernel|bias)", n)
    if m:
        return f"residual_block_{m.group(1)}/conv{m.group(2)}/{m.group(3)}"
    m = re.fullmatch(r"expand_(\w+)_conv/(kernel|bias)", n)
    if m:
        return f"expand_{m.group(1)}/conv/{m.group(2)}"
    if re.fullmatch(r"(StylePredictor|StyleNormPredictor|dummy_conv)/(kernel|bias)", n):
        return n
    if re.match(r"(Conv(_1)?|expanded_conv(_\d+)?)/", n):
        return "mobilenet/" + n
    if re.fullmatch(r"block\d_conv\d/(kernel|bias)", n):
        return n
    return None


def match_checkpoint_to_model(ckpt_vars: Dict[str, dict], model):
    """-> (assignment {our_name: array}, missing [our names], unused [checkpoint keys]).

    Variables are matched by their Keras ``full_name`` from the object graph.  The unnamed BatchNormalization layers of the
    contract blocks (``batch_normalization[_k]/gamma`` ...) are matched in numeric order to contract_start, contract_0, ...
    (the order in which create_style_transfer_model builds them, styleTransfer.py:224-232)."""
    mine = model._all_variables()
    assignment, used = {}, set()
    source_key: Dict[str, str] = {}
    bn_groups: Dict[int, Dict[str, Tuple[str, np.ndarray]]] = {}
    for key, e in sorted(ckpt_vars.items()):
        if not key.endswith(_VALUE_SUFFIX) or "optimizer" in key.split("/")[0] or "/.OPTIMIZER_SLOT/" in key:
            continue
        full = e["full_name"]
        m = re.fullmatch(r"batch_normalization(?:_(\d+))?/(gamma|beta|moving_mean|moving_variance)", full)
        if m:
            bn_groups.setdefault(int(m.group(1) or 0), {})[m.group(2)] = (key, e["value"])
            continue
        ours = _keras_to_ours(full)
        if ours in mine and tuple(e["value"].shape) == mine[ours].shape:
            if ours in source_key:
                # Two checkpoint entries carry the same Keras full_name: a training checkpoint holds the frozen loss model
                # next to the inference model, and StyleLossModelMobileNet shares MobileNetV3's layer names with the predictor.
                # The object-graph path disambiguates: entries below `loss_model` never belong to the inference variables.
                old_is_loss, new_is_loss = "loss_model" in source_key[ours], "loss_model" in key
                if new_is_loss and not old_is_loss:
                    continue
                if not (old_is_loss and not new_is_loss):
                    log.warning("checkpoint entries %s and %s both name %s; keeping the first", source_key[ours], key, full)
                    continue
                used.discard(source_key[ours])
            assignment[ours] = e["value"].astype(np.float32)
            source_key[ours] = key
            used.add(key)
    contract_names = sorted({k.split("/")[0] for k in mine if k.startswith("contract_") and "/bn/" in k},
                            key=lambda s: (-1 if s.endswith("start") else int(s.split("_")[1])))
    for target, idx in zip(contract_names, sorted(bn_groups)):
        for var, (key, value) in bn_groups[idx].items():
            ours = f"{target}/bn/{var}"
            if ours in mine and tuple(value.shape) == mine[ours].shape:
                assignment[ours] = value.astype(np.float32)
                used.add(key)
    missing = [k for k in mine if k not in assignment]
    unused = [k for k in ckpt_vars if k not in used and k.endswith(_VALUE_SUFFIX)]
    return assignment, missing, unused


# ---- writer (tests, and save_weights(..., save_format='tf')) ------------------------------------------------------------------
def _ours_to_keras(name: str, bn_index: Dict[str, int]) -> str:
    m = re.fullmatch(r"(contract_\w+)/conv/(kernel|bias)", name)
    if m:
        return f"{m.group(1)}_conv/{m.group(2)}"
    m = re.fullmatch(r"(contract_\w+)/bn/(\w+)", name)
    if m:
        i = bn_index.setdefault(m.group(1), len(bn_index))
        return f"batch_normalization{'_%d' % i if i else ''}/{m.group(2)}"
    m = re.fullmatch(r"(residual_block_\d)/conv(\d)/(kernel|bias)", name)
    if m:
        return f"{m.group(1)}_conv{m.group(2)}/{m.group(3)}"
    m = re.fullmatch(r"(expand_\w+)/conv/(kernel|bias)", name)
    if m:
        return f"{m.group(1)}_conv/{m.group(2)}"
    if name.startswith("mobilenet/"):
        return name[len("mobilenet/"):]
    return name


def _build_block(entries: List[Tuple[bytes, bytes]]) -> bytes:
    out, last = bytearray(), b""
    restarts = []
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % 16 == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts.append(0)
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_checkpoint(prefix: str, variables: Dict[str, np.ndarray], extra: Dict[str, np.ndarray] = None) -> str:
    """Writes ``variables`` (our registry names) as a TF2 object-based checkpoint (index + one data shard + state file).
    ``extra``: additional checkpoint_key -> array entries (e.g. optimizer slots) without object-graph names."""
    prefix = str(prefix)
    os.makedirs(os.path.dirname(prefix) or ".", exist_ok=True)
    bn_index: Dict[str, int] = {}
    data = bytearray()
    entries: Dict[bytes, bytes] = {}
    graph_nodes = [b""]                                         # node 0: the root object, its children are filled in below
    root_children: List[bytes] = []

    def add_tensor(key: str, arr: np.ndarray, dtype_code: int):
        raw = arr.tobytes()
        shape = b"".join(_field(2, 2, _put_varint(len(d)) + d) for d in
                         (_field(1, 0, _put_varint(int(s))) for s in arr.shape))
        entry = (_field(1, 0, _put_varint(dtype_code)) + _field(2, 2, _put_varint(len(shape)) + shape) +
                 _field(4, 0, _put_varint(len(data))) + _field(5, 0, _put_varint(len(raw))) +
                 _field(6, 5, struct.pack("<I", _mask(crc32c(raw)))))         # masked crc32c of the tensor bytes, every tensor
        entries[key.encode()] = entry
        data.extend(raw)

    for i, (name, arr) in enumerate(variables.items()):
        key = f"layer_with_weights-{i}/v{_VALUE_SUFFIX}"
        full = _ours_to_keras(name, bn_index)
        add_tensor(key, np.ascontiguousarray(arr, np.float32), _DT_FLOAT)
        attr = (_field(1, 2, _put_varint(14) + b"VARIABLE_VALUE") + _field(2, 2, _put_varint(len(full)) + full.encode()) +
                _field(3, 2, _put_varint(len(key)) + key.encode()))
        # object graph: root --layer_with_weights-i--> holder --v--> variable node (children edges as TensorFlow writes them,
        # so the key is the path of edge names from the root followed by the attribute suffix)
        holder, variable = len(graph_nodes), len(graph_nodes) + 1
        root_children.append(_field(1, 2, _len_delimited(_field(1, 0, _put_varint(holder)) + _field(2, 2, _len_delimited(f"layer_with_weights-{i}".encode())))))
        graph_nodes.append(_field(1, 2, _len_delimited(_field(1, 0, _put_varint(variable)) + _field(2, 2, _len_delimited(b"v")))))
        graph_nodes.append(_field(2, 2, _put_varint(len(attr)) + attr))
    for key, arr in (extra or {}).items():
        add_tensor(key, np.ascontiguousarray(arr, np.float32), _DT_FLOAT)
    graph_nodes[0] = b"".join(root_children)
    graph = b"".join(_field(1, 2, _put_varint(len(n)) + n) for n in graph_nodes)
    # scalar DT_STRING tensor: [varint length][masked crc32c of the length bytes][bytes]
    length = _put_varint(len(graph))
    raw = length + struct.pack("<I", _mask(crc32c(struct.pack("<Q", len(graph))))) + graph
    entries[_OBJECT_GRAPH_KEY.encode()] = (_field(1, 0, _put_varint(_DT_STRING)) + _field(2, 2, _put_varint(0)) +
                                           _field(4, 0, _put_varint(len(data))) + _field(5, 0, _put_varint(len(raw))) +
                                           _field(6, 5, struct.pack("<I", 0)))
    data.extend(raw)
    header = _field(1, 0, _put_varint(1)) + _field(3, 2, _put_varint(2) + _field(1, 0, _put_varint(1)))   # num_shards=1, version{producer=1}
    entries[b""] = header

    out = bytearray()

    def emit(block: bytes) -> bytes:
        off = len(out)
        out.extend(block + b"\x00" + struct.pack("<I", _mask(crc32c(block + b"\x00"))))
        return _put_varint(off) + _put_varint(len(block))

    keys = sorted(entries)
    index_entries = []
    for s in range(0, len(keys), 64):                            # several data blocks, like a real table
        chunk = keys[s:s + 64]
        handle = emit(_build_block([(k, entries[k]) for k in chunk]))
        index_entries.append((chunk[-1], handle))
    meta = emit(_build_block([]))
    idx = emit(_build_block(index_entries))
    footer = meta + idx
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", _TABLE_MAGIC))
    with open(prefix + ".index", "wb") as f:
        f.write(out)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(data)
    with open(os.path.join(os.path.dirname(prefix) or ".", "checkpoint"), "w") as f:
        base = os.path.basename(prefix)
        f.write(f'model_checkpoint_path: "{base}"\nall_model_checkpoint_paths: "{base}"\n')
    return prefix
