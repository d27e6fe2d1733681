"""CPU tests of the TensorFlow-free TF2 checkpoint reader / writer."""
import os
import struct

import numpy as np
import pytest

from realtime_style_transfer_b200 import checkpoint as ck
from realtime_style_transfer_b200.models import stylePrediction, styleTransfer, styleTransferInferenceModel


def test_crc32c_known_answers():
    # RFC 3720 test vectors
    assert ck.crc32c(b"123456789") == 0xE3069283
    assert ck.crc32c(bytes(32)) == 0x8A9136AA
    assert ck.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43


def test_bundle_roundtrip_and_format(tmp_path):
    rng = np.random.default_rng(0)
    variables = {"contract_start/conv/kernel": rng.standard_normal((9, 9, 3, 32)).astype(np.float32),
                 "contract_start/conv/bias": rng.standard_normal(32).astype(np.float32),
                 "contract_start/bn/gamma": rng.standard_normal(32).astype(np.float32),
                 "residual_block_2/conv1/kernel": rng.standard_normal((3, 3, 8, 8)).astype(np.float32),
                 "expand_last/conv/kernel": rng.standard_normal((9, 9, 3, 16)).astype(np.float32),
                 "mobilenet/expanded_conv_3/depthwise/depthwise_kernel": rng.standard_normal((5, 5, 96, 1)).astype(np.float32),
                 "StylePredictor/bias": np.full(100, 0.5, np.float32)}
    extra = {"optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE": np.zeros((), np.float32)}
    prefix = ck.write_checkpoint(str(tmp_path / "ckpt-3"), variables, extra)
    raw = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xDB4775248B80FB57              # LevelDB table magic
    index = ck.read_index(prefix)
    assert "_CHECKPOINTABLE_OBJECT_GRAPH" in index and index["_CHECKPOINTABLE_OBJECT_GRAPH"]["dtype"] == 7
    got = ck.read_checkpoint_variables(prefix)
    by_full = {e["full_name"]: e["value"] for e in got.values() if e["full_name"]}
    np.testing.assert_array_equal(by_full["contract_start_conv/kernel"], variables["contract_start/conv/kernel"])
    np.testing.assert_array_equal(by_full["batch_normalization/gamma"], variables["contract_start/bn/gamma"])
    np.testing.assert_array_equal(by_full["expanded_conv_3/depthwise/depthwise_kernel"],
                                  variables["mobilenet/expanded_conv_3/depthwise/depthwise_kernel"])
    # directory / state-file forms resolve through model_checkpoint_path like tf.train.latest_checkpoint
    assert set(ck.read_checkpoint_variables(str(tmp_path))) == set(got)
    assert set(ck.read_checkpoint_variables(str(tmp_path / "checkpoint"))) == set(got)
    # a flipped byte in the index is caught by the block CRC
    bad = bytearray(raw)
    bad[10] ^= 0xFF
    open(prefix + ".index", "wb").write(bad)
    with pytest.raises(ValueError):
        ck.read_index(prefix)
    with pytest.raises(FileNotFoundError):
        ck.read_checkpoint_variables(str(tmp_path / "nope"))


def test_load_weights_from_tf_checkpoint_matches_by_keras_names(tmp_path):
    def build():
        return styleTransferInferenceModel.make_style_transfer_inference_model(
            num_styles=1,
            style_predictor_factory_func=lambda n: stylePrediction.create_style_prediction_model((64, 128, 3), "MOBILE_NET", n),
            style_transfer_factory_func=lambda: styleTransfer.create_style_transfer_model((64, 128, 17), (64, 128, 3), 16, 32, 1))
    src, dst = build(), build()
    rng = np.random.default_rng(1)
    new = {k: rng.standard_normal(v.shape).astype(np.float32) for k, v in src.inference.weights.items()}
    src.inference.set_weights(new)
    prefix = src.inference.save_weights(str(tmp_path / "weights" / "latest_epoch_weights"))     # tracing/checkpoint.py:37
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    status = dst.inference.load_weights(prefix)
    status.assert_nontrivial_match()
    status.assert_existing_objects_matched()
    for k, v in new.items():
        np.testing.assert_array_equal(dst.inference.weights[k], v)
    # the three contract BatchNorm layers are unnamed in Keras: matched in creation order
    assert np.array_equal(dst.transfer.weights["contract_1/bn/moving_variance"], new["contract_1/bn/moving_variance"])
    # a transfer-only model loading the full checkpoint: optimizer slots / predictor variables stay unused
    t2, _ = styleTransfer.create_style_transfer_model((64, 128, 17), (64, 128, 3), 16, 32, 1)
    st = t2.load_weights(prefix)
    st.assert_nontrivial_match()
    assert st.unused and not st.missing
    with pytest.raises(AssertionError):
        st.assert_consumed()


# ---- a bundle assembled by hand from the format description (NOT through checkpoint.write_checkpoint) ------------------
# Sources of the layout: tensorflow/core/util/tensor_bundle (BundleHeaderProto / BundleEntryProto, masked crc32c),
# tensorflow/core/lib/io/table (LevelDB block format: prefix-compressed entries, restart array, 1-byte type + 4-byte crc
# trailer, 48-byte footer), tensorflow/core/protobuf/trackable_object_graph.proto.  The encoder below shares no code with the
# package: it is what a second implementer would write from those descriptions, with the field set a real
# `tf.train.Checkpoint(model).save()` of a Keras model produces (children edges, slot variables, save_counter, unknown fields).
def _vi(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _ld(field, payload):                      # length-delimited field
    return _vi(field << 3 | 2) + _vi(len(payload)) + payload


def _vf(field, value):                        # varint field
    return _vi(field << 3 | 0) + _vi(value)


def _crc32c_bitwise(data):
    crc = 0xFFFFFFFF
    for byte in data:
        crc ^= byte
        for _ in range(8):
            crc = (crc >> 1) ^ (0x82F63B78 & -(crc & 1))
    return crc ^ 0xFFFFFFFF


def _masked(crc):
    return ((crc >> 15 | crc << 17) + 0xA282EAD8) & 0xFFFFFFFF


def _leveldb_block(entries, restart_interval=16):
    out, restarts, last = bytearray(), [], b""
    for i, (key, value) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(key), len(last)) and key[shared] == last[shared]:
                shared += 1
        out += _vi(shared) + _vi(len(key) - shared) + _vi(len(value)) + key[shared:] + value
        last = key
    for r in restarts or [0]:
        out += struct.pack("<I", r)
    out += struct.pack("<I", max(len(restarts), 1))
    return bytes(out)


def test_reads_a_hand_assembled_tf2_bundle(tmp_path):
    rng = np.random.default_rng(7)
    kernel = rng.standard_normal((9, 9, 3, 32)).astype(np.float32)
    bias = rng.standard_normal(32).astype(np.float32)
    gamma = rng.uniform(0.5, 1.5, 32).astype(np.float32)
    rms = rng.uniform(0, 1, (9, 9, 3, 32)).astype(np.float32)
    tensors = [   # (checkpoint key, Keras full_name or None, dtype enum, array)  -- keys as tf.train.Checkpoint(model) writes them
        ("layer_with_weights-0/kernel/.ATTRIBUTES/VARIABLE_VALUE", "contract_start_conv/kernel", 1, kernel),
        ("layer_with_weights-0/bias/.ATTRIBUTES/VARIABLE_VALUE", "contract_start_conv/bias", 1, bias),
        ("layer_with_weights-1/gamma/.ATTRIBUTES/VARIABLE_VALUE", "batch_normalization/gamma", 1, gamma),
        ("layer_with_weights-0/kernel/.OPTIMIZER_SLOT/optimizer/rms/.ATTRIBUTES/VARIABLE_VALUE", "RMSprop/contract_start_conv/kernel/rms", 1, rms),
        ("save_counter/.ATTRIBUTES/VARIABLE_VALUE", "save_counter", 9, np.asarray(3, np.int64)),
    ]
    # ---- data shard + BundleEntryProto per tensor: dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6 (fixed32)
    data = bytearray()
    index_entries = {}
    for key, _full, dtype, arr in tensors:
        raw = arr.tobytes()
        shape = b"".join(_ld(2, _vf(1, int(d))) for d in arr.shape)            # TensorShapeProto.dim { size }
        entry = _vf(1, dtype) + _ld(2, shape) + (_vf(4, len(data)) if len(data) else b"") + _vf(5, len(raw)) + \
            _vi(6 << 3 | 5) + struct.pack("<I", _masked(_crc32c_bitwise(raw)))
        index_entries[key.encode()] = entry
        data += raw
    # ---- TrackableObjectGraph: node 0 = root with children; variables carry SerializedTensor{name, full_name, checkpoint_key}
    def variable_node(key, full):
        attr = _ld(1, b"VARIABLE_VALUE") + _ld(2, full.encode()) + _ld(3, key.encode())
        return _ld(2, attr) + _ld(5, _ld(1, b"") + _vf(2, 1))                   # + has_checkpoint_values wrapper (unknown to readers)
    def child(node_id, name):
        return _ld(1, _vf(1, node_id) + _ld(2, name.encode()))
    nodes = [
        child(1, "layer_with_weights-0") + child(2, "layer_with_weights-1") + child(6, "save_counter") + child(7, "optimizer"),   # root
        child(3, "kernel") + child(4, "bias"),                                                                 # conv layer
        child(5, "gamma"),                                                                                     # batch norm layer
        variable_node(tensors[0][0], tensors[0][1]) ,
        variable_node(tensors[1][0], tensors[1][1]),
        variable_node(tensors[2][0], tensors[2][1]),
        variable_node(tensors[4][0], tensors[4][1]),
        _ld(3, _vf(1, 3) + _ld(2, b"rms") + _vf(3, 8)),                         # optimizer: slot_variables {original 3, "rms", slot node 8}
        variable_node(tensors[3][0], tensors[3][1]),
    ]
    graph = b"".join(_ld(1, n) for n in nodes)
    # a scalar DT_STRING tensor is stored as [varint length][masked crc32c of the length as fixed 8 bytes... 4 bytes][bytes]
    graph_raw = _vi(len(graph)) + struct.pack("<I", _masked(_crc32c_bitwise(struct.pack("<Q", len(graph))))) + graph
    index_entries[b"_CHECKPOINTABLE_OBJECT_GRAPH"] = _vf(1, 7) + _ld(2, b"") + _vf(4, len(data)) + _vf(5, len(graph_raw)) + \
        _vi(6 << 3 | 5) + struct.pack("<I", 0)
    data += graph_raw
    index_entries[b""] = _vf(1, 1) + _ld(3, _vf(1, 1))                           # BundleHeaderProto{num_shards 1, version{producer 1}}
    # ---- the table: sorted keys in two data blocks, empty metaindex, index block, 48-byte footer
    keys = sorted(index_entries)
    table = bytearray()

    def emit(block):
        handle = _vi(len(table)) + _vi(len(block))
        table.extend(block + b"\x00" + struct.pack("<I", _masked(_crc32c_bitwise(block + b"\x00"))))
        return handle
    split = len(keys) // 2
    h0 = emit(_leveldb_block([(k, index_entries[k]) for k in keys[:split]], restart_interval=2))
    h1 = emit(_leveldb_block([(k, index_entries[k]) for k in keys[split:]], restart_interval=2))
    meta = emit(_leveldb_block([]))
    idx = emit(_leveldb_block([(keys[split - 1] + b"\x00", h0), (keys[-1] + b"~", h1)]))     # separators >= last key of the block
    footer = meta + idx
    table += footer + bytes(40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    prefix = tmp_path / "ckpt-3"
    (tmp_path / "ckpt-3.index").write_bytes(bytes(table))
    (tmp_path / "ckpt-3.data-00000-of-00001").write_bytes(bytes(data))
    (tmp_path / "checkpoint").write_text('model_checkpoint_path: "ckpt-3"\nall_model_checkpoint_paths: "ckpt-3"\n')

    got = ck.read_checkpoint_variables(str(prefix))
    assert got["save_counter/.ATTRIBUTES/VARIABLE_VALUE"]["value"].dtype == np.int64 and int(got["save_counter/.ATTRIBUTES/VARIABLE_VALUE"]["value"]) == 3
    np.testing.assert_array_equal(got[tensors[0][0]]["value"], kernel)
    assert got[tensors[0][0]]["full_name"] == "contract_start_conv/kernel"
    assert got[tensors[2][0]]["full_name"] == "batch_normalization/gamma"
    # ... and load_weights maps it onto a model, skipping the optimizer slot and the counter
    model, _ = styleTransfer.create_style_transfer_model((64, 128, 3), (64, 128, 3), 16, 32, 1)
    status = model.load_weights(str(tmp_path))                                   # directory -> follows the state file
    status.assert_nontrivial_match()
    np.testing.assert_array_equal(model.weights["contract_start/conv/kernel"], kernel)
    np.testing.assert_array_equal(model.weights["contract_start/conv/bias"], bias)
    np.testing.assert_array_equal(model.weights["contract_start/bn/gamma"], gamma)
    assert sorted(status.matched) == ["contract_start/bn/gamma", "contract_start/conv/bias", "contract_start/conv/kernel"]
    assert any("OPTIMIZER_SLOT" in k for k in status.unused) or all("rms" not in m for m in status.matched)


def test_native_crc32c_and_tensor_verification(tmp_path):
    """rst_host_crc32c (slicing-by-8 in the library, host code) against the RFC 3720 vectors and the Python table loop; every
    tensor of a written bundle carries its masked crc32c (also above 1 MiB) and read_checkpoint_variables(verify_tensors=True)
    checks it; the object graph has children edges whose names spell the checkpoint keys."""
    rng = np.random.default_rng(0)
    blob = rng.integers(0, 256, 100_003, dtype=np.uint8).tobytes()
    assert ck.crc32c(blob) == ck._crc32c_python(blob)
    assert ck.crc32c(b"123456789" * 1000) == ck._crc32c_python(b"123456789" * 1000)
    big = rng.standard_normal((3, 3, 256, 256)).astype(np.float32)              # 2.4 MB
    variables = {"residual_block_1/conv0/kernel": big, "residual_block_1/conv0/bias": rng.standard_normal(256).astype(np.float32)}
    prefix = ck.write_checkpoint(str(tmp_path / "ckpt-1"), variables)
    index = ck.read_index(prefix)
    key = "layer_with_weights-0/v/.ATTRIBUTES/VARIABLE_VALUE"
    assert index[key]["crc32c"] == ck._mask(ck._crc32c_python(big.tobytes())) != 0
    got = ck.read_checkpoint_variables(prefix, verify_tensors=True)
    np.testing.assert_array_equal(got[key]["value"], big)
    # corrupt one byte of the big tensor in the data shard
    data_path = prefix + ".data-00000-of-00001"
    raw = bytearray(open(data_path, "rb").read())
    raw[index[key]["offset"] + 12345] ^= 0x40
    open(data_path, "wb").write(raw)
    with pytest.raises(ValueError):
        ck.read_checkpoint_variables(prefix, verify_tensors=True)
    ck.read_checkpoint_variables(prefix)                                          # unverified read still returns the bytes
    # children edges: root -> layer_with_weights-i -> v
    e = index["_CHECKPOINTABLE_OBJECT_GRAPH"]
    shard = open(data_path, "rb").read()[e["offset"]:e["offset"] + e["size"]]
    n, pos = ck._get_varint(shard, 0)
    nodes = [v for fn, _wt, v in ck._parse_fields(shard[pos + 4:pos + 4 + n]) if fn == 1]
    root_children = [dict((f2, v2) for f2, _w, v2 in ck._parse_fields(v)) for fn, _wt, v in ck._parse_fields(nodes[0]) if fn == 1]
    assert [c[2] for c in root_children] == [b"layer_with_weights-0", b"layer_with_weights-1"]
    holder = nodes[root_children[0][1]]
    child = dict((f2, v2) for f2, _w, v2 in ck._parse_fields([v for fn, _wt, v in ck._parse_fields(holder) if fn == 1][0]))
    assert child[2] == b"v" and any(fn == 2 for fn, _wt, _v in ck._parse_fields(nodes[child[1]]))
