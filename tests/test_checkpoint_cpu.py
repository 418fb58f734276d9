"""Saver-checkpoint interchange: variable names, tensor-bundle reader/writer.  CPU only."""
import os
import struct

import numpy as np
import pytest

from facet_graph_convolution_b200 import checkpoint as ck


def test_crc32c_known_answers_and_lane_path():
    # published check values of CRC-32C (iSCSI): "123456789", 32 zero bytes, 32 0xFF bytes
    assert ck.crc32c(b"123456789") == 0xE3069283
    assert ck.crc32c(bytes(32)) == 0x8A9136AA
    assert ck.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert ck.crc32c(b"") == 0
    rs = np.random.RandomState(0)
    for n in (8191, 16383, 16384, 16385, 32768 + 17, 70001, 1 << 20):
        data = rs.randint(0, 256, n).astype(np.uint8).tobytes()
        want = ck._crc_raw_scalar(0xFFFFFFFF, data) ^ 0xFFFFFFFF  # byte-at-a-time recurrence
        assert ck.crc32c(data) == want, n
    # the mask used by the table and bundle formats: rotate right by 15, add a constant
    assert ck._mask_crc(0) == 0xA282EAD8
    assert ck._mask_crc(0x00008000) == (1 + 0xA282EAD8) & 0xFFFFFFFF


def test_varints_round_trip():
    for v in (0, 1, 127, 128, 300, 2 ** 32 - 1, 2 ** 63 - 1):
        enc = ck._put_varint(v)
        assert ck._get_varint(enc, 0) == (v, len(enc))
    assert ck._put_varint(300) == b"\xac\x02"
    with pytest.raises(ck.CheckpointError):
        ck._get_varint(b"\x80", 0)


def test_variable_names_follow_the_reference_scoping():
    """Worked out by hand from Code/model.py:31-44 (leaf names), :428 / :764 (Conv / MLP scopes), :853-925
    (Level0, Level1, Level2, then Level1 and Level0 entered a second time) and train.py:72 (outer scope)."""
    names = [n for n, _ in ck.network_variables()]
    conv = ["weight", "bias", "assignment", "assignment_1", "assignment_2"]
    mlp = ["weight", "bias"]
    want = []
    for scope, leaves in (("Level0/Conv", conv), ("Level1/Conv", conv), ("Level2/Conv", conv),
                          ("Level2/Conv_1", conv), ("Level1_1/Conv", conv), ("Level1_1/Conv_1", conv),
                          ("Level0_1/Conv", conv), ("Level0_1/Conv_1", conv), ("Level0_1/MLP", mlp),
                          ("Level0_1/MLP_1", mlp)):
        want += ["model/%s/%s" % (scope, leaf) for leaf in leaves]
    assert names == want
    ms = ck.network_variables(multi_scale=True)
    ms_names = [n for n, _ in ms]
    assert len(ms) == len(want) + 8 and len(set(ms_names)) == len(ms_names)
    i = ms_names.index("model/Level2/MLP/weight")
    assert ms_names[i - 1] == "model/Level2/Conv_1/assignment_2" and ms_names[i + 4] == "model/Level1_1/Conv/weight"
    assert dict(ms)["model/Level1_1/MLP/weight"] == (64, 1024) and dict(ms)["model/Level2/MLP_1/weight"] == (1024, 3)
    shapes = dict(ck.network_variables())
    assert shapes["model/Level0/Conv/weight"] == (9, 32, 6)
    assert shapes["model/Level1_1/Conv/weight"] == (9, 64, 128) and shapes["model/Level1_1/Conv_1/assignment"] == (9, 128)
    assert shapes["model/Level0_1/Conv_1/assignment_1"] == (9,)
    assert shapes["model/Level0_1/MLP/weight"] == (32, 1024) and shapes["model/Level0_1/MLP_1/bias"] == (3,)


def test_name_scope_uniquifier_matches_tf_rules():
    ns = ck._NameScope("model")
    assert ns.variable("w") == "model/w" and ns.variable("w") == "model/w_1" and ns.variable("w") == "model/w_2"
    assert ns.variable("w_1") == "model/w_1_1"  # the explicit name is taken by the uniquified one above
    ns.enter("A")
    assert ns.variable("w") == "model/A/w"
    ns.leave()
    ns.enter("A")
    assert ns.variable("w") == "model/A_1/w"


def _params(multi_scale, seed=0):
    rs = np.random.RandomState(seed)
    return [rs.randn(*s).astype(np.float32) for _, s in ck.network_variables(multi_scale=multi_scale)]


@pytest.mark.parametrize("multi_scale", [False, True])
def test_network_round_trip_with_training_extras(tmp_path, multi_scale):
    params = _params(multi_scale)
    prefix = ck.save_network(str(tmp_path / "net"), params, multi_scale=multi_scale, global_step=1234)
    assert os.path.basename(prefix) == "net-1234"
    assert ck.latest_checkpoint(str(tmp_path)) == prefix
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    back = ck.load_network(prefix, multi_scale=multi_scale)
    assert len(back) == len(params) and all(np.array_equal(a, b) for a, b in zip(back, params))
    # a training checkpoint also carries optimiser slots and counters: they are ignored by name
    names = [n for n, _ in ck.network_variables(multi_scale=multi_scale)]
    tensors = dict(zip(names, params))
    for n, p in zip(names, params):
        tensors[n + "/Adam"] = np.zeros_like(p)
        tensors[n + "/Adam_1"] = np.ones_like(p)
    tensors["beta1_power"] = np.float32(0.9).reshape(())
    tensors["Variable"] = np.int32(1234).reshape(())
    tensors["big"] = np.arange(7, dtype=np.int64)
    ck.write_bundle(str(tmp_path / "full"), tensors)
    back = ck.load_network(str(tmp_path / "full"), multi_scale=multi_scale)
    assert all(np.array_equal(a, b) for a, b in zip(back, params))
    everything = ck.read_bundle(str(tmp_path / "full"))
    assert set(everything) == set(tensors)
    assert everything["Variable"].shape == () and int(everything["Variable"]) == 1234
    assert everything["big"].dtype == np.int64 and everything["beta1_power"].dtype == np.float32


def test_index_layout_is_the_published_table_format(tmp_path):
    prefix = str(tmp_path / "t")
    ck.write_bundle(prefix, {"b/x": np.arange(6, dtype=np.float32).reshape(2, 3), "a": np.float32([1.5])})
    raw = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xDB4775248B80FB57 and len(raw) > 48
    items = ck._read_table(prefix + ".index", True)
    assert [k for k, _ in items] == [b"", b"a", b"b/x"]  # header first, then names in byte order
    e = ck._decode_entry(items[2][1])
    assert e["dtype"] == 1 and e["shape"] == [2, 3] and e["offset"] == 4 and e["size"] == 24
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    assert data == np.float32([1.5]).tobytes() + np.arange(6, dtype=np.float32).tobytes()
    assert e["crc"] == ck._mask_crc(ck.crc32c(data[4:]))


def test_many_variables_span_several_index_blocks(tmp_path):
    tensors = {"scope_%03d/some/long/variable/name" % i: np.full((3,), i, np.float32) for i in range(400)}
    ck.write_bundle(str(tmp_path / "m"), tensors)
    raw = open(str(tmp_path / "m") + ".index", "rb").read()
    footer = raw[-48:]
    pos = 0
    for _ in range(2):
        _, pos = ck._get_varint(footer, pos)
    ioff, pos = ck._get_varint(footer, pos)
    isize, pos = ck._get_varint(footer, pos)
    assert len(ck._parse_block(raw, ioff, isize, True)) > 1
    back = ck.read_bundle(str(tmp_path / "m"))
    assert all(np.array_equal(back[k], v) for k, v in tensors.items())


def test_corruption_and_mismatches_are_reported(tmp_path):
    params = _params(False)
    prefix = ck.save_network(str(tmp_path / "net"), params)
    dpath = prefix + ".data-00000-of-00001"
    raw = bytearray(open(dpath, "rb").read())
    raw[100] ^= 0x40
    open(dpath, "wb").write(bytes(raw))
    with pytest.raises(ck.CheckpointError, match="checksum"):
        ck.load_network(prefix)
    assert len(ck.load_network(prefix, verify=False)) == len(params)
    open(dpath, "wb").write(bytes(raw[:-8]))
    with pytest.raises(ck.CheckpointError, match="truncated"):
        ck.load_network(prefix, verify=False)
    # single-scale file into the multi-scale network: the extra heads are missing
    prefix2 = ck.save_network(str(tmp_path / "b" / "net"), params)
    with pytest.raises(ck.CheckpointError, match="not in the checkpoint"):
        ck.load_network(prefix2, multi_scale=True)
    with pytest.raises(ck.CheckpointError, match="expects"):
        ck.load_network(prefix2, in_channels=3)
    ipath = prefix2 + ".index"
    idx = bytearray(open(ipath, "rb").read())
    idx[10] ^= 1
    open(ipath, "wb").write(bytes(idx))
    with pytest.raises(ck.CheckpointError):
        ck.read_bundle(prefix2)
    open(ipath, "wb").write(b"not a table")
    with pytest.raises(ck.CheckpointError, match="magic"):
        ck.read_bundle(prefix2)
    assert ck.latest_checkpoint(str(tmp_path / "nowhere")) is None
    with pytest.raises(ck.CheckpointError, match="expected"):
        ck.save_network(str(tmp_path / "c" / "net"), params[:-1])


def test_crc_and_protobuf_pieces_against_tensorboard_implementations():
    """TensorBoard ships an independent masked CRC-32C (TFRecord framing uses the same mask as the table and
    bundle formats) and TensorFlow's generated protobuf classes for shapes, dtypes and versions: the pieces of
    the bundle encoding they cover must parse / agree."""
    tb = pytest.importorskip("tensorboard.compat.tensorflow_stub.pywrap_tensorflow")
    shape_pb2 = pytest.importorskip("tensorboard.compat.proto.tensor_shape_pb2")
    types_pb2 = pytest.importorskip("tensorboard.compat.proto.types_pb2")
    versions_pb2 = pytest.importorskip("tensorboard.compat.proto.versions_pb2")
    rs = np.random.RandomState(1)
    for n in (0, 1, 9, 4096, 20000):
        data = rs.randint(0, 256, n).astype(np.uint8).tobytes()
        assert ck.crc32c(data) == tb.crc32c(data)
        assert ck._mask_crc(ck.crc32c(data)) == tb.masked_crc32c(data)
    assert {types_pb2.DT_FLOAT: "<f4", types_pb2.DT_DOUBLE: "<f8", types_pb2.DT_INT32: "<i4",
            types_pb2.DT_INT64: "<i8"} == {k: v.str for k, v in ck._DTYPES.items()}
    arr = np.zeros((9, 64, 128), np.float32)
    fields = {fn: v for fn, _, v in ck._pb_fields(ck._encode_entry(arr, 1234, 0xDEADBEEF))}
    shape = shape_pb2.TensorShapeProto()
    shape.ParseFromString(fields[2])
    assert [d.size for d in shape.dim] == [9, 64, 128] and not shape.unknown_rank
    assert fields[1] == types_pb2.DT_FLOAT and fields[4] == 1234 and fields[5] == arr.nbytes and fields[6] == 0xDEADBEEF
    # and the other way: a shape serialised by the real class is decoded by the reader
    real = shape_pb2.TensorShapeProto(dim=[shape_pb2.TensorShapeProto.Dim(size=s) for s in (3, 1, 70000)])
    entry = b"\x08\x01\x12" + ck._put_varint(len(real.SerializeToString())) + real.SerializeToString() + b"\x28\x04"
    assert ck._decode_entry(entry)["shape"] == [3, 1, 70000]
    scalar = ck._decode_entry(ck._encode_entry(np.zeros((), np.int64), 0, 1))
    assert scalar["shape"] == [] and scalar["dtype"] == types_pb2.DT_INT64 and scalar["size"] == 8
    ver = versions_pb2.VersionDef()
    ver.ParseFromString(b"\x08\x01")  # the header's version submessage written by write_bundle
    assert ver.producer == 1


def test_training_state_resumes_adam_exactly(tmp_path):
    """Saver-style training checkpoint (variables, `/Adam`, `/Adam_1`, beta powers, step counter): five Adam
    updates, save, restore into fresh objects, five more == ten uninterrupted updates, bit for bit."""
    import torch
    from facet_graph_convolution_b200 import train as T
    spec = ck.network_variables()

    def fresh():
        g = torch.Generator().manual_seed(0)
        ps = [(torch.randn(s, generator=g) * 0.05).requires_grad_(True) for _, s in spec]
        b = T.GradBucket(ps)
        return ps, b, T.Adam(b)

    def update(ps, b, opt, k):
        g = torch.Generator().manual_seed(100 + k)
        for p in ps:
            p.grad = torch.randn(p.shape, generator=g)
        b.pack()
        opt.step()

    ps_a, b_a, opt_a = fresh()
    for k in range(10):
        update(ps_a, b_a, opt_a, k)
    ps_b, b_b, opt_b = fresh()
    for k in range(5):
        update(ps_b, b_b, opt_b, k)
    m, v = opt_b.state_lists()
    prefix = ck.save_training_state(str(tmp_path / "net"), ps_b, m, v, opt_b.t, global_step=5)
    assert ck.latest_checkpoint(str(tmp_path)) == prefix and prefix.endswith("net-5")
    everything = ck.read_bundle(prefix)
    assert len(everything) == 3 * len(spec) + 3
    assert np.float32(everything["beta1_power"]) == np.float32(0.9 ** 6) and int(everything["Variable"]) == 5
    assert "model/Level1_1/Conv_1/assignment_2/Adam_1" in everything
    st = ck.load_training_state(prefix)
    assert st["step"] == 5
    ps_c = [torch.from_numpy(a.copy()).requires_grad_(True) for a in st["params"]]
    b_c = T.GradBucket(ps_c)
    opt_c = T.Adam(b_c)
    opt_c.load_state(st["m"], st["v"], st["step"])
    for k in range(5, 10):
        update(ps_c, b_c, opt_c, k)
    for a, c in zip(ps_a, ps_c):
        assert torch.equal(a.detach(), c.detach())
    # the inference loader reads the same file and ignores the optimiser's entries
    back = ck.load_network(prefix)
    assert all(np.array_equal(x, p.detach().numpy()) for x, p in zip(back, ps_b))
    # an inference export resumes with zero moments
    p2 = ck.save_network(str(tmp_path / "inf" / "net"), ps_b)
    st2 = ck.load_training_state(p2)
    assert st2["step"] == 0 and all(not a.any() for a in st2["m"]) and all(not a.any() for a in st2["v"])
    with pytest.raises(ck.CheckpointError, match="moment"):
        ck.save_training_state(str(tmp_path / "x"), ps_b, m[:-1], v, 1)
