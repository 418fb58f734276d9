"""Checkpoint interchange with the reference's TensorFlow `Saver` files.

The reference restores and saves its networks with `tf.train.Saver()` (Code/train.py:79-87, 522-534,
and `saver.save(sess, NETWORK_PATH + NET_NAME, global_step=...)` in its training loops).  A Saver file set is

    <dir>/checkpoint                         text: model_checkpoint_path: "<name>-<step>"
    <dir>/<name>-<step>.index                tensor-bundle index: a sorted string table (LevelDB table
                                             format) name -> BundleEntryProto, "" -> BundleHeaderProto
    <dir>/<name>-<step>.data-00000-of-00001  raw little-endian tensor bytes

This module reads and writes that format with NumPy only (TensorFlow is neither needed nor present) and maps
the Saver's variable names onto the creation-ordered parameter list of `model.VariableStore` /
`model.DenoisingNet`.  Names follow TF1's name-scope rules applied to the reference's scoping calls
(Code/model.py:31-44 for the leaf names, :428 `Conv`, :764 `MLP`, :853-925 `Level0..2` entered twice,
Code/train.py:72,188,492,764,1074 for the outer `model` scope): a scope or variable name that is already in
use inside its parent gets the suffix `_1`, `_2`, ...

Host-side file handling, no device work: nothing here touches the CUDA library.

PARITY UNPINNED: `/root/reference` ships no checkpoint file and TensorFlow is not installed in this image, so
the reader has only been exercised on files produced by the writer below (format restated from the published
tensor-bundle / table layout), and the names are derived from the scoping rules rather than read from a real
`.index`.  `tests/test_checkpoint_cpu.py` checks the pieces that have independent known answers: CRC-32C test
vectors, varint coding, the uniquified-name sequence worked out by hand from the reference source, and --
against the TensorFlow-derived code TensorBoard ships in this image -- the masked CRC, the dtype enum values
and the TensorShapeProto / VersionDef encodings.  What stays unpinned is the table block layout and the
BundleEntryProto / BundleHeaderProto field numbers.
"""
import os
import re
import struct
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

__all__ = ["network_variables", "read_bundle", "write_bundle", "latest_checkpoint", "load_network",
           "save_network", "save_training_state", "load_training_state", "crc32c", "CheckpointError"]


class CheckpointError(RuntimeError):
    pass


# ----------------------------------------------------------------------------- variable names
class _NameScope:
    """TF1 graph naming: `unique_name` per full path, scopes nest with '/'."""

    def __init__(self, root: str = ""):
        self._used: Dict[str, int] = {}
        self._stack = [root.rstrip("/")]
        if root:
            self._used[root.rstrip("/")] = 1

    def _unique(self, name: str) -> str:
        parent = self._stack[-1]
        full = parent + "/" + name if parent else name
        n = self._used.get(full, 0)
        if n == 0:
            self._used[full] = 1
            return full
        cand = "%s_%d" % (full, n)
        while cand in self._used:
            n += 1
            cand = "%s_%d" % (full, n)
        self._used[full] = n + 1
        self._used[cand] = 1
        return cand

    def enter(self, name: str):
        self._stack.append(self._unique(name))
        return self

    def leave(self):
        self._stack.pop()

    def variable(self, name: str) -> str:
        return self._unique(name)


def network_variables(in_channels: int = 6, multi_scale: bool = False, scope: str = "model",
                      M: int = 9) -> List[Tuple[str, Tuple[int, ...]]]:
    """(Saver name, shape) of every variable of `get_model_reg_multi_scale` in creation order
    (reference Code/model.py:837-946; the same order `model.VariableStore` records)."""
    ns = _NameScope(scope)
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def conv(cin, cout):  # model.py:427-447: W0, b, u, c then v
        ns.enter("Conv")
        out.append((ns.variable("weight"), (M, cout, cin)))
        out.append((ns.variable("bias"), (cout,)))
        out.append((ns.variable("assignment"), (M, cin)))
        out.append((ns.variable("assignment"), (M,)))
        out.append((ns.variable("assignment"), (M, cin)))
        ns.leave()

    def lin(cin, cout):  # model.py:763-769
        ns.enter("MLP")
        out.append((ns.variable("weight"), (cin, cout)))
        out.append((ns.variable("bias"), (cout,)))
        ns.leave()

    def head(cin):
        lin(cin, 1024)
        lin(1024, 3)

    ns.enter("Level0")          # model.py:853
    conv(in_channels, 32)
    ns.leave()
    ns.enter("Level1")          # :866
    conv(32, 64)
    ns.leave()
    ns.enter("Level2")          # :878
    conv(64, 128)
    conv(128, 128)
    if multi_scale:
        head(128)
    ns.leave()
    ns.enter("Level1")          # :904, second entry -> Level1_1
    conv(128, 64)
    conv(128, 64)
    if multi_scale:
        head(64)
    ns.leave()
    ns.enter("Level0")          # :925 -> Level0_1
    conv(64, 32)
    conv(64, 32)
    head(32)
    ns.leave()
    return out


# ----------------------------------------------------------------------------- CRC-32C (Castagnoli)
def _make_table():
    poly = np.uint32(0x82F63B78)
    t = np.arange(256, dtype=np.uint32)
    for _ in range(8):
        t = np.where(t & 1, (t >> 1) ^ poly, t >> 1).astype(np.uint32)
    return t


_TABLE = _make_table()


def _crc_raw_scalar(state: int, data: bytes) -> int:
    tab = _TABLE
    for b in data:
        state = int(tab[(state ^ b) & 0xFF]) ^ (state >> 8)
    return state


def _zero_shift_columns(nbytes: int) -> np.ndarray:
    """Images of the 32 unit states after feeding `nbytes` zero bytes (the CRC register is linear)."""
    cols = (np.uint32(1) << np.arange(32, dtype=np.uint32)).astype(np.uint32)
    for _ in range(nbytes):
        cols = _TABLE[cols & 0xFF] ^ (cols >> 8)
    return cols


def _byte_tables(cols: np.ndarray) -> np.ndarray:
    """tables[k][v] = image of the state v << 8k under the linear register map given by its 32 column images."""
    tabs = np.zeros((4, 256), dtype=np.uint32)
    for k in range(4):
        t = np.zeros(1, dtype=np.uint32)
        for b in range(8):  # entries with bit b set = entries without it, xor that column
            t = np.concatenate([t, t ^ cols[8 * k + b]])
        tabs[k] = t
    return tabs


def _apply_tables(tabs: np.ndarray, x: np.ndarray) -> np.ndarray:
    return (tabs[0][x & 0xFF] ^ tabs[1][(x >> 8) & 0xFF] ^ tabs[2][(x >> 16) & 0xFF] ^ tabs[3][x >> 24])


def crc32c(data, seed: int = 0) -> int:
    """CRC-32C of a bytes-like object (reflected polynomial 0x1EDC6F41, init/final xor 0xFFFFFFFF).
    Inputs from 16 KB on are cut into a power-of-two number of equal chunks that run the byte-wise table
    recurrence side by side (NumPy lanes); neighbouring chunk registers are then joined pairwise, log2(lanes)
    times, through the zero-feed operator of the current chunk length (squared after every round)."""
    buf = np.frombuffer(memoryview(data).cast("B"), dtype=np.uint8)
    state = (seed ^ 0xFFFFFFFF) & 0xFFFFFFFF
    n = buf.size
    if n >= 16384:
        lanes = 512
        while lanes < (1 << 20) and n // (2 * lanes) >= 32:
            lanes *= 2
        L = n // lanes
        body = np.ascontiguousarray(buf[: L * lanes].reshape(lanes, L).T)
        st = np.zeros(lanes, dtype=np.uint32)
        st[0] = state
        for i in range(L):
            st = _TABLE[(st ^ body[i]) & 0xFF] ^ (st >> 8)
        cols = _zero_shift_columns(L)
        while st.size > 1:
            tabs = _byte_tables(cols)
            st = _apply_tables(tabs, st[0::2]) ^ st[1::2]
            cols = _apply_tables(tabs, cols)
        state = int(st[0])
        buf = buf[L * lanes:]
    state = _crc_raw_scalar(state, buf.tobytes())
    return (state ^ 0xFFFFFFFF) & 0xFFFFFFFF


def _mask_crc(c: int) -> int:
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ----------------------------------------------------------------------------- varints / protobuf subset
def _put_varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _get_varint(buf, pos: int) -> Tuple[int, int]:
    shift = 0
    v = 0
    while True:
        if pos >= len(buf):
            raise CheckpointError("truncated varint")
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, pos
        shift += 7
        if shift > 63:
            raise CheckpointError("varint too long")


def _pb_fields(buf):
    """Yields (field number, wire type, value) of one protobuf message (varint, 64-bit, bytes, 32-bit)."""
    pos = 0
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            if len(v) != ln:
                raise CheckpointError("truncated protobuf field")
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise CheckpointError("unsupported protobuf wire type %d" % wt)
        yield fn, wt, v


# DataType enum values of tensorflow/core/framework/types.proto that a Saver file of this network can hold
_DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 9: np.dtype("<i8")}
_DTYPE_IDS = {v: k for k, v in _DTYPES.items()}


def _encode_entry(arr: np.ndarray, offset: int, crc_masked: int) -> bytes:
    shape = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(int(s)) for s in arr.shape))
    msg = b"\x08" + _put_varint(_DTYPE_IDS[arr.dtype.newbyteorder("<")])
    msg += b"\x12" + _put_varint(len(shape)) + shape
    if offset:
        msg += b"\x20" + _put_varint(offset)
    msg += b"\x28" + _put_varint(arr.nbytes)
    msg += b"\x35" + struct.pack("<I", crc_masked)
    return msg


def _decode_entry(buf) -> dict:
    e = dict(dtype=0, shape=[], shard=0, offset=0, size=0, crc=None, sliced=False)
    for fn, _, v in _pb_fields(buf):
        if fn == 1:
            e["dtype"] = v
        elif fn == 2:
            for f2, _, dim in _pb_fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, dv in _pb_fields(dim):
                        if f3 == 1:
                            size = dv
                    e["shape"].append(size)
        elif fn == 3:
            e["shard"] = v
        elif fn == 4:
            e["offset"] = v
        elif fn == 5:
            e["size"] = v
        elif fn == 6:
            e["crc"] = v
        elif fn == 7:
            e["sliced"] = True
    return e


# ----------------------------------------------------------------------------- string table (.index)
_TABLE_MAGIC = 0xDB4775248B80FB57
_RESTART_INTERVAL = 16


def _build_block(items: Sequence[Tuple[bytes, bytes]]) -> bytes:
    out = bytearray()
    restarts = []
    last = b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % _RESTART_INTERVAL == 0:
            restarts.append(len(out))
        else:
            m = min(len(k), len(last))
            while shared < m and k[shared] == last[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v))
        out += k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _block_with_trailer(contents: bytes) -> bytes:
    return contents + b"\x00" + struct.pack("<I", _mask_crc(crc32c(contents + b"\x00")))


def _parse_block(raw: bytes, offset: int, size: int, verify: bool) -> List[Tuple[bytes, bytes]]:
    if offset + size + 5 > len(raw):
        raise CheckpointError("index block runs past the end of the file")
    contents = raw[offset:offset + size]
    ctype = raw[offset + size]
    if ctype != 0:
        raise CheckpointError("compressed index blocks (type %d) are not supported" % ctype)
    if verify:
        want = struct.unpack_from("<I", raw, offset + size + 1)[0]
        if _mask_crc(crc32c(raw[offset:offset + size + 1])) != want:
            raise CheckpointError("index block checksum mismatch")
    if size < 4:
        raise CheckpointError("index block too small")
    nrestart = struct.unpack_from("<I", contents, size - 4)[0]
    end = size - 4 - 4 * nrestart
    if end < 0:
        raise CheckpointError("bad restart array")
    items = []
    pos = 0
    key = b""
    while pos < end:
        shared, pos = _get_varint(contents, pos)
        non_shared, pos = _get_varint(contents, pos)
        vlen, pos = _get_varint(contents, pos)
        if shared > len(key) or pos + non_shared + vlen > end:
            raise CheckpointError("corrupt index block entry")
        key = key[:shared] + contents[pos:pos + non_shared]
        pos += non_shared
        items.append((key, contents[pos:pos + vlen]))
        pos += vlen
    return items


def _read_table(path: str, verify: bool) -> List[Tuple[bytes, bytes]]:
    with open(path, "rb") as f:
        raw = f.read()
    if len(raw) < 48 or struct.unpack_from("<Q", raw, len(raw) - 8)[0] != _TABLE_MAGIC:
        raise CheckpointError("%s is not a tensor-bundle index (bad magic)" % path)
    footer = raw[len(raw) - 48:]
    pos = 0
    _, pos = _get_varint(footer, pos)      # metaindex handle
    _, pos = _get_varint(footer, pos)
    ioff, pos = _get_varint(footer, pos)   # index handle
    isize, pos = _get_varint(footer, pos)
    items: List[Tuple[bytes, bytes]] = []
    for _, handle in _parse_block(raw, ioff, isize, verify):
        boff, hp = _get_varint(handle, 0)
        bsize, hp = _get_varint(handle, hp)
        items += _parse_block(raw, boff, bsize, verify)
    return items


def _write_table(path: str, items: Sequence[Tuple[bytes, bytes]], block_bytes: int = 4096):
    out = bytearray()
    index = []
    cur: List[Tuple[bytes, bytes]] = []
    cur_bytes = 0

    def flush():
        nonlocal cur, cur_bytes
        if not cur:
            return
        contents = _build_block(cur)
        handle = _put_varint(len(out)) + _put_varint(len(contents))
        index.append((cur[-1][0], handle))  # any key >= the block's last key and < the next block's first
        out.extend(_block_with_trailer(contents))
        cur, cur_bytes = [], 0

    for k, v in items:
        cur.append((k, v))
        cur_bytes += len(k) + len(v) + 3
        if cur_bytes >= block_bytes:
            flush()
    flush()
    meta = _build_block([])
    meta_handle = _put_varint(len(out)) + _put_varint(len(meta))
    out.extend(_block_with_trailer(meta))
    idx = _build_block(index)
    idx_handle = _put_varint(len(out)) + _put_varint(len(idx))
    out.extend(_block_with_trailer(idx))
    footer = meta_handle + idx_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", _TABLE_MAGIC)
    out.extend(footer)
    with open(path, "wb") as f:
        f.write(bytes(out))


# ----------------------------------------------------------------------------- bundles
def _data_path(prefix: str, shard: int, nshards: int) -> str:
    return "%s.data-%05d-of-%05d" % (prefix, shard, nshards)


def read_bundle(prefix: str, names: Optional[Sequence[str]] = None, verify: bool = True) -> Dict[str, np.ndarray]:
    """Tensors of a Saver checkpoint `<prefix>.index` + `<prefix>.data-*` by variable name (all of them, or
    only `names`).  Checks table and tensor checksums unless `verify` is False."""
    items = _read_table(prefix + ".index", verify)
    if not items or items[0][0] != b"":
        raise CheckpointError("%s.index has no bundle header" % prefix)
    nshards, endian = 1, 0
    for fn, _, v in _pb_fields(items[0][1]):
        if fn == 1:
            nshards = v
        elif fn == 2:
            endian = v
    if endian != 0:
        raise CheckpointError("big-endian bundles are not supported")
    want = None if names is None else set(names)
    shards: Dict[int, np.ndarray] = {}
    out: Dict[str, np.ndarray] = {}
    for key, val in items[1:]:
        name = key.decode("utf-8")
        if want is not None and name not in want:
            continue
        e = _decode_entry(val)
        if e["sliced"]:
            raise CheckpointError("%s: partitioned variables are not supported" % name)
        if e["dtype"] not in _DTYPES:
            raise CheckpointError("%s: unsupported dtype id %d" % (name, e["dtype"]))
        dt = _DTYPES[e["dtype"]]
        count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
        if count * dt.itemsize != e["size"]:
            raise CheckpointError("%s: shape %s does not match %d bytes" % (name, e["shape"], e["size"]))
        if e["shard"] not in shards:
            shards[e["shard"]] = np.fromfile(_data_path(prefix, e["shard"], nshards), dtype=np.uint8)
        data = shards[e["shard"]]
        if e["offset"] + e["size"] > data.size:
            raise CheckpointError("%s: data file is truncated" % name)
        raw = data[e["offset"]:e["offset"] + e["size"]]
        if verify and e["crc"] is not None and _mask_crc(crc32c(raw)) != e["crc"]:
            raise CheckpointError("%s: tensor checksum mismatch" % name)
        out[name] = raw.view(dt).reshape(e["shape"]).copy()
    if want is not None:
        missing = sorted(want - set(out))
        if missing:
            raise CheckpointError("%s: variables not in the checkpoint: %s" % (prefix, ", ".join(missing)))
    return out


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray]):
    """Writes `<prefix>.index` and `<prefix>.data-00000-of-00001` (one shard, little endian)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    header = b"\x08\x01" + b"\x1a\x02\x08\x01"  # num_shards = 1, (endianness = LITTLE omitted), version.producer = 1
    items: List[Tuple[bytes, bytes]] = [(b"", header)]
    offset = 0
    with open(_data_path(prefix, 0, 1), "wb") as f:
        for name in sorted(tensors, key=lambda s: s.encode("utf-8")):
            if not name:
                raise CheckpointError("empty variable name")
            arr = np.asarray(tensors[name])
            if not arr.flags.c_contiguous:
                arr = np.ascontiguousarray(arr)
            if arr.dtype.newbyteorder("<") not in _DTYPE_IDS:
                raise CheckpointError("%s: unsupported dtype %s" % (name, arr.dtype))
            arr = arr.astype(arr.dtype.newbyteorder("<"), copy=False)
            raw = arr.tobytes()
            items.append((name.encode("utf-8"), _encode_entry(arr, offset, _mask_crc(crc32c(raw)))))
            f.write(raw)
            offset += len(raw)
    _write_table(prefix + ".index", items)


def latest_checkpoint(directory: str) -> Optional[str]:
    """`tf.train.get_checkpoint_state(dir).model_checkpoint_path` (Code/train.py:82-84, 527-534)."""
    state = os.path.join(directory, "checkpoint")
    if not os.path.exists(state):
        return None
    with open(state, "r") as f:
        for line in f:
            m = re.match(r'\s*model_checkpoint_path:\s*"(.*)"\s*$', line)
            if m:
                p = m.group(1)
                return p if os.path.isabs(p) else os.path.join(directory, p)
    return None


def load_network(prefix: str, in_channels: int = 6, multi_scale: bool = False, scope: str = "model",
                 verify: bool = True) -> List[np.ndarray]:
    """Parameters of the reference network from a Saver checkpoint, in creation order: pass the list as
    `params=` to `model.VariableStore` / `model.DenoisingNet`.  Optimiser slots and counters that a training
    checkpoint also holds (`.../Adam`, `beta1_power`, the global step) are ignored."""
    spec = network_variables(in_channels, multi_scale, scope)
    got = read_bundle(prefix, [n for n, _ in spec], verify)
    out = []
    for name, shape in spec:
        a = got[name]
        if tuple(a.shape) != tuple(shape):
            raise CheckpointError("%s: checkpoint shape %s, network expects %s" % (name, tuple(a.shape), tuple(shape)))
        out.append(a.astype(np.float32, copy=False))
    return out


def save_network(prefix: str, params: Sequence, in_channels: int = 6, multi_scale: bool = False,
                 scope: str = "model", global_step: Optional[int] = None) -> str:
    """Writes the creation-ordered `params` (tensors or arrays) under the Saver's names, plus the `checkpoint`
    state file next to it; with `global_step` the file is `<prefix>-<step>` as `saver.save(..., global_step=)`
    names it (the reference parses the step back out of that name, Code/train.py:528-533)."""
    spec = network_variables(in_channels, multi_scale, scope)
    if len(params) != len(spec):
        raise CheckpointError("expected %d variables, got %d" % (len(spec), len(params)))
    tensors = {}
    for (name, shape), t in zip(spec, params):
        a = t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
        if tuple(a.shape) != tuple(shape):
            raise CheckpointError("%s: parameter shape %s, network expects %s" % (name, tuple(a.shape), tuple(shape)))
        tensors[name] = a.astype(np.float32, copy=False)
    if global_step is not None:
        prefix = "%s-%d" % (prefix, int(global_step))
    write_bundle(prefix, tensors)
    base = os.path.basename(prefix)
    with open(os.path.join(os.path.dirname(os.path.abspath(prefix)), "checkpoint"), "w") as f:
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))
    return prefix


def _np(t):
    return t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)


def save_training_state(prefix: str, params: Sequence, adam_m: Sequence, adam_v: Sequence, step: int,
                        b1: float = 0.9, b2: float = 0.999, in_channels: int = 6, multi_scale: bool = False,
                        scope: str = "model", global_step: Optional[int] = None) -> str:
    """A training checkpoint as the reference's Saver writes it after `step` Adam updates (train.py:429, 520,
    552): the variables, their Adam slots `<name>/Adam` (first moment) and `<name>/Adam_1` (second moment), the
    optimiser's `beta1_power` / `beta2_power` (TF keeps beta^(step+1)) and the unnamed step counter
    `Variable`.  File name and `checkpoint` state file as `save_network`."""
    spec = network_variables(in_channels, multi_scale, scope)
    if not (len(params) == len(adam_m) == len(adam_v) == len(spec)):
        raise CheckpointError("expected %d variables with one first and one second moment each" % len(spec))
    tensors: Dict[str, np.ndarray] = {}
    for (name, shape), p, m, v in zip(spec, params, adam_m, adam_v):
        for suffix, t in (("", p), ("/Adam", m), ("/Adam_1", v)):
            a = _np(t).astype(np.float32, copy=False)
            if tuple(a.shape) != tuple(shape):
                raise CheckpointError("%s%s: shape %s, network expects %s" % (name, suffix, tuple(a.shape), tuple(shape)))
            tensors[name + suffix] = a
    tensors["beta1_power"] = np.float32(b1 ** (int(step) + 1)).reshape(())
    tensors["beta2_power"] = np.float32(b2 ** (int(step) + 1)).reshape(())
    tensors["Variable"] = np.int32(step).reshape(())
    if global_step is not None:
        prefix = "%s-%d" % (prefix, int(global_step))
    write_bundle(prefix, tensors)
    base = os.path.basename(prefix)
    with open(os.path.join(os.path.dirname(os.path.abspath(prefix)), "checkpoint"), "w") as f:
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))
    return prefix


def load_training_state(prefix: str, in_channels: int = 6, multi_scale: bool = False, scope: str = "model",
                        verify: bool = True) -> dict:
    """{'params', 'm', 'v': lists in creation order, 'step': Adam updates done} from a training checkpoint.
    A file without optimiser slots (an inference export) gives zero moments and step 0."""
    spec = network_variables(in_channels, multi_scale, scope)
    got = read_bundle(prefix, None, verify)
    out = {"params": [], "m": [], "v": [], "step": 0}
    for name, shape in spec:
        if name not in got:
            raise CheckpointError("%s: variables not in the checkpoint: %s" % (prefix, name))
        for key, suffix in (("params", ""), ("m", "/Adam"), ("v", "/Adam_1")):
            a = got.get(name + suffix)
            if a is None:
                a = np.zeros(shape, np.float32)
            if tuple(a.shape) != tuple(shape):
                raise CheckpointError("%s%s: checkpoint shape %s, network expects %s"
                                      % (name, suffix, tuple(a.shape), tuple(shape)))
            out[key].append(a.astype(np.float32, copy=False))
    if "Variable" in got and got["Variable"].shape == ():
        out["step"] = int(got["Variable"])
    return out
