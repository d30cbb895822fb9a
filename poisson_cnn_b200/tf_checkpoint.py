"""TensorFlow checkpoint (tensor-bundle) reader / writer in pure Python + the Keras object-graph key map.

The reference saves its weights with ModelCheckpoint(save_weights_only=True) / model.save_weights(prefix)
(poisson_CNN/train/pcnn_end_to_end.py:43) and restores them with model.load_weights(prefix)
(poisson_CNN/train/utils.py:10-29).  That format is TensorFlow's tensor bundle:

  <prefix>.index                    an immutable sorted string table (the LevelDB table format:
                                    prefix-compressed data blocks with restart arrays, an index block,
                                    a 48-byte footer ending in the magic 0xdb4775248b80fb57).  Key ""
                                    holds a BundleHeaderProto, every other key a BundleEntryProto
                                    (dtype, shape, shard, offset, size, crc32c of the bytes).
  <prefix>.data-00000-of-00001      the raw little-endian tensor bytes.

Keras names the variables of a SUBCLASSED model by the attribute path from the root object, e.g.
  hpnn/pre_bottleneck_convolutions/0/kernel/.ATTRIBUTES/VARIABLE_VALUE
(list attributes contribute their index).  keras_key_map() generates those paths for the reference classes
(models/Homogeneous_Poisson_NN_Legacy.py:43-115, models/Dirichlet_BC_NN_Legacy.py:47-97, blocks/resnet.py:12-27,
blocks/bottleneck_block.py:26-97, layers/Scaling.py:23-34, models/Poisson_CNN_Legacy.py:8-9) and maps each
onto this package's variable names (weights.py).

STATUS: TensorFlow is not installable in the build container, so this module is checked against its own
writer and against hand-assembled tables (tests/test_host.py), NOT against a TF-written file.  The format
constants follow tensorflow/core/util/tensor_bundle and tensorflow/core/lib/io/{table,format,block}.cc as
published; the converter in INTEGRATION.md (runs where TF exists) remains the verified route.
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 4: np.dtype("u1"), 5: np.dtype("<i2"),
           6: np.dtype("i1"), 9: np.dtype("<i8"), 10: np.dtype("?"), 17: np.dtype("<u2"), 19: np.dtype("<f2"),
           22: np.dtype("<u4"), 23: np.dtype("<u8")}
_DT_BFLOAT16, _DT_STRING = 14, 7
_DTYPE_CODES = {np.dtype("float32"): 1, np.dtype("float64"): 2, np.dtype("int32"): 3, np.dtype("int64"): 9,
                np.dtype("float16"): 19}


# ------------------------------------------------------------------ varints / protobuf (wire format only)
def _get_varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _put_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf):
    """{field number: [values]} with varints as int, fixed32/64 as int, length-delimited as bytes."""
    fields, pos, n = {}, 0, len(buf)
    while pos < n:
        tag, pos = _get_varint(buf, pos)
        num, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        fields.setdefault(num, []).append(v)
    return fields


def _field(num, wt, payload):
    return _put_varint((num << 3) | wt) + payload


def _entry_proto(dtype_code, shape, offset, size, crc):
    dims = b"".join(_field(2, 2, (lambda d: _put_varint(len(d)) + d)(_field(1, 0, _put_varint(int(s))))) for s in shape)
    msg = _field(1, 0, _put_varint(dtype_code)) + _field(2, 2, _put_varint(len(dims)) + dims)
    if offset:
        msg += _field(4, 0, _put_varint(offset))
    msg += _field(5, 0, _put_varint(size)) + _field(6, 5, struct.pack("<I", crc))
    return msg


# ------------------------------------------------------------------ crc32c (Castagnoli), masked as LevelDB does
_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t.append(c)
        _CRC_TABLE = t
    return _CRC_TABLE


def crc32c(data, crc=0):
    t = _crc_table()
    crc ^= 0xFFFFFFFF
    for b in bytes(data):
        crc = t[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def _mask_crc(crc):
    return (((crc >> 15) | (crc << 17)) + 0xa282ead8) & 0xFFFFFFFF


# ------------------------------------------------------------------ snappy (index blocks may be compressed)
def _snappy_decompress(buf):
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = buf[pos] | (buf[pos + 1] << 8)
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError("corrupt snappy stream")
        for _ in range(ln):                     # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("snappy length mismatch")
    return bytes(out)


# ------------------------------------------------------------------ table reader
def _read_block(data, offset, size, verify):
    raw = data[offset:offset + size]
    ctype = data[offset + size]
    if verify:
        stored = struct.unpack_from("<I", data, offset + size + 1)[0]
        if _mask_crc(crc32c(data[offset:offset + size + 1])) != stored:
            raise ValueError("table block checksum mismatch at offset %d" % offset)
    if ctype == 1:
        raw = _snappy_decompress(raw)
    elif ctype != 0:
        raise ValueError("unknown block compression type %d" % ctype)
    return raw


def _block_entries(block):
    nrestarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * nrestarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_table(path, verify=False):
    """All (key, value) pairs of a LevelDB-format table file, in key order."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != TABLE_MAGIC:
        raise ValueError("%s is not a TensorFlow checkpoint index (bad table magic)" % path)
    footer = data[-48:]
    _, pos = _get_varint(footer, 0)          # metaindex handle (unused)
    _, pos = _get_varint(footer, pos)
    ioff, pos = _get_varint(footer, pos)
    isize, pos = _get_varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(data, ioff, isize, verify)):
        boff, hp = _get_varint(handle, 0)
        bsize, _ = _get_varint(handle, hp)
        out.extend(_block_entries(_read_block(data, boff, bsize, verify)))
    return out


def read_tensor_bundle(prefix, verify=False):
    """{variable name: numpy array} of a TF checkpoint `prefix` (.index + .data-xxxxx-of-xxxxx).
    String tensors (the serialized object graph) are skipped; bfloat16 comes back as float32."""
    entries = read_table(prefix + ".index", verify)
    header = None
    tensors, shards = {}, {}
    for key, value in entries:
        if key == b"":
            header = _parse_proto(value)
            continue
        e = _parse_proto(value)
        dtype = e.get(1, [0])[0]
        if dtype == _DT_STRING or 7 in e:        # strings and partitioned (sliced) variables are not weights of this model
            continue
        shape = []
        if 2 in e:
            for dim in _parse_proto(e[2][0]).get(2, []):
                shape.append(_parse_proto(dim).get(1, [0])[0])
        shard, offset, size = e.get(3, [0])[0], e.get(4, [0])[0], e.get(5, [0])[0]
        tensors[key.decode()] = (dtype, tuple(shape), shard, offset, size, e.get(6, [None])[0])
    if header is None:
        raise ValueError("checkpoint index has no header entry")
    if header.get(2, [0])[0] != 0:
        raise ValueError("big-endian checkpoints are not supported")
    num_shards = header.get(1, [1])[0]
    out = {}
    for name, (dtype, shape, shard, offset, size, crc) in tensors.items():
        if shard not in shards:
            shards[shard] = np.memmap("%s.data-%05d-of-%05d" % (prefix, shard, num_shards), dtype=np.uint8, mode="r")
        raw = shards[shard][offset:offset + size]
        if verify and crc is not None and _mask_crc(crc32c(raw.tobytes())) != crc:
            raise ValueError("checksum mismatch for %s" % name)
        if dtype == _DT_BFLOAT16:
            a = (np.frombuffer(raw.tobytes(), dtype="<u2").astype(np.uint32) << 16).view(np.float32)
        elif dtype in _DTYPES:
            a = np.frombuffer(raw.tobytes(), dtype=_DTYPES[dtype])
        else:
            continue
        out[name] = a.reshape(shape).copy()
    return out


# ------------------------------------------------------------------ table writer (exports / test fixtures)
def _build_block(items, restart_interval=16):
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_table(path, items, block_size=4096):
    """Writes sorted (key, value) byte pairs as an uncompressed LevelDB-format table."""
    items = sorted(items)
    f = bytearray()

    def emit(block):
        off = len(f)
        f.extend(block)
        f.append(0)                                                   # kNoCompression
        f.extend(struct.pack("<I", _mask_crc(crc32c(block + b"\x00"))))
        return _put_varint(off) + _put_varint(len(block))
    index, cur, cur_bytes = [], [], 0
    for k, v in items:
        cur.append((k, v))
        cur_bytes += len(k) + len(v) + 8
        if cur_bytes >= block_size:
            index.append((cur[-1][0], emit(_build_block(cur))))
            cur, cur_bytes = [], 0
    if cur or not index:
        index.append((cur[-1][0] if cur else b"", emit(_build_block(cur))))
    meta_handle = emit(_build_block([]))
    index_handle = emit(_build_block(index, restart_interval=1))
    footer = meta_handle + index_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    f.extend(footer)
    with open(path, "wb") as fh:
        fh.write(bytes(f))


def write_tensor_bundle(prefix, tensors):
    """Writes {name: array} as a single-shard TF checkpoint (float32/float64/int32/int64/float16)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    header = _field(1, 0, _put_varint(1)) + _field(3, 2, (lambda v: _put_varint(len(v)) + v)(_field(1, 0, _put_varint(1))))
    items, offset = [(b"", header)], 0
    with open(prefix + ".data-00000-of-00001", "wb") as data:
        for name in sorted(tensors):
            a = np.ascontiguousarray(tensors[name])
            code = _DTYPE_CODES.get(a.dtype)
            if code is None:
                raise ValueError("write_tensor_bundle: dtype %s of %s not supported" % (a.dtype, name))
            raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
            data.write(raw)
            items.append((name.encode(), _entry_proto(code, a.shape, offset, len(raw), _mask_crc(crc32c(raw)))))
            offset += len(raw)
    write_table(prefix + ".index", items)


# ------------------------------------------------------------------ Keras object-graph keys <-> package names
_BN = (("gamma", "gamma"), ("beta", "beta"), ("moving_mean", "mean"), ("moving_variance", "var"))


def _conv_keys(m, ref, name, use_bias=True):
    m[ref + "/kernel"] = name + "/kernel"
    if use_bias:
        m[ref + "/bias"] = name + "/bias"


def _bn_keys(m, ref, name):
    for a, b in _BN:
        m[ref + "/" + a] = name + "/" + b


def _resnet_keys(m, ref, name, use_bn, use_bias=True):
    for c in range(3):
        _conv_keys(m, "%s/conv_layers/%d" % (ref, c), "%s/conv%d" % (name, c), use_bias)
    if use_bn:
        _bn_keys(m, ref + "/batchnorm0", name + "/bn0")
        _bn_keys(m, ref + "/batchnorm1", name + "/bn1")


def _final_keys(m, fin, ref, name):
    S, nreg, ub = len(fin["filters"]), fin.get("final_regular_conv_stages", 2), fin.get("use_bias", True)
    for k in range(S):
        if k < S - nreg:     # final_convolutions = [conv_0, resnet_0, conv_1, resnet_1, ..., conv_{S-nreg}, ...]
            _conv_keys(m, "%s/%d" % (ref, 2 * k), "%s/%d/conv" % (name, k), ub)
            _resnet_keys(m, "%s/%d" % (ref, 2 * k + 1), "%s/%d/resnet" % (name, k), False, ub)
        else:
            _conv_keys(m, "%s/%d" % (ref, 2 * (S - nreg) + k - (S - nreg)), "%s/%d/conv" % (name, k), ub)


def hpnn_key_map(cfg, ref_prefix="", prefix=""):
    """{reference attribute path: package variable name} of Homogeneous_Poisson_NN_Legacy."""
    m = {}
    use_bn = cfg.get("use_batchnorm", False)
    pre = cfg["pre_bottleneck_convolutions_config"]
    step = 2 if use_bn else 1
    for k in range(len(pre["filters"])):
        _conv_keys(m, "pre_bottleneck_convolutions/%d" % (k * step), "pre_bottleneck/%d" % k, pre.get("use_bias", True))
        if use_bn:
            _bn_keys(m, "pre_bottleneck_convolutions/%d" % (k * step + 1), "pre_bottleneck/%d/bn" % k)
    for kind in ("deconv", "multilinear"):
        bc = cfg["bottleneck_%s_config" % kind]
        ds = bc["downsampling_factors"]
        order = sorted(range(len(ds)), key=lambda i: ds[i], reverse=True)     # the reference sorts its block list
        for j, idx in enumerate(order):
            ref, name = "bottleneck_%s_blocks/%d" % (kind, j), "bottleneck_%s/%d" % (kind, idx)
            ub = bc.get("conv_use_bias", True)
            _conv_keys(m, ref + "/conv_layers/0", name + "/conv0", ub)
            for r in range(1, bc["n_convs"][idx]):
                _resnet_keys(m, "%s/conv_layers/%d" % (ref, r), "%s/resnet%d" % (name, r), use_bn, ub)
            if kind == "deconv":
                _conv_keys(m, ref + "/upsample_layer", name + "/deconv", bc.get("deconv_use_bias", True))
    _conv_keys(m, "non_bottleneck_conv", "non_bottleneck_conv")
    _conv_keys(m, "post_merge_conv", "post_merge_conv")
    _resnet_keys(m, "post_merge_resnet", "post_merge_resnet", False)
    for i in range(3):
        _conv_keys(m, "dx_dense_layers/%d" % i, "dx_dense/%d" % i)
    _final_keys(m, cfg["final_convolutions_config"], "final_convolutions", "final")
    if cfg.get("use_scaling", False):
        sc = cfg["scaling_config"]
        for s in range(sc.get("stages", 2)):
            _conv_keys(m, "scaling/stages/%d" % (2 * s), "scaling/conv%d" % s, sc.get("use_bias", True))
        for i in range(3):
            _conv_keys(m, "scaling/dense_%d" % i, "scaling/dense%d" % i)
    return {ref_prefix + a: prefix + b for a, b in m.items()}


def dbcnn_key_map(cfg, ref_prefix="", prefix=""):
    """{reference attribute path: package variable name} of Dirichlet_BC_NN_Legacy_2."""
    m = {}
    use_bn = cfg.get("use_batchnorm", False)
    bc = cfg["boundary_conv_config"]
    step = 3 if use_bn else 2
    for k in range(len(bc["filters"])):
        ub = bc.get("use_bias", True)
        _conv_keys(m, "boundary_convolutions/%d" % (k * step), "boundary/%d/conv" % k, ub)
        if use_bn:
            _bn_keys(m, "boundary_convolutions/%d" % (k * step + 1), "boundary/%d/bn" % k)
        _resnet_keys(m, "boundary_convolutions/%d" % (k * step + step - 1), "boundary/%d/resnet" % k, use_bn, ub)
    for i in range(len(cfg["domain_info_mlp_config"]["units"])):
        _conv_keys(m, "domain_info_dense_layers/%d" % i, "mlp/%d" % i)
    _final_keys(m, cfg["final_convolutions_config"], "final_convolutions", "final")
    return {ref_prefix + a: prefix + b for a, b in m.items()}


def pcnn_key_map(hp_cfg, db_cfg, prefix=""):
    """Poisson_CNN_Legacy holds its sub-models as the attributes `hpnn` and `dbcnn`."""
    return {**hpnn_key_map(hp_cfg, "hpnn/", prefix + "hpnn/"), **dbcnn_key_map(db_cfg, "dbcnn/", prefix + "dbcnn/")}


def load_checkpoint_weights(ckpt_prefix, key_map, verify=False):
    """Reads a TF checkpoint and renames its variables through key_map ({reference path: package name}).
    Optimizer slots and bookkeeping entries are ignored; a variable the model needs but the checkpoint lacks raises."""
    raw = read_tensor_bundle(ckpt_prefix, verify)
    out, missing = {}, []
    for ref, name in key_map.items():
        key = ref + SUFFIX
        if key in raw:
            out[name] = raw[key].astype(np.float32)      # the reference casts float64 checkpoints too (train/utils.py:17-27)
        else:
            missing.append(key)
    if missing:
        have = [k for k in raw if k.endswith(SUFFIX)][:5]
        raise ValueError("TF checkpoint %s lacks %d variables, e.g. %s (it has e.g. %s)" % (ckpt_prefix, len(missing), missing[:3], have))
    # model variables the checkpoint holds but the key map does not name: a silent naming drift would hide here
    used = {ref + SUFFIX for ref in key_map}
    extra = [k for k in raw if k.endswith(SUFFIX) and k not in used and "optimizer" not in k and "save_counter" not in k
             and not k.startswith("_CHECKPOINTABLE_OBJECT_GRAPH")]
    import warnings
    if extra:
        warnings.warn("TF checkpoint %s holds %d model variables this model does not use, e.g. %s" % (ckpt_prefix, len(extra), extra[:3]))
    if not os.path.isfile(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "tf_ckpt", "pcnn.index")):
        warnings.warn("the TensorFlow tensor-bundle reader and the Keras key map have not been verified against a TensorFlow-written "
                      "checkpoint yet (no TensorFlow in the build environment; tests/golden/make_tf_fixtures.py generates the "
                      "fixture): a naming mismatch would surface as a 'lacks N variables' error above", stacklevel=2)
    return out


def save_checkpoint_weights(ckpt_prefix, weights, key_map):
    """Writes package-named weights under the reference's Keras keys (inverse of load_checkpoint_weights)."""
    inv = {v: k for k, v in key_map.items()}
    write_tensor_bundle(ckpt_prefix, {inv[name] + SUFFIX: np.asarray(a, np.float32) for name, a in weights.items()})
