"""Oracle-side reader for the two on-disk formats at the hot path's boundary:
TensorFlow SavedModel (``saved_model.pb``) and the tensor-bundle V2 checkpoint
(``variables/variables.index`` + ``.data-00000-of-00001``).

TEST INFRASTRUCTURE ONLY (see oracle/shifu_oracle.py header).  Independent of the
product's C++ reader/writer (shifu-tensorflow_b200/csrc/savedmodel.cpp) so that each
checks the other.  Written from the published format descriptions (protobuf wire
format; leveldb table format: block = entries + restart array, 5-byte trailer
{type, masked crc32c}, 48-byte footer with magic 0xdb4775248b80fb57; TF
tensor_bundle.proto BundleHeaderProto/BundleEntryProto) because TF itself is not
vendored in the reference (shifu-tensorflow-eval/pom.xml:45,59-73).

What the reference does with these files:
  writer  res/ssgd_monitor.py:457-490 (simple_save + export_generic_config)
  reader  shifu-tensorflow-eval/src/main/java/ml/shifu/shifu/tensorflow/TensorflowModel.java:169
          (SavedModelBundle.load) and :71,85 (feed / fetch by raw op name)
"""
from __future__ import annotations

import os
import struct
from typing import Dict, List, Tuple

import numpy as np

# ----------------------------- protobuf wire walker -----------------------------

def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    r, s = 0, 0
    while True:
        b = buf[pos]; pos += 1
        r |= (b & 0x7F) << s
        if not (b & 0x80):
            return r, pos
        s += 7


def parse_proto(buf: bytes) -> List[Tuple[int, int, object]]:
    """-> [(field_no, wire_type, value)]; value is int for varint/fixed, bytes for len-delimited."""
    out, pos, n = [], 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]; pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = bytes(buf[pos:pos + ln]); pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]; pos += 4
        else:
            raise ValueError("unsupported wire type %d" % wt)
        out.append((fn, wt, v))
    return out


def _fields(msg, fn):
    return [v for f, _, v in msg if f == fn]


# ----------------------------- crc32c (Castagnoli) -----------------------------
_CRC_TABLE = None


def crc32c(data: bytes, crc: int = 0) -> int:
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t.append(c)
        _CRC_TABLE = np.array(t, dtype=np.uint32)
    c = crc ^ 0xFFFFFFFF
    tab = _CRC_TABLE
    for b in data:
        c = int(tab[(c ^ b) & 0xFF]) ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def crc_mask(crc: int) -> int:
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ----------------------------- leveldb table (tensor-bundle index) -----------------------------
TABLE_MAGIC = 0xDB4775248B80FB57


def _read_block(buf: bytes, off: int, size: int, verify: bool = True) -> List[Tuple[bytes, bytes]]:
    data = buf[off:off + size]
    ctype = buf[off + size]
    if ctype != 0:
        raise ValueError("compressed table blocks unsupported (type %d)" % ctype)
    if verify:
        stored = struct.unpack_from("<I", buf, off + size + 1)[0]
        if crc_mask(crc32c(buf[off:off + size + 1])) != stored:
            raise ValueError("table block crc mismatch")
    n_restarts = struct.unpack_from("<I", data, len(data) - 4)[0]
    end = len(data) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _varint(data, pos)
        non_shared, pos = _varint(data, pos)
        vlen, pos = _varint(data, pos)
        key = key[:shared] + data[pos:pos + non_shared]; pos += non_shared
        out.append((key, data[pos:pos + vlen])); pos += vlen
    return out


def read_table(path: str, verify: bool = True) -> List[Tuple[bytes, bytes]]:
    buf = open(path, "rb").read()
    footer = buf[-48:]
    if struct.unpack_from("<Q", footer, 40)[0] != TABLE_MAGIC:
        raise ValueError("bad table magic")
    pos = 0
    _, pos = _varint(footer, pos); _, pos = _varint(footer, pos)      # metaindex handle
    ioff, pos = _varint(footer, pos); isz, pos = _varint(footer, pos)  # index handle
    entries = []
    for _, handle in _read_block(buf, ioff, isz, verify):
        boff, p = _varint(handle, 0); bsz, p = _varint(handle, p)
        entries.extend(_read_block(buf, boff, bsz, verify))
    return entries


DT_FLOAT, DT_INT32, DT_STRING, DT_INT64, DT_BOOL = 1, 3, 7, 9, 10


def read_bundle(prefix: str, verify_crc: bool = False) -> Dict[str, np.ndarray]:
    """prefix = ".../variables/variables".  Returns {name: ndarray} for DT_FLOAT/INT32/INT64 entries."""
    entries = read_table(prefix + ".index")
    out, data_cache = {}, {}
    for key, val in entries:
        if key == b"":
            continue  # BundleHeaderProto {num_shards=1, endianness=2, version=3}
        m = parse_proto(val)
        dtype = (_fields(m, 1) or [0])[0]
        shape = []
        for sh in _fields(m, 2):
            for d in _fields(parse_proto(sh), 2):
                shape.append((_fields(parse_proto(d), 1) or [0])[0])
        shard = (_fields(m, 3) or [0])[0]
        off = (_fields(m, 4) or [0])[0]
        size = (_fields(m, 5) or [0])[0]
        np_dt = {DT_FLOAT: np.float32, DT_INT32: np.int32, DT_INT64: np.int64}.get(dtype)
        if np_dt is None:
            continue
        if shard not in data_cache:
            data_cache[shard] = np.memmap("%s.data-%05d-of-%05d" % (prefix, shard, 1), dtype=np.uint8, mode="r")
        raw = bytes(data_cache[shard][off:off + size])
        if verify_crc:
            stored = (_fields(m, 6) or [0])[0]
            if crc_mask(crc32c(raw)) != stored:
                raise ValueError("tensor crc mismatch for %r" % key)
        out[key.decode()] = np.frombuffer(raw, dtype=np_dt).reshape(shape).copy()
    return out


# ----------------------------- SavedModel graph walk -----------------------------

def read_graph_nodes(saved_model_pb: str, tag: str = "serve"):
    """-> ({node_name: (op, [inputs], {attr: raw AttrValue bytes})}, signature dict)."""
    sm = parse_proto(open(saved_model_pb, "rb").read())
    for mg_raw in _fields(sm, 2):
        mg = parse_proto(mg_raw)
        tags = []
        for mi in _fields(mg, 1):
            tags += [t.decode() for t in _fields(parse_proto(mi), 4)]
        if tag not in tags:
            continue
        nodes = {}
        for gd in _fields(mg, 2):
            for nd_raw in _fields(parse_proto(gd), 1):
                nd = parse_proto(nd_raw)
                name = _fields(nd, 1)[0].decode()
                op = _fields(nd, 2)[0].decode()
                inputs = [i.decode() for i in _fields(nd, 3)]
                attrs = {}
                for a in _fields(nd, 5):
                    am = parse_proto(a)
                    attrs[_fields(am, 1)[0].decode()] = _fields(am, 2)[0]
                nodes[name] = (op, inputs, attrs)
        sigs = {}
        for s in _fields(mg, 5):
            sm_ = parse_proto(s)
            sigs[_fields(sm_, 1)[0].decode()] = _fields(sm_, 2)[0]
        return nodes, sigs
    raise ValueError("no meta graph with tag %r" % tag)


_ACT_OPS = {"Sigmoid": 0, "Tanh": 1, "Relu": 2, "LeakyRelu": 3}


def _strip(name: str) -> str:
    name = name.lstrip("^")
    return name.split(":")[0]


def extract_mlp(model_dir: str, input_name: str, output_name: str, tag: str = "serve"):
    """Walk output -> input through (act?) <- BiasAdd|Add <- MatMul chains, skipping Identity and
    inference-disabled Keras dropout (cond/Switch/Merge) nodes.  Returns
    (layers=[(W[in,out], b[out], act_id or -1)], variable-name list)."""
    nodes, _ = read_graph_nodes(os.path.join(model_dir, "saved_model.pb"), tag)
    bundle = read_bundle(os.path.join(model_dir, "variables", "variables"))

    def resolve_var(n):
        n = _strip(n)
        for _ in range(8):
            op, ins, _a = nodes[n]
            if op in ("VariableV2", "Variable", "VarHandleOp"):
                return n
            if op in ("Identity", "ReadVariableOp"):
                n = _strip(ins[0]); continue
            raise ValueError("cannot resolve variable from %s (%s)" % (n, op))
        raise ValueError("variable chain too long")

    def skip_passthrough(n):
        """follow Identity / dropout-cond Merge(Switch) plumbing down to the producing tensor"""
        n = _strip(n)
        while True:
            op, ins, _a = nodes[n]
            if op in ("Identity", "StopGradient"):
                n = _strip(ins[0])
            elif op == "Merge":
                # keras dropout in_train_phase: Merge(cond/Switch_1 (inference branch), cond/dropout/mul)
                nxt = None
                for i in ins:
                    si = _strip(i)
                    if nodes[si][0] == "Switch":
                        nxt = _strip(nodes[si][1][0]); break
                if nxt is None:
                    raise ValueError("unsupported Merge at %s" % n)
                n = nxt
            elif op == "Switch":
                n = _strip(ins[0])
            else:
                return n

    layers, names = [], []
    cur = skip_passthrough(output_name)
    target = _strip(input_name)
    while cur != target:
        op, ins, attrs = nodes[cur]
        act = -1
        if op in _ACT_OPS:
            act = _ACT_OPS[op]
            cur = skip_passthrough(ins[0]); op, ins, attrs = nodes[cur]
        if op not in ("BiasAdd", "Add", "AddV2"):
            raise ValueError("expected BiasAdd/Add at %s, got %s" % (cur, op))
        a, b = skip_passthrough(ins[0]), ins[1]
        if nodes[a][0] != "MatMul":
            a, b = skip_passthrough(ins[1]), ins[0]
        bias_var = resolve_var(b)
        mop, mins, _m = nodes[a]
        if mop != "MatMul":
            raise ValueError("expected MatMul at %s, got %s" % (a, mop))
        w_var = resolve_var(mins[1])
        layers.append((bundle[w_var], bundle[bias_var], act))
        names.append((w_var, bias_var))
        cur = skip_passthrough(mins[0])
    layers.reverse(); names.reverse()
    return layers, names


def mlp_forward(layers, X: np.ndarray) -> np.ndarray:
    from . import shifu_oracle as so
    A = X.astype(np.float32)
    for W, b, act in layers:
        z = A @ W + b
        A = so.act_forward(z, act) if act >= 0 else z
    return A
