#!/usr/bin/env python3
"""Development aid: dump the op list of a TFLite flatbuffer (schema v3) without
any TensorFlow / flatbuffers dependency.  Used while writing oracle/ and the
product-side weight loader to read the three Lyra graphs
(lyra/model_coeffs/{soundstream_encoder,quantizer,lyragan}.tflite in the
reference tree).  Not imported by the product or the tests.

usage: tflite_dump.py model.tflite [--subgraph N] [--consts]
"""
import struct
import sys

import numpy as np

BUILTIN = {0: "ADD", 2: "CONCATENATION", 3: "CONV_2D", 4: "DEPTHWISE_CONV_2D", 6: "DEQUANTIZE",
           18: "MUL", 22: "RESHAPE", 36: "GATHER", 41: "SUB", 45: "STRIDED_SLICE", 49: "SPLIT",
           53: "CAST", 55: "MAXIMUM", 56: "ARG_MAX", 58: "LESS", 67: "TRANSPOSE_CONV",
           72: "NOT_EQUAL", 74: "SUM", 79: "ARG_MIN", 83: "PACK", 85: "ONE_HOT", 98: "LEAKY_RELU",
           99: "SQUARED_DIFFERENCE", 114: "QUANTIZE", 129: "CALL_ONCE", 142: "VAR_HANDLE",
           143: "READ_VARIABLE", 144: "ASSIGN_VARIABLE"}
TYPES = {0: "f32", 2: "i32", 3: "u8", 4: "i64", 6: "bool", 9: "i8", 13: "resource"}


class FB:
    def __init__(self, buf):
        self.b = buf

    def u8(self, o): return self.b[o]
    def i8(self, o): return struct.unpack_from("<b", self.b, o)[0]
    def u16(self, o): return struct.unpack_from("<H", self.b, o)[0]
    def i32(self, o): return struct.unpack_from("<i", self.b, o)[0]
    def u32(self, o): return struct.unpack_from("<I", self.b, o)[0]
    def f32(self, o): return struct.unpack_from("<f", self.b, o)[0]

    def root(self): return self.u32(0)

    def field(self, tab, fid):
        """absolute offset of field `fid` in table at `tab`, or 0 if absent."""
        vt = tab - self.i32(tab)
        vsz = self.u16(vt)
        slot = 4 + 2 * fid
        if slot >= vsz:
            return 0
        off = self.u16(vt + slot)
        return tab + off if off else 0

    def indirect(self, o): return o + self.u32(o)

    def table(self, tab, fid):
        o = self.field(tab, fid)
        return self.indirect(o) if o else 0

    def vec(self, tab, fid):
        """(start, length) of a vector field"""
        o = self.field(tab, fid)
        if not o:
            return 0, 0
        v = self.indirect(o)
        return v + 4, self.u32(v)

    def vec_tables(self, tab, fid):
        s, n = self.vec(tab, fid)
        return [self.indirect(s + 4 * i) for i in range(n)]

    def vec_i32(self, tab, fid):
        s, n = self.vec(tab, fid)
        return list(struct.unpack_from("<%di" % n, self.b, s)) if n else []

    def vec_f32(self, tab, fid):
        s, n = self.vec(tab, fid)
        return list(struct.unpack_from("<%df" % n, self.b, s)) if n else []

    def vec_i64(self, tab, fid):
        s, n = self.vec(tab, fid)
        return list(struct.unpack_from("<%dq" % n, self.b, s)) if n else []

    def string(self, tab, fid):
        s, n = self.vec(tab, fid)
        return bytes(self.b[s:s + n]).decode() if s else ""

    def scalar(self, tab, fid, fmt, default=0):
        o = self.field(tab, fid)
        return struct.unpack_from(fmt, self.b, o)[0] if o else default


def load(path):
    buf = open(path, "rb").read()
    fb = FB(buf)
    model = fb.root()
    opcodes = []
    for oc in fb.vec_tables(model, 1):
        dep = fb.scalar(oc, 0, "<b")
        new = fb.scalar(oc, 3, "<i")
        opcodes.append(max(dep, new))
    buffers = []
    for bt in fb.vec_tables(model, 4):
        s, n = fb.vec(bt, 0)
        buffers.append((s, n))
    subgraphs = []
    for sg in fb.vec_tables(model, 2):
        tensors = []
        for t in fb.vec_tables(sg, 0):
            q = fb.table(t, 4)
            tensors.append(dict(
                shape=fb.vec_i32(t, 0), type=TYPES.get(fb.scalar(t, 1, "<b"), "?"),
                buffer=fb.scalar(t, 2, "<I"), name=fb.string(t, 3),
                scale=fb.vec_f32(q, 2) if q else [], zp=fb.vec_i64(q, 3) if q else [],
                qdim=fb.scalar(q, 6, "<i") if q else 0,
                is_variable=fb.scalar(t, 5, "<b")))
        ops = []
        for op in fb.vec_tables(sg, 3):
            code = opcodes[fb.scalar(op, 0, "<I")]
            ops.append(dict(code=code, name=BUILTIN.get(code, str(code)),
                            inputs=fb.vec_i32(op, 1), outputs=fb.vec_i32(op, 2),
                            opt=fb.table(op, 4)))
        subgraphs.append(dict(tensors=tensors, ops=ops, inputs=fb.vec_i32(sg, 1),
                              outputs=fb.vec_i32(sg, 2), name=fb.string(sg, 4)))
    sigs = []
    for sd in fb.vec_tables(model, 7):
        sigs.append(dict(key=fb.string(sd, 2), subgraph=fb.scalar(sd, 4, "<I"),
                         inputs=[(fb.string(m, 0), fb.scalar(m, 1, "<I")) for m in fb.vec_tables(sd, 0)],
                         outputs=[(fb.string(m, 0), fb.scalar(m, 1, "<I")) for m in fb.vec_tables(sd, 1)]))
    return fb, subgraphs, buffers, sigs


def const_data(fb, buffers, t):
    s, n = buffers[t["buffer"]]
    if n == 0:
        return None
    dt = {"f32": np.float32, "i32": np.int32, "i8": np.int8, "i64": np.int64, "bool": np.bool_, "u8": np.uint8}[t["type"]]
    return np.frombuffer(fb.b, dtype=dt, count=n // np.dtype(dt).itemsize, offset=s).reshape(t["shape"] or [-1])


def opt_str(fb, op):
    o, n = op["opt"], op["name"]
    if not o:
        return ""
    g = lambda fid, fmt="<i", d=0: fb.scalar(o, fid, fmt, d)
    if n == "CONV_2D":
        return "pad=%d sw=%d sh=%d act=%d dw=%d dh=%d" % (g(0, "<b"), g(1), g(2), g(3, "<b"), g(4, "<i", 1), g(5, "<i", 1))
    if n == "DEPTHWISE_CONV_2D":
        return "pad=%d sw=%d sh=%d mult=%d act=%d dw=%d dh=%d" % (g(0, "<b"), g(1), g(2), g(3), g(4, "<b"), g(5, "<i", 1), g(6, "<i", 1))
    if n == "TRANSPOSE_CONV":
        return "pad=%d sw=%d sh=%d" % (g(0, "<b"), g(1), g(2))
    if n == "LEAKY_RELU":
        return "alpha=%g" % g(0, "<f")
    if n == "STRIDED_SLICE":
        return "bm=%d em=%d ell=%d na=%d sh=%d" % (g(0), g(1), g(2), g(3), g(4))
    if n == "CONCATENATION":
        return "axis=%d act=%d" % (g(0), g(1, "<b"))
    if n == "SPLIT":
        return "num=%d" % g(0)
    if n in ("ADD", "SUB", "MUL"):
        return "act=%d" % g(0, "<b")
    if n == "VAR_HANDLE":
        return "name=%s" % fb.string(o, 1)
    if n == "CALL_ONCE":
        return "init=%d" % g(0)
    if n == "GATHER":
        return "axis=%d batch_dims=%d" % (g(0), g(1))
    if n == "SUM":
        return "keep=%d" % g(0, "<b")
    if n == "PACK":
        return "n=%d axis=%d" % (g(0), g(1))
    if n == "ONE_HOT":
        return "axis=%d" % g(0)
    if n in ("ARG_MIN", "ARG_MAX"):
        return "out_type=%d" % g(0, "<b")
    if n == "CAST":
        return "in=%d out=%d" % (g(0, "<b"), g(1, "<b"))
    return ""


def main():
    path = sys.argv[1]
    only = int(sys.argv[sys.argv.index("--subgraph") + 1]) if "--subgraph" in sys.argv else None
    consts = "--consts" in sys.argv
    fb, sgs, buffers, sigs = load(path)
    for s in sigs:
        print("signature", s)
    for si, sg in enumerate(sgs):
        if only is not None and si != only:
            continue
        T = sg["tensors"]
        print("== subgraph %d '%s' inputs=%s outputs=%s  #tensors=%d #ops=%d" % (si, sg["name"], sg["inputs"], sg["outputs"], len(T), len(sg["ops"])))

        def td(i):
            if i < 0:
                return "-"
            t = T[i]
            q = ""
            if t["scale"]:
                q = " s=%s zp=%s" % (("%.6g" % t["scale"][0]) if len(t["scale"]) == 1 else "[%d]" % len(t["scale"]),
                                     t["zp"][0] if len(t["zp"]) == 1 else "[%d]" % len(t["zp"]))
            c = "C" if buffers[t["buffer"]][1] else ""
            return "#%d%s:%s%s%s" % (i, c, t["type"], t["shape"], q)
        for oi, op in enumerate(sg["ops"]):
            print("%3d %-18s %s -> %s  %s" % (oi, op["name"], " ".join(td(i) for i in op["inputs"]),
                                              " ".join(td(i) for i in op["outputs"]), opt_str(fb, op)))
        if consts:
            for i, t in enumerate(T):
                d = const_data(fb, buffers, t)
                if d is not None:
                    flat = d.reshape(-1)
                    print("const #%d %s %s %s : %s" % (i, t["name"], t["type"], t["shape"], flat[:8]))


if __name__ == "__main__":
    main()
