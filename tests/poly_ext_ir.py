"""TEST INFRASTRUCTURE: `CircuitCoreDef::poly_ext` (risc0/zkp/src/adapter.rs, called by verify/mod.rs:370-390) for the
rv32im and recursion circuits, evaluated from the committed circuit IR (risc0_b200/circuits/<name>.ir.json.gz - the
constraint polynomial as a DAG, parsed from the reference's generated poly_fp by tools/circuit_ir.py).

The verifier's validity check is  check(z) * ((3z)^N - 1) == poly_ext(poly_mix, eval_u, [out, mix])  where eval_u are
the tap evaluations at z (extension field). rv32im's own poly_ext.rs is a missing large blob in the reference snapshot
(.MISSING_LARGE_BLOBS:7); the polynomial is the same one poly_fp evaluates over the base field, so it is evaluated here
over FpExt instead. Plain Python big-int arithmetic (normal form), ~0.3 s per evaluation."""
import ctypes as C
import gzip
import json
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 15 * 2**27 + 1
R = (1 << 32) % P
RINV = pow(R, P - 2, P)
NB = P - 11


def _table(header, name):
    text = open(os.path.join(ROOT, "risc0_b200", "csrc", "tables", header)).read()
    m = re.search(r"%s\[\d*\]\s*=\s*\{([^}]*)\}" % name, text)
    return [int(x, 0) for x in re.findall(r"0x[0-9a-fA-F]+|\d+", m.group(1))]


class CircuitPolyExt:
    def __init__(self, name):
        up = name.upper()
        d = json.load(gzip.open(os.path.join(ROOT, "risc0_b200", "circuits", name + ".ir.json.gz")))
        self.nodes, self.types, self.root = d["nodes"], d["types"], d["root"]
        flat = _table("circuit_%s.h" % name, up + "_TAPS")
        # taps: (offset, back, group, combo, skip) per tap, in the verifier's eval_u order
        group_names = {0: "accum", 1: "code", 2: "data"}
        self.tap_index = {}
        for i in range(len(flat) // 5):
            off, back, group = flat[5 * i], flat[5 * i + 1], flat[5 * i + 2]
            self.tap_index[(group_names[group], off, back)] = i
        self.pows = _table("circuit_%s.h" % name, up + "_POLY_MIX_POWERS")
        # reachable nodes in topological (index) order: operands always precede their users in the DAG
        live = set()
        stack = [self.root]
        while stack:
            n = stack.pop()
            if n in live:
                continue
            live.add(n)
            k = self.nodes[n]
            if k[0] in "+-*":
                stack += [k[1], k[2]]
        self.order = sorted(live)

    def __call__(self, poly_mix, eval_u, out, mix):
        """all arguments in normal form: poly_mix 4-tuple, eval_u list of 4-tuples, out / mix lists of ints"""
        def emul(a, b):
            return ((a[0] * b[0] + NB * (a[1] * b[3] + a[2] * b[2] + a[3] * b[1])) % P,
                    (a[0] * b[1] + a[1] * b[0] + NB * (a[2] * b[3] + a[3] * b[2])) % P,
                    (a[0] * b[2] + a[1] * b[1] + a[2] * b[0] + NB * (a[3] * b[3])) % P,
                    (a[0] * b[3] + a[1] * b[2] + a[2] * b[1] + a[3] * b[0]) % P)

        def epow(a, n):
            r = (1, 0, 0, 0)
            while n:
                if n & 1:
                    r = emul(r, a)
                a = emul(a, a)
                n >>= 1
            return r

        pm_cache = {}
        val = {}
        for n in self.order:
            k = self.nodes[n]
            op = k[0]
            if op == "c":
                v = (k[1] % P, 0, 0, 0)
            elif op == "ce":
                v = tuple(x % P for x in k[1:5])
            elif op == "t":
                name = {"ctrl": "code"}.get(k[1], k[1])
                v = eval_u[self.tap_index[(name, k[2], k[3])]]
            elif op == "g":
                src = mix if k[1] == "mix" else out
                v = (src[k[2]], 0, 0, 0)
            elif op == "pm":
                if k[1] not in pm_cache:
                    pm_cache[k[1]] = epow(poly_mix, self.pows[k[1]])
                v = pm_cache[k[1]]
            elif op == "+":
                a, b = val[k[1]], val[k[2]]
                v = tuple((x + y) % P for x, y in zip(a, b))
            elif op == "-":
                a, b = val[k[1]], val[k[2]]
                v = tuple((x - y) % P for x, y in zip(a, b))
            else:
                v = emul(val[k[1]], val[k[2]])
            val[n] = v
        return val[self.root]

    def callback(self):
        """ctypes callback for orc_verify_*_ext (Montgomery words in and out)"""
        u32p = C.POINTER(C.c_uint32)
        proto = C.CFUNCTYPE(None, u32p, u32p, C.c_uint64, u32p, C.c_uint64, u32p, C.c_uint64, u32p)

        def dec(ptr, n):
            return [int(x) * RINV % P for x in np.ctypeslib.as_array(ptr, shape=(n,))]

        def fn(poly_mix, eval_u, ntaps, out, nout, mix, nmix, result):
            pm = tuple(dec(poly_mix, 4))
            eu = dec(eval_u, 4 * ntaps)
            res = self(pm, [tuple(eu[4 * i:4 * i + 4]) for i in range(ntaps)], dec(out, nout), dec(mix, nmix))
            for i in range(4):
                result[i] = res[i] * R % P

        self._cb = proto(fn)   # keep alive
        return self._cb


_CACHE = {}


def poly_ext(name):
    if name not in _CACHE:
        _CACHE[name] = CircuitPolyExt(name)
    return _CACHE[name]
