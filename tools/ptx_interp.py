#!/usr/bin/env python3
"""Single-thread interpreter for the PTX subset tools/gen_eval_check.py emits (test infrastructure, CPU only).

The generated eval_check kernels are straight-line 32-bit integer code plus two forward branches; running one thread
of them in exact integer arithmetic lets the CPU test suite check the EMITTED TEXT (register naming, carry chains,
address arithmetic, constant offsets) against the reference's compiled poly_fp, not just the scalar DAG it was
emitted from. Anything outside the subset raises, so a generator change that starts emitting a new instruction fails
loudly here instead of being skipped.

    k = Kernel(open("eval_check_rv32im_p0.ptx").read())
    k.run(params={"p_check": 0x1000_0000_0000, ..., "p_cst": bytes}, mem=Memory({...}), tid=5, ctaid=0, ntid=128)
"""
import re
import struct

M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF


class Memory:
    """global memory as {base address: numpy uint32 array}; addresses are byte addresses"""

    def __init__(self, regions):
        self.regions = sorted(regions.items())

    def _find(self, addr):
        for base, arr in self.regions:
            off = addr - base
            if 0 <= off < 4 * len(arr):
                assert off % 4 == 0, "unaligned access"
                return arr, off // 4
        raise IndexError("global access outside every buffer: 0x%x" % addr)

    def load(self, addr):
        arr, i = self._find(addr)
        return int(arr[i])

    def store(self, addr, v):
        arr, i = self._find(addr)
        arr[i] = v


_TOK = re.compile(r"\s*(@!?%\w+\s+)?([\w.]+)\s*(.*?);\s*$")


def _split_operands(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "{[":
            depth += 1
        elif ch in "}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


class Kernel:
    def __init__(self, text):
        self.params = []     # (name, kind) in declaration order; kind = "u32" | "u64" | ("b8", size)
        self.code = []       # (pred, negate, op, operands)
        self.labels = {}
        in_body = False
        header = ""
        for raw in text.split("\n"):
            line = raw.split("//")[0].strip()
            if not line:
                continue
            if not in_body:
                header += " " + line
                if line == "{":
                    in_body = True
                    for m in re.finditer(r"\.param\s+(?:\.align\s+\d+\s+)?\.(\w+)\s+(\w+)(?:\[(\d+)\])?", header):
                        ty, name, size = m.group(1), m.group(2), m.group(3)
                        self.params.append((name, ("b8", int(size)) if size else ty))
                continue
            if line == "}":
                break
            if line.startswith(".reg"):
                continue
            if line.endswith(":"):
                self.labels[line[:-1]] = len(self.code)
                continue
            m = _TOK.match(line)
            if not m:
                raise ValueError("cannot parse: " + raw)
            pred = m.group(1)
            neg = False
            if pred:
                pred = pred.strip()[1:]
                if pred.startswith("!"):
                    neg, pred = True, pred[1:]
            self.code.append((pred, neg, m.group(2), _split_operands(m.group(3))))

    def run(self, params, mem, tid=0, ctaid=0, ntid=128, max_steps=10_000_000):
        R = {"%tid.x": tid, "%ctaid.x": ctaid, "%ntid.x": ntid}
        carry = 0

        def val(x):
            if x[0] == "%":
                return R[x]
            return int(x, 0)

        def param_load(expr, width):
            m = re.match(r"\[(\w+)(?:\+(\d+))?\]", expr)
            name, off = m.group(1), int(m.group(2) or 0)
            p = params[name]
            if isinstance(p, (bytes, bytearray)):
                return struct.unpack_from("<I" if width == 32 else "<Q", p, off)[0]
            assert off == 0
            return p & (M32 if width == 32 else M64)

        def addr_of(expr):
            m = re.match(r"\[(%\w+)(?:\+(\d+))?\]", expr)
            return (R[m.group(1)] + int(m.group(2) or 0)) & M64

        pc, steps = 0, 0
        while pc < len(self.code):
            steps += 1
            if steps > max_steps:
                raise RuntimeError("step limit")
            pred, neg, op, a = self.code[pc]
            pc += 1
            if pred is not None and bool(R[pred]) == neg:
                continue
            if op == "ld.param.u32":
                R[a[0]] = param_load(a[1], 32)
            elif op == "ld.param.u64":
                R[a[0]] = param_load(a[1], 64)
            elif op == "cvta.to.global.u64":
                R[a[0]] = val(a[1])
            elif op in ("mov.u32", "mov.u64"):
                R[a[0]] = val(a[1]) & (M32 if op.endswith("32") else M64)
            elif op == "mov.b64":
                if a[0].startswith("{"):            # unpack: {lo, hi} = src
                    lo, hi = [x.strip() for x in a[0][1:-1].split(",")]
                    v = val(a[1])
                    R[lo], R[hi] = v & M32, (v >> 32) & M32
                else:
                    R[a[0]] = val(a[1]) & M64
            elif op == "add.u32":
                R[a[0]] = (val(a[1]) + val(a[2])) & M32
            elif op == "sub.u32":
                R[a[0]] = (val(a[1]) - val(a[2])) & M32
            elif op == "neg.s32":
                R[a[0]] = (-val(a[1])) & M32
            elif op == "min.u32":
                R[a[0]] = min(val(a[1]) & M32, val(a[2]) & M32)
            elif op == "and.b32":
                R[a[0]] = val(a[1]) & val(a[2]) & M32
            elif op == "shl.b32":
                R[a[0]] = (val(a[1]) << val(a[2])) & M32
            elif op == "mul.lo.u32":
                R[a[0]] = ((val(a[1]) & M32) * (val(a[2]) & M32)) & M32
            elif op == "mul.hi.u32":
                R[a[0]] = (((val(a[1]) & M32) * (val(a[2]) & M32)) >> 32) & M32
            elif op == "mul.wide.u32":
                R[a[0]] = (val(a[1]) & M32) * (val(a[2]) & M32)
            elif op == "mad.lo.u32":
                R[a[0]] = ((val(a[1]) & M32) * (val(a[2]) & M32) + val(a[3])) & M32
            elif op == "mad.wide.u32":
                R[a[0]] = ((val(a[1]) & M32) * (val(a[2]) & M32) + val(a[3])) & M64
            elif op == "mad.lo.cc.u32":
                t = (((val(a[1]) & M32) * (val(a[2]) & M32)) & M32) + (val(a[3]) & M32)
                R[a[0]], carry = t & M32, t >> 32
            elif op == "madc.hi.u32":
                t = ((((val(a[1]) & M32) * (val(a[2]) & M32)) >> 32) & M32) + (val(a[3]) & M32) + carry
                R[a[0]] = t & M32
            elif op == "add.u64":
                R[a[0]] = (val(a[1]) + val(a[2])) & M64
            elif op == "cvt.u64.u32":
                R[a[0]] = val(a[1]) & M32
            elif op in ("setp.eq.u32", "setp.ne.u32", "setp.ge.u32"):
                x, y = val(a[1]) & M32, val(a[2]) & M32
                R[a[0]] = {"eq": x == y, "ne": x != y, "ge": x >= y}[op.split(".")[1]]
            elif op == "selp.u32":
                R[a[0]] = val(a[1]) if R[a[3]] else val(a[2])
            elif op in ("ld.global.nc.u32", "ld.global.u32"):
                R[a[0]] = mem.load(addr_of(a[1]))
            elif op == "st.global.u32":
                mem.store(addr_of(a[0]), val(a[1]) & M32)
            elif op in ("prefetch.global.L1", "prefetch.global.L2"):
                mem.load(addr_of(a[0]))        # must at least be a valid address
            elif op == "bra":
                pc = self.labels[a[0]]
            elif op == "bar.sync":
                pass
            elif op == "ret":
                return steps
            else:
                raise NotImplementedError("PTX op outside the generator's subset: %s" % op)
        return steps
