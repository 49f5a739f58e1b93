#!/usr/bin/env python3
"""Front end of the eval_check code generator: turns the reference's zirgen-generated constraint polynomial
(`poly_fp`, straight-line C++ split over 21 functions with per-thread scratch arrays) into one flat, hash-consed
expression DAG.

Input (read-only, build container only):
    /root/reference/risc0/circuit/rv32im-sys/kernels/cxx/rust_poly_fp_{0..3}.cpp   (rv32im, 52.9 k lines)
Output: a `Dag` object; `tools/gen_eval_check.py` lowers and schedules it and emits sm_100a PTX, and also serialises it
to `risc0_b200/circuits/*.ir.json.gz` so the generated kernels can be rebuilt without the reference tree.

Because BabyBear arithmetic is exact, ANY re-association / re-ordering of the DAG yields bit-identical results; the
generator is free to restructure as long as it evaluates the same polynomial. The oracle for the result is the
reference's own compiled C++ (oracle/_ref), compared point by point in tests/.

Statement forms handled (census in SURVEY §7): Fp/FpExt constants, tap loads
`buf[col*steps + ((cycle - kInvRate*back) & mask)]`, global loads `buf[i]`, scratch array loads/stores, `+ - *`,
`acc + v * poly_mix[k]`, `acc + a * b * poly_mix[k]`, tail calls into the next sub-function.
"""
import re
import sys

P = 15 * 2**27 + 1

FP, EXT = 0, 1


class Dag:
    def __init__(self):
        self.nodes = []      # tuples
        self.types = []      # FP / EXT
        self.index = {}
        self.root = None

    def add(self, key, ty):
        i = self.index.get(key)
        if i is None:
            i = len(self.nodes)
            self.nodes.append(key)
            self.types.append(ty)
            self.index[key] = i
        return i

    # constructors with light canonicalisation (commutative operand order) so CSE catches re-loads / re-computations
    def const(self, v):
        return self.add(("c", v % P), FP)

    def const_ext(self, a, b, c, d):
        return self.add(("ce", a % P, b % P, c % P, d % P), EXT)

    def tap(self, buf, col, back):
        return self.add(("t", buf, col, back), FP)

    def glob(self, buf, idx):
        return self.add(("g", buf, idx), FP)

    def pmix(self, k):
        return self.add(("pm", k), EXT)

    def op(self, o, a, b):
        ty = EXT if (self.types[a] == EXT or self.types[b] == EXT) else FP
        if o in "+*" and a > b:
            a, b = b, a
        return self.add((o, a, b), ty)


_re_fn = re.compile(r"^FpExt (\w+)\((.*)\) \{$")
_re_const = re.compile(r"^constexpr Fp (\w+)\((\d+)\);$")
_re_constext = re.compile(r"^constexpr FpExt (\w+)\((\d+),(\d+),(\d+),(\d+)\);$")
_re_ext0 = re.compile(r"^FpExt (\w+) = FpExt\(0\);$")
_re_arr = re.compile(r"^(Fp|FpExt) (\w+)\[(\d+)\];$")
_re_tap = re.compile(r"^auto (\w+) = ([\w\[\]]+)\[(\d+) \* steps \+ \(\(cycle - kInvRate \* (\d+)\) & mask\)\];$")
_re_load = re.compile(r"^auto (\w+) = ([\w\[\]]+)\[(\d+)\];$")
_re_store = re.compile(r"^(\w+)\[(\d+)\] = (\w+);$")
_re_bin = re.compile(r"^auto (\w+) = (\w+) ([-+*]) (\w+);$")
_re_fma = re.compile(r"^FpExt (\w+) = (\w+) \+ (\w+) \* poly_mix\[(\d+)\];$")
_re_fma2 = re.compile(r"^FpExt (\w+) = (\w+) \+ (\w+) \* (\w+) \* poly_mix\[(\d+)\];$")
_re_call = re.compile(r"^auto (\w+) = (\w+)\(cycle, steps, poly_mix, (.*)\);$")
_re_ret = re.compile(r"^return (\w+);$")
_re_comment = re.compile(r"/\*.*?\*/")


def parse_functions(paths):
    fns = {}
    for path in paths:
        cur = None
        for raw in open(path):
            line = raw.strip()
            if not line or line.startswith("//") or line.startswith("#"):
                continue
            line = _re_comment.sub("", line).strip()
            m = _re_fn.match(line)
            if m:
                params = []
                for p in m.group(2).split(","):
                    p = p.strip()
                    ty, name = p.rsplit(" ", 1)
                    params.append((ty.strip(), name))
                cur = (params, [])
                fns[m.group(1)] = cur
                continue
            if cur is None:
                continue
            if line == "}":
                cur = None
                continue
            cur[1].append(line)
    return fns


class ScratchArray:
    def __init__(self, ty, size):
        self.ty = ty
        self.cells = {}


def build_dag(fns, entry="poly_fp", arg_names=("accum", "data", "global", "mix")):
    """`args[k]` of the entry function are the reference's arg order: accum, data, out(global), mix
    (rv32im/src/prove/hal/cpu.rs:178)."""
    dag = Dag()

    def run(name, actuals):
        params, body = fns[name]
        env = {}
        # first three params are cycle, steps, poly_mix
        formal = params[3:]
        assert len(formal) == len(actuals), (name, len(formal), len(actuals))
        for (ty, pname), val in zip(formal, actuals):
            env[pname] = val
        for line in body:
            if line.startswith("size_t mask"):
                continue
            m = _re_const.match(line)
            if m:
                env[m.group(1)] = dag.const(int(m.group(2)))
                continue
            m = _re_constext.match(line)
            if m:
                env[m.group(1)] = dag.const_ext(*(int(m.group(i)) for i in range(2, 6)))
                continue
            m = _re_ext0.match(line)
            if m:
                env[m.group(1)] = dag.const_ext(0, 0, 0, 0)
                continue
            m = _re_arr.match(line)
            if m:
                env[m.group(2)] = ScratchArray(m.group(1), int(m.group(3)))
                continue
            m = _re_tap.match(line)
            if m:
                buf = resolve_buf(env, m.group(2))
                assert isinstance(buf, str), line
                env[m.group(1)] = dag.tap(buf, int(m.group(3)), int(m.group(4)))
                continue
            m = _re_load.match(line)
            if m:
                buf = resolve_buf(env, m.group(2))
                idx = int(m.group(3))
                if isinstance(buf, ScratchArray):
                    env[m.group(1)] = buf.cells[idx]
                else:
                    env[m.group(1)] = dag.glob(buf, idx)
                continue
            m = _re_store.match(line)
            if m:
                arr = env[m.group(1)]
                assert isinstance(arr, ScratchArray), line
                arr.cells[int(m.group(2))] = env[m.group(3)]
                continue
            m = _re_bin.match(line)
            if m:
                env[m.group(1)] = dag.op(m.group(3), env[m.group(2)], env[m.group(4)])
                continue
            m = _re_fma.match(line)
            if m:
                prod = dag.op("*", env[m.group(3)], dag.pmix(int(m.group(4))))
                env[m.group(1)] = dag.op("+", env[m.group(2)], prod)
                continue
            m = _re_fma2.match(line)
            if m:
                inner = dag.op("*", env[m.group(3)], env[m.group(4)])
                prod = dag.op("*", inner, dag.pmix(int(m.group(5))))
                env[m.group(1)] = dag.op("+", env[m.group(2)], prod)
                continue
            m = _re_call.match(line)
            if m:
                actual = []
                for a in m.group(3).split(","):
                    a = a.strip()
                    actual.append(resolve_buf(env, a) if (a.startswith("args[") or isinstance(env.get(a), (str, ScratchArray))) else env[a])
                env[m.group(1)] = run(m.group(2), actual)
                continue
            m = _re_ret.match(line)
            if m:
                return env[m.group(1)]
            raise ValueError("unhandled statement in %s: %s" % (name, line))
        raise ValueError("no return in " + name)

    def resolve_buf(env, tok):
        m = re.match(r"^args\[(\d+)\]$", tok)
        if m:
            return arg_names[int(m.group(1))]
        return env[tok]

    sys.setrecursionlimit(10000)
    # entry has signature (cycle, steps, poly_mix, Fp** args): no extra formals
    params, body = fns[entry]
    fns[entry] = (params[:3], body)
    dag.root = run(entry, [])
    return dag


def stats(dag):
    from collections import Counter
    c = Counter()
    for k, t in zip(dag.nodes, dag.types):
        c[(k[0], "ext" if t == EXT else "fp")] += 1
    return c


def live_nodes(dag):
    """ids reachable from the root"""
    seen = set()
    stack = [dag.root]
    while stack:
        n = stack.pop()
        if n in seen:
            continue
        seen.add(n)
        k = dag.nodes[n]
        if k[0] in "+-*":
            stack.append(k[1])
            stack.append(k[2])
    return seen


if __name__ == "__main__":
    base = "/root/reference/risc0/circuit/rv32im-sys/kernels/cxx/rust_poly_fp_%d.cpp"
    fns = parse_functions([base % i for i in range(4)])
    print("functions:", len(fns))
    dag = build_dag(fns)
    live = live_nodes(dag)
    print("nodes:", len(dag.nodes), "live:", len(live))
    c = stats(dag)
    for k in sorted(c):
        print(k, c[k])
