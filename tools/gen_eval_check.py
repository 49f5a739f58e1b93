#!/usr/bin/env python3
"""Back end of the eval_check code generator: lowers the constraint DAG built by tools/circuit_ir.py to 32-bit
scalar operations, schedules them and emits sm_100a PTX (risc0_b200/csrc/gen/eval_check_<circuit>_p<j>.ptx, assembled
by ptxas and embedded in libr0b200.so by risc0_b200/build.py) plus the host launcher (eval_check_<circuit>.cu).

    python tools/gen_eval_check.py rv32im            # needs /root/reference OR the committed IR
    python tools/gen_eval_check.py rv32im --from-ir  # rebuild from risc0_b200/circuits/rv32im.ir.json.gz only

What the kernel computes is CircuitHal::eval_check (risc0/zkp/src/hal/mod.rs:279-289; CPU spec
risc0/circuit/rv32im/src/prove/hal/cpu.rs:145-208): for every point i of the 4N domain,
    check[k*D + i] = (poly_fp(i) * ((3 * w_4N^i)^N - 1)^-1)[k].
How it is computed differs from both reference back ends:
  * one flat DAG (common sub-expressions merged across the reference's generated sub-functions), lowered to 32-bit
    scalar field operations;
  * the top-level sum is cut into units of bounded cost (oversized terms are split by distributing products over sums)
    and binned into ~45 small part kernels: small straight-line kernels run at about twice the issue efficiency of
    large ones;
  * every `acc + v * poly_mix[k]` chain is flattened into a sum of products and evaluated as a streamed 64-bit
    multiply-accumulate (carry-chained mad.lo.cc / madc.hi pairs that ptxas folds into single IMAD.WIDE accumulates),
    one Montgomery reduction per sum - exact arithmetic makes any re-association bit-identical;
  * values are emitted on demand right before first use; cheap ones and taps are recomputed / re-loaded instead of
    being kept alive; a never-taken branch every 1600 instructions bounds ptxas' scheduling window;
  * poly_mix powers (and their -11 multiples for the X^4 = -11 wrap), globals and mix values travel in the kernel
    parameter block (constant bank), so they are free operands;
  * the divisor only takes 4 values ((3w^i)^N = 3^N * w_4^(i mod 4)): its inverses are computed once on the host.
"""
import gzip
import json
import os
import re
import sys

sys.setrecursionlimit(200000)

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import circuit_ir as ir  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = ir.P
R = 2**32 % P
FP, EXT = ir.FP, ir.EXT


def mont(x):
    return (x % P) * R % P


CIRCUITS = {
    "rv32im": dict(
        srcs=["/root/reference/risc0/circuit/rv32im-sys/kernels/cxx/rust_poly_fp_%d.cpp" % i for i in range(4)],
        arg_names=("accum", "data", "global", "mix"),
        cols=dict(accum=103, data=211, code=1),
        n_global=90, n_mix=36, parts=8,
    ),
    "recursion": dict(
        srcs=["/root/reference/risc0/circuit/recursion-sys/kernels/cxx/poly_fp.cpp"],
        arg_names=("code", "global", "data", "mix", "accum"),   # recursion-sys/kernels/cxx/ffi.cpp:224-230
        cols=dict(accum=12, data=128, code=23),
        n_global=32, n_mix=20, parts=4,
    ),
}


# ------------------------------------------------------------------------------------------------ IR (de)serialise
def save_ir(dag, path):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with gzip.open(path, "wt") as f:
        json.dump(dict(nodes=[list(k) for k in dag.nodes], types=dag.types, root=dag.root), f, separators=(",", ":"))


def load_ir(path):
    with gzip.open(path, "rt") as f:
        d = json.load(f)
    dag = ir.Dag()
    dag.nodes = [tuple(k) for k in d["nodes"]]
    dag.types = d["types"]
    dag.root = d["root"]
    return dag


# ------------------------------------------------------------------------------------------------ partitioning
def reach(dag, roots):
    seen, st = set(), list(roots)
    while st:
        x = st.pop()
        if x in seen:
            continue
        seen.add(x)
        k = dag.nodes[x]
        if k[0] in "+-*":
            st.append(k[1])
            st.append(k[2])
    return seen


def node_cost(dag, x):
    k = dag.nodes[x]
    if k[0] == "t":
        return 1
    if k[0] not in "+-*":
        return 0
    if dag.types[x] == FP:
        return 1
    return 16 if (k[0] == "*" and dag.types[k[1]] == EXT and dag.types[k[2]] == EXT) else 4


def reach_cost(dag, roots):
    return sum(node_cost(dag, x) for x in reach(dag, roots))


def signed_leaves(dag, uses, n):
    """n as a signed sum of ext values if it is a single-use +/- tree: [(sign, node)] in the tree's own order"""
    out, st = [], [(1, n, True)]
    while st:
        sg, x, is_root = st.pop()
        k = dag.nodes[x]
        if k[0] in "+-" and (is_root or uses[x] == 1) and dag.types[x] == EXT and dag.types[k[1]] == EXT and \
                dag.types[k[2]] == EXT:
            st.append((sg if k[0] == "+" else -sg, k[2], False))
            st.append((sg, k[1], False))
        else:
            out.append((sg, x))
    return out


def make_sum(dag, uses, leaves):
    """append nodes for sum(sign * leaf) to the DAG and return the id of the result"""
    def add(key, ty):
        dag.nodes.append(key)
        dag.types.append(ty)
        uses.append(1)
        return len(dag.nodes) - 1
    leaves = sorted(leaves, key=lambda sl: -sl[0])   # a positive leaf first when there is one (stable otherwise)
    sg, acc = leaves[0]
    if sg < 0:
        acc = add(("-", add(("ce", 0, 0, 0, 0), EXT), acc), EXT)
    for sg, x in leaves[1:]:
        acc = add(("+" if sg > 0 else "-", acc, x), EXT)
    return acc


def split_units(dag, uses, limit):
    """The constraint polynomial as a list of independent units (factors, node), each meaning prod(factors) * node,
    whose sum is the root. Units above `limit` cost are split by distributing products over single-use sums:
    a * (b1 + ... + bn) -> a * (b1 + .. + bk) + a * (bk+1 + ...), recomputing the (light) factor in each piece.
    Small kernels matter: measured issue efficiency is ~50 % for 6-10 k-instruction kernels and ~20-25 % for
    20-40 k-instruction ones (gpurun_out/evalcheck_variants7.log)."""
    units = [((), t) for sg, t in signed_leaves(dag, uses, dag.root)]
    assert all(sg > 0 for sg, _ in signed_leaves(dag, uses, dag.root))
    done = []
    while units:
        factors, node = units.pop()
        cost = reach_cost(dag, list(factors) + [node])
        if cost <= limit or len(factors) > 6:
            done.append((factors, node))
            continue
        k = dag.nodes[node]
        leaves = signed_leaves(dag, uses, node)
        if len(leaves) > 1:
            # pack consecutive leaves into chunks under the limit
            chunks, cur = [], []
            for sl in leaves:
                trial = cur + [sl]
                if cur and reach_cost(dag, list(factors) + [x for _, x in trial]) > limit:
                    chunks.append(cur)
                    cur = [sl]
                else:
                    cur = trial
            chunks.append(cur)
            if len(chunks) == 1:      # a single leaf is over the limit on its own: descend into it
                big = max(leaves, key=lambda sl: reach_cost(dag, [sl[1]]))
                rest = [sl for sl in leaves if sl is not big]
                pieces = [[big]] + ([rest] if rest else [])
            else:
                pieces = chunks
            if len(pieces) == 1 and len(pieces[0]) == 1 and pieces[0][0][0] > 0:
                node = pieces[0][0][1]
                k = dag.nodes[node]
            else:
                for piece in pieces:
                    if len(piece) == 1 and piece[0][0] > 0:
                        units.append((factors, piece[0][1]))
                    else:
                        units.append((factors, make_sum(dag, uses, piece)))
                continue
        if k[0] == "*" and dag.types[node] == EXT and (uses[node] <= 1 or not factors):
            a, b = k[1], k[2]
            heavy, light = (a, b) if reach_cost(dag, [a]) >= reach_cost(dag, [b]) else (b, a)
            units.append((factors + (light,), heavy))
            continue
        done.append((factors, node))      # cannot be split further
    return done


def partition(dag, uses, nparts, limit=None):
    """groups of units with balanced cost; sub-expressions shared between groups are recomputed in each"""
    limit = limit or int(os.environ.get("EVAL_UNIT_LIMIT", "3000"))
    units = split_units(dag, uses, limit)
    info = []
    for u in units:
        r = reach(dag, list(u[0]) + [u[1]])
        info.append((sum(node_cost(dag, x) for x in r), u, r))
    total = sum(c for c, _, _ in info)
    if not nparts:
        nparts = max(1, (total + limit - 1) // limit)
    info.sort(key=lambda z: -z[0])
    bins = [dict(terms=[], nodes=set(), cost=0) for _ in range(nparts)]
    for c, u, r in info:
        best, best_cost = None, None
        for b in bins:
            extra = sum(node_cost(dag, x) for x in r - b["nodes"])
            tot = b["cost"] + extra
            if best is None or tot < best_cost:
                best, best_cost = b, tot
        best["terms"].append(u)
        best["cost"] = best_cost
        best["nodes"] |= r
        best.setdefault("reaches", []).append(r)
    # local refinement: move a unit to another bin when that lowers the total recomputed work (more sharing) and the
    # receiving bin stays within 1.25x the unit limit
    def bin_cost(reaches):
        nodes = set()
        for r in reaches:
            nodes |= r
        return sum(node_cost(dag, x) for x in nodes)

    if os.environ.get("EVAL_REFINE", "1") == "1":
        refine_cap = float(os.environ.get("EVAL_REFINE_CAP", "1.25"))
        for _ in range(int(os.environ.get("EVAL_REFINE_PASSES", "2"))):
            moved = 0
            for b in bins:
                j = 0
                while j < len(b["terms"]) and len(b["terms"]) > 1:
                    r = b["reaches"][j]
                    without = bin_cost(b["reaches"][:j] + b["reaches"][j + 1:])
                    gain_here = b["cost"] - without
                    best_t, best_delta, best_new = None, 0, None
                    for t in bins:
                        if t is b or not t["terms"]:
                            continue
                        new_cost = t["cost"] + sum(node_cost(dag, x) for x in r - t["nodes"])
                        delta = (new_cost - t["cost"]) - gain_here
                        if delta < best_delta and new_cost <= refine_cap * limit:
                            best_t, best_delta, best_new = t, delta, new_cost
                    if best_t is not None:
                        best_t["terms"].append(b["terms"].pop(j))
                        best_t["reaches"].append(b["reaches"].pop(j))
                        best_t["nodes"] |= r
                        best_t["cost"] = best_new
                        b["cost"] = without
                        b["nodes"] = set().union(*b["reaches"]) if b["reaches"] else set()
                        moved += 1
                    else:
                        j += 1
            if not moved:
                break
        # dissolve bins that ended up tiny (a launch plus the check read-modify-write costs more than they compute)
        for b in sorted(bins, key=lambda b: b["cost"]):
            if not b["terms"] or b["cost"] >= 0.3 * limit:
                continue
            if not any(t is not b and t["terms"] for t in bins):
                continue     # a single bin: nothing to merge into
            for u, r in list(zip(b["terms"], b["reaches"])):
                cands = [t for t in bins if t is not b and t["terms"]]
                t = min(cands, key=lambda t: (t["cost"] + sum(node_cost(dag, x) for x in r - t["nodes"]) > 1.25 * limit,
                                              sum(node_cost(dag, x) for x in r - t["nodes"])))
                t["cost"] += sum(node_cost(dag, x) for x in r - t["nodes"])
                t["terms"].append(u)
                t["reaches"].append(r)
                t["nodes"] |= r
            b["terms"], b["reaches"], b["nodes"], b["cost"] = [], [], set(), 0
    return [b for b in bins if b["terms"]]


# ------------------------------------------------------------------------------------------------ scalar lowering
NINV = 0x77FFFFFF            # -P^-1 mod 2^32
PINV = 0x88000001            # P^-1 mod 2^32
MONT_ONE = R
PLAIN_ADD = os.environ.get("EVAL_PLAIN_ADD", "1") == "1"
ADDR_TABLE = os.environ.get("EVAL_ADDR_TABLE", "0") == "1"
RINV = pow(R, -1, P)
NBETA_M = mont(P - 11)


def mont_mul(a, b):
    return a * b * RINV % P


class Scalars:
    """hash-consed DAG of 32-bit field operations (every value canonical Montgomery, < P)"""

    def __init__(self):
        self.nodes, self.index = [], {}

    def add_node(self, key):
        i = self.index.get(key)
        if i is None:
            i = len(self.nodes)
            self.nodes.append(key)
            self.index[key] = i
        return i

    def imm(self, v):
        return self.add_node(("i", v % P))

    def cst(self, off):
        return self.add_node(("k", off))

    def tap(self, buf, col, back):
        return self.add_node(("t", buf, col, back))

    def is_imm(self, a, v=None):
        k = self.nodes[a]
        return k[0] == "i" and (v is None or k[1] == v)

    def add(self, a, b):
        if self.is_imm(a, 0):
            return b
        if self.is_imm(b, 0):
            return a
        if self.is_imm(a) and self.is_imm(b):
            return self.imm(self.nodes[a][1] + self.nodes[b][1])
        if a > b:
            a, b = b, a
        return self.add_node(("+", a, b))

    def sub(self, a, b):
        if self.is_imm(b, 0):
            return a
        if a == b:
            return self.imm(0)
        if self.is_imm(a) and self.is_imm(b):
            return self.imm(self.nodes[a][1] - self.nodes[b][1])
        if self.is_imm(a, 0):
            return self.add_node(("n", b))
        return self.add_node(("-", a, b))

    def neg(self, a):
        return self.sub(self.imm(0), a)

    def mul(self, a, b):
        if self.is_imm(a, 0) or self.is_imm(b, 0):
            return self.imm(0)
        if self.is_imm(a, MONT_ONE):
            return b
        if self.is_imm(b, MONT_ONE):
            return a
        if self.is_imm(a) and self.is_imm(b):
            return self.imm(mont_mul(self.nodes[a][1], self.nodes[b][1]))
        if a > b:
            a, b = b, a
        return self.add_node(("*", a, b))

    def dot(self, terms):
        terms = [(min(a, b), max(a, b)) for a, b in terms if not (self.is_imm(a, 0) or self.is_imm(b, 0))]
        if not terms:
            return self.imm(0)
        if len(terms) == 1:
            return self.mul(*terms[0])
        return self.add_node(("d", tuple(sorted(terms))))

    def operands(self, i):
        k = self.nodes[i]
        if k[0] in "+-*":
            return [k[1], k[2]]
        if k[0] in ("n", "N"):
            return [k[1]]
        if k[0] == "d":
            return [x for ab in k[1] for x in ab]
        return []


class Layout:
    """byte offsets inside the kernel's constant parameter block"""

    def __init__(self, npm, n_global, n_mix, ncols=0):
        self.pm = 0
        self.npm = 16 * npm
        self.glob = 32 * npm
        self.mix = self.glob + 4 * n_global
        self.inv_y = self.mix + 4 * n_mix
        self.size = self.inv_y + 16
        self.n = npm
        # EVAL_ADDR_TABLE: byte offsets col * stride of the tap columns as 64-bit constants, so that a tap address is
        # base + c[...] (IADD3 + IADD3.X with constant-bank operands) instead of an IMAD.WIDE on the fma-heavy pipe
        self.ncols = ncols if ADDR_TABLE else 0
        if self.ncols:
            self.coloff = (self.size + 7) & ~7
            self.size = self.coloff + 8 * self.ncols


def lower(dag, terms, lay):
    """ext-level DAG -> scalar DAG; returns (Scalars, [4 output scalars])"""
    S = Scalars()
    memo = {}
    zero = S.imm(0)

    def nbeta_of(s):
        k = S.nodes[s]
        if k[0] == "k" and lay.pm <= k[1] < lay.npm:
            return S.cst(k[1] + lay.npm)       # the precomputed (-11 * pm[k][c]) copy
        return S.mul(s, S.imm(NBETA_M))

    def ext_mul(x, y):
        # prefer the side made of constants as `y` so that its -11 multiples are free
        def constness(v):
            return sum(1 for s in v if S.nodes[s][0] in "ik")
        if constness(x) > constness(y):
            x, y = y, x
        nb, nc, nd = nbeta_of(y[1]), nbeta_of(y[2]), nbeta_of(y[3])
        return (S.dot([(x[0], y[0]), (x[1], nd), (x[2], nc), (x[3], nb)]),
                S.dot([(x[0], y[1]), (x[1], y[0]), (x[2], nd), (x[3], nc)]),
                S.dot([(x[0], y[2]), (x[1], y[1]), (x[2], y[0]), (x[3], nd)]),
                S.dot([(x[0], y[3]), (x[1], y[2]), (x[2], y[1]), (x[3], y[0])]))

    def get(n):
        """iterative post-order evaluation; FP nodes -> scalar id, EXT nodes -> 4-tuple"""
        stack = [n]
        while stack:
            cur = stack[-1]
            if cur in memo:
                stack.pop()
                continue
            k = dag.nodes[cur]
            o = k[0]
            if o in "+-*":
                pend = [x for x in (k[1], k[2]) if x not in memo]
                if pend:
                    stack.extend(pend)
                    continue
                a, b = memo[k[1]], memo[k[2]]
                ea, eb = isinstance(a, tuple), isinstance(b, tuple)
                if not ea and not eb:
                    r = {"+": S.add, "-": S.sub, "*": S.mul}[o](a, b)
                elif ea and eb:
                    if o == "*":
                        r = ext_mul(a, b)
                    else:
                        f = S.add if o == "+" else S.sub
                        r = tuple(f(a[c], b[c]) for c in range(4))
                elif o == "*":
                    e, f = (a, b) if ea else (b, a)
                    r = tuple(S.mul(e[c], f) for c in range(4))
                elif o == "+":
                    e, f = (a, b) if ea else (b, a)
                    r = (S.add(e[0], f), e[1], e[2], e[3])
                elif ea:   # ext - fp
                    r = (S.sub(a[0], b), a[1], a[2], a[3])
                else:      # fp - ext
                    r = (S.sub(a, b[0]), S.neg(b[1]), S.neg(b[2]), S.neg(b[3]))
            elif o == "c":
                r = S.imm(mont(k[1]))
            elif o == "ce":
                r = tuple(S.imm(mont(v)) for v in k[1:5])
            elif o == "pm":
                r = tuple(S.cst(lay.pm + 16 * k[1] + 4 * c) for c in range(4))
            elif o == "g":
                base = lay.glob if k[1] == "global" else lay.mix
                r = S.cst(base + 4 * k[2])
            elif o == "t":
                r = S.tap(k[1], k[2], k[3])
            else:
                raise ValueError(k)
            memo[cur] = r
            stack.pop()
        return memo[n]

    def as_ext(v):
        return v if isinstance(v, tuple) else (v, zero, zero, zero)

    tot = [zero] * 4
    for unit in terms:
        factors, node = unit if isinstance(unit, tuple) else ((), unit)
        v = get(node)
        for f in factors:
            fv = get(f)
            if isinstance(v, tuple) and isinstance(fv, tuple):
                v = ext_mul(fv, v)
            elif isinstance(v, tuple):
                v = tuple(S.mul(v[c], fv) for c in range(4))
            elif isinstance(fv, tuple):
                v = tuple(S.mul(fv[c], v) for c in range(4))
            else:
                v = S.mul(fv, v)
        v = as_ext(v)
        tot = [S.add(tot[c], v[c]) for c in range(4)]
    return S, tot


def flatten_sums(S, outs):
    """v2 optimisation: turn single-use add/sub trees (leaves: single-use products or plain values) into one lazy
    64-bit dot product. Exact arithmetic -> identical results. Returns (new Scalars, new outs)."""
    n = len(S.nodes)
    uses = [0] * n
    for i in range(n):
        for o in S.operands(i):
            uses[o] += 1
    for o in outs:
        uses[o] += 1
    T = Scalars()
    memo = {}
    one = ("i", MONT_ONE)

    def leaves(root):
        """[(sign, node)] of the maximal single-use add/sub tree under root"""
        out, st = [], [(1, root, True)]
        while st:
            sg, x, is_root = st.pop()
            k = S.nodes[x]
            if (is_root or uses[x] == 1) and k[0] in "+-":
                # push right first so that the left operand (the earlier part of the chain) is visited first:
                # the leaves come out in the generator's original constraint order, which has good operand locality
                st.append((sg if k[0] == "+" else -sg, k[2], False))
                st.append((sg, k[1], False))
            elif (is_root or uses[x] == 1) and k[0] == "n":
                st.append((-sg, k[1], False))
            else:
                out.append((sg, x))
        return out

    def conv(root):
        stack = [root]
        while stack:
            cur = stack[-1]
            if cur in memo:
                stack.pop()
                continue
            k = S.nodes[cur]
            lv = leaves(cur) if k[0] in "+-n" else []
            if len(lv) >= 3:
                # product leaves used once are folded in as terms; their operands are what we need converted
                need = []
                for sg, x in lv:
                    kx = S.nodes[x]
                    if kx[0] == "*" or (uses[x] == 1 and kx[0] == "d"):
                        need.extend(S.operands(x))
                    else:
                        need.append(x)
                pend = [x for x in need if x not in memo]
                if pend:
                    stack.extend(pend)
                    continue
                terms, plain = [], []
                for sg, x in lv:
                    kx = S.nodes[x]
                    if kx[0] == "*":      # a product costs one MAC wherever it is used: never worth a register
                        prods = [(memo[kx[1]], memo[kx[2]])]
                    elif uses[x] == 1 and kx[0] == "d":
                        prods = [(memo[a], memo[b]) for a, b in kx[1]]
                    elif PLAIN_ADD:
                        # a plain addend joins AFTER the reduction (IADD3 + VIADDMNMX on the alu pipe) instead of
                        # entering the dot product as x * 1 (an IMAD.WIDE on the fma-heavy pipe, the saturated one)
                        plain.append((sg, memo[x]))
                        continue
                    else:
                        prods = [(memo[x], T.imm(MONT_ONE))]
                    for a, b in prods:
                        if sg < 0:
                            # negate the cheaper side: an immediate is free, otherwise one subtraction
                            if T.is_imm(b):
                                b = T.imm(-T.nodes[b][1])
                            elif T.is_imm(a):
                                a = T.imm(-T.nodes[a][1])
                            else:
                                b = T.add_node(("N", b))   # lazy negation P - b (may equal P; fine inside a product)
                        terms.append((a, b))
                r = T.add_node(("d", tuple(terms))) if terms else None
                for sg, v in plain:
                    if r is None:
                        r = v if sg > 0 else T.neg(v)
                    else:
                        r = T.add(r, v) if sg > 0 else T.sub(r, v)
            else:
                ops = S.operands(cur)
                pend = [x for x in ops if x not in memo]
                if pend:
                    stack.extend(pend)
                    continue
                if k[0] == "+":
                    r = T.add(memo[k[1]], memo[k[2]])
                elif k[0] == "-":
                    r = T.sub(memo[k[1]], memo[k[2]])
                elif k[0] == "n":
                    r = T.neg(memo[k[1]])
                elif k[0] == "*":
                    r = T.mul(memo[k[1]], memo[k[2]])
                elif k[0] == "d":
                    r = T.add_node(("d", tuple((memo[a], memo[b]) for a, b in k[1])))
                else:
                    r = T.add_node(k)
            memo[cur] = r
            stack.pop()
        return memo[root]

    new_outs = [conv(o) for o in outs]
    return T, new_outs


def schedule_scalars(S, outs):
    n = len(S.nodes)

    def ops(i):
        k = S.nodes[i]
        if k[0] == "N":
            return [k[1]]
        return S.operands(i)

    weight = [1] * n
    for i in range(n):
        w = 1
        for o in ops(i):
            w += weight[o]
        weight[i] = min(w, 10**9)
    order, seen = [], set()
    for root in outs:
        stack = [(root, False)]
        while stack:
            node, done = stack.pop()
            if done:
                order.append(node)
                continue
            if node in seen:
                continue
            seen.add(node)
            stack.append((node, True))
            for o in sorted(set(ops(node)), key=lambda x: weight[x]):   # heaviest pushed last -> visited first
                stack.append((o, False))
    return order


# ------------------------------------------------------------------------------------------------ PTX emission
_TUNE = None


def part_opt(part, key, default):
    """generator option for one part kernel: the measured per-part choice (risc0_b200/circuits/<circuit>.tune.json,
    written by tools/autotune_eval_check.py) unless EVAL_IGNORE_TUNE=1, else the environment, else the default.
    Every option only changes the ORDER / placement of exact field operations: results are bit-identical."""
    global _TUNE
    if _TUNE is None:
        _TUNE = {}
        if os.environ.get("EVAL_IGNORE_TUNE") != "1":
            import glob
            for f in glob.glob(os.path.join(ROOT, "risc0_b200", "circuits", "*.tune.json")):
                _TUNE.update(json.load(open(f)).get("gen", {}))
    v = _TUNE.get(part, {}).get(key)
    if v is None:
        v = os.environ.get(key, default)
    return v


class Ptx:
    def __init__(self, S, name, lay):
        self.S, self.name, self.lay = S, name, lay
        self.body = []
        self.nt = 0   # u32 temporaries
        self.nw = 0   # u64 temporaries
        self.nl = 0   # tap load registers
        self.pos = 0
        self.remat = int(part_opt(name, "EVAL_REMAT_DIST", "400"))   # re-load / recompute instead of keeping alive
        self.remat_cost = int(part_opt(name, "EVAL_REMAT_COST", "8"))
        self.split = int(part_opt(name, "EVAL_SPLIT", "1600"))
        self.last_fence = 0
        self.use_bar = os.environ.get("EVAL_BAR", "0") == "1"
        self.first_wide = os.environ.get("EVAL_FIRST_WIDE", "1") == "1"
        self.mshift = os.environ.get("EVAL_MSHIFT", "0") == "1"
        self.pin_adds = os.environ.get("EVAL_PIN", "0") == "1"
        self.additive = part_opt(name, "EVAL_ADDITIVE", "0") == "1"
        self.bases = {}

    def t(self):
        self.nt += 1
        return "%%t%d" % self.nt

    def w(self):
        self.nw += 1
        return "%%w%d" % self.nw

    def opnd(self, i):
        k = self.S.nodes[i]
        if k[0] == "i":
            return str(k[1])
        return "%%v%d" % i

    def bound(self, i):
        k = self.S.nodes[i]
        if k[0] == "i":
            return k[1]
        if k[0] == "N":
            return P
        return P - 1

    def emit(self, line):
        self.body.append("    " + line)
        self.pos += 1

    def fence(self):
        """basic-block boundary (a never-taken branch): keeps ptxas' scheduler from hoisting work across it, which is
        what blows up register pressure on a 30k-instruction block"""
        if self.split and self.pos - self.last_fence >= self.split:
            # bar.sync also keeps the warps of a block within the same instruction-cache window: straight-line code
            # is fetched once per block instead of once per warp
            self.body.append("    bar.sync 0;" if self.use_bar else "    @%p0 bra DONE;")
            self.last_fence = self.pos

    def fold_hi(self, hi, bound):
        """bring a 64-bit accumulator (lo, hi) with value <= bound under P * 2^32 by reducing its high word in place;
        returns the new bound"""
        lim = P << 32
        if bound < lim:
            return bound
        if bound >= 2 * lim:
            h2 = self.t()
            self.emit("add.u32 %s, %s, %d;" % (h2, hi, (1 << 32) - 2 * P))
            self.emit("min.u32 %s, %s, %s;" % (hi, hi, h2))
        h4 = self.t()
        self.emit("add.u32 %s, %s, %d;" % (h4, hi, (1 << 32) - P))
        self.emit("min.u32 %s, %s, %s;" % (hi, hi, h4))
        return lim - 1

    def mac(self, lo, hi, a, b, first):
        """(lo, hi) += a * b as the carry-chained 32-bit pair that ptxas folds into ONE IMAD.WIDE with a 64-bit
        accumulator operand (a mad.wide chain is instead re-associated into IMAD.WIDE + IADD3/IADD3.X trees, which
        puts ~1.5 extra instructions per product on the half-rate alu pipe)"""
        if first and self.first_wide:
            # one IMAD.WIDE for certain: ptxas leaves ~2/3 of the mul.lo / mul.hi pairs as IMAD + IMAD.HI (6 fma-pipe
            # cycles instead of 4)
            w = self.w()
            self.emit("mul.wide.u32 %s, %s, %s;" % (w, a, b))
            self.emit("mov.b64 {%s, %s}, %s;" % (lo, hi, w))
        elif first:
            self.emit("mul.lo.u32 %s, %s, %s;" % (lo, a, b))
            self.emit("mul.hi.u32 %s, %s, %s;" % (hi, a, b))
        else:
            self.emit("mad.lo.cc.u32 %s, %s, %s, %s;" % (lo, a, b, lo))
            self.emit("madc.hi.u32 %s, %s, %s, %s;" % (hi, a, b, hi))

    def pin(self, r):
        """r += 0 with a zero ptxas cannot see through: the preceding two-input add / sub becomes a three-input IADD3,
        which ptxas cannot re-issue as IMAD.IADD on the fma pipe (it balances instruction COUNTS between the integer
        pipes and does not know that IMAD.WIDE / IMAD.HI hold the fma pipe twice as long as an IMAD)"""
        if self.pin_adds:
            self.emit("add.u32 %s, %s, %%zr;" % (r, r))

    def mont_finish(self, dst, lo, hi):
        # subtractive Montgomery reduction (csrc/fp.cuh mont_reduce): m = lo * P^-1; r = hi - hi(m * P) in (-P, P);
        # canonical = min.u32(r, r + P). No carry chain, 4 instructions.
        m, h, r, r2 = self.t(), self.t(), self.t(), self.t()
        if self.additive:
            # additive form: m = lo * (-P^-1); (lo, hi) + m * P has a zero low word, its high word is the result in
            # [0, 2P) (the accumulator is < P 2^32 and m P < 2^32 P, so the sum stays under 2^64). The carry-chained pair
            # folds into ONE IMAD.WIDE with a 64-bit addend: IMAD + IMAD.WIDE + VIADDMNMX instead of IMAD + IMAD.HI +
            # IADD3 + VIADDMNMX - one alu-pipe instruction less per reduction.
            z = self.t()
            self.emit("mul.lo.u32 %s, %s, %d;" % (m, lo, NINV))
            self.emit("mad.lo.cc.u32 %s, %s, %d, %s;" % (z, m, P, lo))
            self.emit("madc.hi.u32 %s, %s, %d, %s;" % (r, m, P, hi))
            self.emit("add.u32 %s, %s, %d;" % (r2, r, (1 << 32) - P))
            self.emit("min.u32 %s, %s, %s;" % (dst, r, r2))
            return
        if self.mshift:
            # m = lo * P^-1 with P^-1 = 1 + 2^27 + 2^31 as two shifts and one three-input add on the alu pipe (which is
            # half idle in these kernels) instead of an IMAD on the saturated fma pipe. The shift amounts are opaque
            # registers (27 / 31 + (domain & 1)): with immediates ptxas turns the shift-adds back into IMADs.
            a, b, c = self.t(), self.t(), self.t()
            self.emit("shl.b32 %s, %s, %%sh27;" % (a, lo))
            self.emit("shl.b32 %s, %s, %%sh31;" % (b, lo))
            self.emit("add.u32 %s, %s, %s;" % (c, lo, a))
            self.emit("add.u32 %s, %s, %s;" % (m, c, b))
        else:
            self.emit("mul.lo.u32 %s, %s, %d;" % (m, lo, PINV))
        self.emit("mul.hi.u32 %s, %s, %d;" % (h, m, P))
        self.emit("sub.u32 %s, %s, %s;" % (r, hi, h))
        self.pin(r)
        self.emit("add.u32 %s, %s, %d;" % (r2, r, P))
        self.emit("min.u32 %s, %s, %s;" % (dst, r, r2))

    def node(self, i):
        S = self.S
        k = S.nodes[i]
        o = k[0]
        dst = self.reg.get(i, "%%v%d" % i)
        if o == "i":
            return
        if o == "k":
            self.emit("ld.param.u32 %s, [p_cst+%d];" % (dst, k[1]))
        elif o in "+-":
            a, b = self.use(k[1]), self.use(k[2])
            if S.nodes[k[1]][0] == "i":   # immediate first operand: materialise (only for '-', '+' is normalised)
                m = self.t()
                self.emit("mov.u32 %s, %s;" % (m, a))
                a = m
            t1, t2 = self.t(), self.t()
            if o == "+":
                self.emit("add.u32 %s, %s, %s;" % (t1, a, b))
                self.pin(t1)
                self.emit("add.u32 %s, %s, %d;" % (t2, t1, (1 << 32) - P))
            else:
                self.emit("sub.u32 %s, %s, %s;" % (t1, a, b))
                self.pin(t1)
                self.emit("add.u32 %s, %s, %d;" % (t2, t1, P))
            self.emit("min.u32 %s, %s, %s;" % (dst, t1, t2))
        elif o == "n":
            t1, t2 = self.t(), self.t()
            self.emit("neg.s32 %s, %s;" % (t1, self.use(k[1])))
            self.emit("add.u32 %s, %s, %d;" % (t2, t1, P))
            self.emit("min.u32 %s, %s, %s;" % (dst, t1, t2))
        elif o == "N":
            self.emit("sub.u32 %s, %d, %s;" % (dst, P, self.use(k[1])))
        elif o == "*":
            lo, hi = self.t(), self.t()
            a, b = k[1], k[2]
            if S.nodes[a][0] == "i":
                a, b = b, a
            self.mac(lo, hi, self.use(a), self.use(b), True)
            self.mont_finish(dst, lo, hi)
        else:
            raise ValueError(k)

    # ---- demand-driven emission: a value is computed right before its first use; sums of products are streamed
    # (operands of term j are computed, then multiplied into the running accumulators of every sibling sum that uses
    # them) so that a long constraint sum never has more than one term's worth of temporaries alive.
    def prepare(self):
        S = self.S
        n = len(S.nodes)
        self.weight = [1] * n
        for i in range(n):
            w = 1
            for o in self.ops(i):
                w += self.weight[o]
            self.weight[i] = min(w, 10**9)
        self.done = set()
        self.pos = 0
        self.tap_state = {}
        self.reg = {}        # node -> current register name
        self.last_use = {}
        self.ncopy = 0
        # cost of recomputing a value from taps / constants (instructions), capped
        self.cost = [0] * n
        for i in range(n):
            k = S.nodes[i]
            if k[0] in "ik":
                c = 0
            elif k[0] == "t":
                c = 2
            else:
                c = {"+": 3, "-": 3, "n": 3, "N": 1, "*": 6}.get(k[0], 0)
                if k[0] == "d":
                    c = 2 * len(k[1]) + 6
                c += sum(self.cost[o] for o in set(self.ops(i)))
            self.cost[i] = min(c, 10**6)
        self.siblings = {}
        for i, k in enumerate(S.nodes):
            if k[0] == "d":
                self.siblings.setdefault(self.signature(k), []).append(i)

    def ops(self, i):
        k = self.S.nodes[i]
        if k[0] == "N":
            return [k[1]]
        return self.S.operands(i)

    def is_const(self, i):
        return self.S.nodes[i][0] in "ik"

    def signature(self, k):
        return tuple(sorted(set(x for ab in k[1] for x in ab if not self.is_const(x))))

    def ensure(self, i):
        if i in self.done:
            return
        S = self.S
        k = S.nodes[i]
        if k[0] == "d":
            self.stream_dots([d for d in self.siblings[self.signature(k)] if d not in self.done])
            return
        if k[0] != "t":
            for o in sorted(set(self.ops(i)), key=lambda x: -self.weight[x]):
                self.ensure(o)
            self.node(i)
            self.fence()
        self.done.add(i)

    def use(self, i):
        """operand string for node i at the current position (taps are re-loaded when the last load is far behind)"""
        k = self.S.nodes[i]
        if k[0] == "i":
            return str(k[1])
        if k[0] == "t":
            st = self.tap_state.get(i)
            if st is None or self.pos - st[1] > self.remat:
                reg = "%%l%d" % self.nl
                self.nl += 1
                base = "%%b_%s_%d" % (k[1], k[3])
                self.bases[(k[1], k[3])] = base
                a = self.w()
                if self.lay.ncols:
                    o = self.w()
                    self.emit("ld.param.u64 %s, [p_cst+%d];" % (o, self.lay.coloff + 8 * k[2]))
                    self.emit("add.u64 %s, %s, %s;" % (a, base, o))
                elif self.pin_adds:
                    # explicit uniform product + carry-chained add: with every other add pinned to the alu pipe ptxas
                    # would otherwise fold this into a per-thread IMAD.WIDE (4 fma-pipe cycles per tap load)
                    o, ol, oh, bl, bh = self.w(), self.t(), self.t(), self.t(), self.t()
                    self.emit("mul.wide.u32 %s, %%stride, %d;" % (o, k[2]))
                    self.emit("mov.b64 {%s, %s}, %s;" % (ol, oh, o))
                    self.emit("mov.b64 {%s, %s}, %s;" % (bl, bh, base))
                    self.emit("add.cc.u32 %s, %s, %s;" % (ol, ol, bl))
                    self.emit("addc.u32 %s, %s, %s;" % (oh, oh, bh))
                    self.emit("mov.b64 %s, {%s, %s};" % (a, ol, oh))
                else:
                    self.emit("mad.wide.u32 %s, %%stride, %d, %s;" % (a, k[2], base))
                self.emit("ld.global.nc.u32 %s, [%s];" % (reg, a))
                st = [reg, self.pos]
                self.tap_state[i] = st
            st[1] = self.pos
            return st[0]
        if k[0] != "k" and self.cost[i] <= self.remat_cost and self.pos - self.last_use.get(i, self.pos) > self.remat:
            # cheap value whose previous use is far behind: recompute into a fresh register instead of keeping it alive
            self.ncopy += 1
            self.reg[i] = "%%c%d" % self.ncopy
            self.node(i)
        self.last_use[i] = self.pos
        return self.reg.get(i, "%%v%d" % i)

    def stream_dots(self, dots):
        S = self.S
        # align the siblings' terms by their non-constant operands
        keyed = {}
        for d in dots:
            for a, b in S.nodes[d][1]:
                if S.nodes[a][0] == "i":
                    a, b = b, a
                key = tuple(sorted(x for x in (a, b) if not self.is_const(x)))
                keyed.setdefault(key, []).append((d, a, b))
        acc = {d: None for d in dots}      # (lo, hi) register pair once the first product has been issued
        bound = {d: 0 for d in dots}
        if part_opt(self.name, "EVAL_ORDER", "program") == "weight":
            order = sorted(keyed, key=lambda key: -max([self.weight[x] for x in key] or [0]))
        else:
            order = list(keyed)    # first-appearance order = the constraint system's own order
        for key in order:
            for x in sorted(key, key=lambda x: -self.weight[x]):
                self.ensure(x)
            for d, a, b in keyed[key]:
                for x in (a, b):
                    self.ensure(x)      # constants: ld.param emitted on first use
                pb = self.bound(a) * self.bound(b)
                if acc[d] is not None and bound[d] + pb >= (1 << 64):
                    bound[d] = self.fold_hi(acc[d][1], bound[d])
                ua, ub = self.use(a), self.use(b)
                if acc[d] is None:
                    acc[d] = (self.t(), self.t())
                    self.mac(acc[d][0], acc[d][1], ua, ub, True)
                else:
                    self.mac(acc[d][0], acc[d][1], ua, ub, False)
                bound[d] += pb
        for d in dots:
            self.fold_hi(acc[d][1], bound[d])
            self.mont_finish("%%v%d" % d, acc[d][0], acc[d][1])
            self.done.add(d)
        self.fence()

    def kernel(self, order, outs, threads):
        self.prepare()
        for o in outs:
            self.ensure(o)
        nv = len(self.S.nodes)
        L = []
        L.append("// GENERATED by tools/gen_eval_check.py - do not edit.")
        L.append(".version 8.6")
        L.append(".target sm_100a")
        L.append(".address_size 64")
        L.append("")
        L.append(".visible .entry %s(" % self.name)
        L.append("    .param .u64 p_check, .param .u64 p_accum, .param .u64 p_code, .param .u64 p_data, .param .u32 p_domain, .param .u32 p_first, .param .u32 p_i0,")
        L.append("    .param .align 16 .b8 p_cst[%d])" % self.lay.size)
        L.append(".maxntid %d, 1, 1" % threads)
        if os.environ.get("EVAL_MAXNREG"):
            L.append(".maxnreg %d" % int(os.environ["EVAL_MAXNREG"]))
        L.append("{")
        L.append("    .reg .pred %p<4>;")
        L.append("    .reg .u32 %%v<%d>;" % (nv + 1))
        L.append("    .reg .u32 %%t<%d>;" % (self.nt + 40))
        L.append("    .reg .u32 %%l<%d>;" % (self.nl + 1))
        L.append("    .reg .u32 %%c<%d>;" % (self.ncopy + 2))
        L.append("    .reg .u64 %%w<%d>;" % (self.nw + 40))
        L.append("    .reg .u32 %i, %i0, %domain, %mask, %first, %stride, %bdim, %bidx, %lane;")
        L.append("    .reg .u32 %o<8>, %q<8>, %iy<6>, %res<4>, %sh27, %sh31, %zr;")
        L.append("    .reg .u64 %check, %accum, %code, %data, %oaddr<4>, %off;")
        for (buf, back), base in sorted(self.bases.items()):
            L.append("    .reg .u64 %s;" % base)
        if int(os.environ.get("EVAL_PREFETCH", "0")):
            L.append("    .reg .u64 %pfa;")
        L.append("    ld.param.u64 %check, [p_check];")
        L.append("    ld.param.u64 %accum, [p_accum];")
        L.append("    ld.param.u64 %code, [p_code];")
        L.append("    ld.param.u64 %data, [p_data];")
        L.append("    ld.param.u32 %domain, [p_domain];")
        L.append("    ld.param.u32 %first, [p_first];")
        L.append("    ld.param.u32 %i0, [p_i0];")
        L.append("    cvta.to.global.u64 %check, %check;")
        L.append("    cvta.to.global.u64 %accum, %accum;")
        L.append("    cvta.to.global.u64 %code, %code;")
        L.append("    cvta.to.global.u64 %data, %data;")
        L.append("    mov.u32 %bdim, %ntid.x;")
        L.append("    mov.u32 %bidx, %ctaid.x;")
        L.append("    mov.u32 %lane, %tid.x;")
        L.append("    mad.lo.u32 %i, %bidx, %bdim, %lane;")
        L.append("    add.u32 %i, %i, %i0;")
        L.append("    setp.ge.u32 %p1, %i, %domain;")
        L.append("    @%p1 bra DONE;")
        L.append("    add.u32 %mask, %domain, -1;")
        L.append("    shl.b32 %stride, %domain, 2;")
        L.append("    setp.eq.u32 %p0, %domain, 3;")
        if self.mshift or self.pin_adds:
            L.append("    and.b32 %zr, %domain, 1;")     # 0 (the domain is a power of two >= 4), unknown to ptxas
            L.append("    add.u32 %sh31, %zr, 31;")
            L.append("    add.u32 %sh27, %zr, 27;")
        n = 0
        for (buf, back), base in sorted(self.bases.items()):
            # element index (i - 4*back) & mask, byte address = buf + 4 * index
            L.append("    add.u32 %%q0, %%i, %d;" % ((-4 * back) % (1 << 32)))
            L.append("    and.b32 %q0, %q0, %mask;")
            L.append("    mad.wide.u32 %s, %%q0, 4, %%%s;" % (base, buf))
            n += 1
        L.extend(self.with_prefetch(self.body))
        # scale by inv_y[i & 3], accumulate into check unless this is the first part
        L.append("    and.b32 %q1, %i, 3;")
        L.append("    ld.param.u32 %%iy0, [p_cst+%d];" % self.lay.inv_y)
        L.append("    ld.param.u32 %%iy1, [p_cst+%d];" % (self.lay.inv_y + 4))
        L.append("    ld.param.u32 %%iy2, [p_cst+%d];" % (self.lay.inv_y + 8))
        L.append("    ld.param.u32 %%iy3, [p_cst+%d];" % (self.lay.inv_y + 12))
        L.append("    setp.eq.u32 %p2, %q1, 1;")
        L.append("    selp.u32 %iy4, %iy1, %iy0, %p2;")
        L.append("    setp.eq.u32 %p2, %q1, 2;")
        L.append("    selp.u32 %iy4, %iy2, %iy4, %p2;")
        L.append("    setp.eq.u32 %p2, %q1, 3;")
        L.append("    selp.u32 %iy4, %iy3, %iy4, %p2;")
        self.body = []
        for c in range(4):
            lo, hi = self.t(), self.t()
            self.mac(lo, hi, self.use(outs[c]) if self.S.nodes[outs[c]][0] != "i" else self._mat(outs[c]), "%iy4", True)
            self.mont_finish("%%res%d" % c, lo, hi)
        L.extend(self.body)
        L.append("    mul.wide.u32 %off, %i, 4;")
        L.append("    add.u64 %oaddr0, %check, %off;")
        L.append("    cvt.u64.u32 %off, %stride;")
        L.append("    add.u64 %oaddr1, %oaddr0, %off;")
        L.append("    add.u64 %oaddr2, %oaddr1, %off;")
        L.append("    add.u64 %oaddr3, %oaddr2, %off;")
        L.append("    setp.ne.u32 %p3, %first, 0;")
        L.append("    @%p3 bra STORE;")
        for c in range(4):
            L.append("    ld.global.u32 %%o%d, [%%oaddr%d];" % (c, c))
        for c in range(4):
            L.append("    add.u32 %%q2, %%res%d, %%o%d;" % (c, c))
            L.append("    add.u32 %%q3, %%q2, %d;" % ((1 << 32) - P))
            L.append("    min.u32 %%res%d, %%q2, %%q3;" % c)
        L.append("STORE:")
        for c in range(4):
            L.append("    st.global.u32 [%%oaddr%d], %%res%d;" % (c, c))
        L.append("DONE:")
        L.append("    ret;")
        L.append("}")
        return "\n".join(L) + "\n"

    def with_prefetch(self, body):
        """EVAL_PREFETCH=1|2: for every fence-delimited window, prefetch (L1 | L2) the tap lines the NEXT window loads,
        spread over the first half of the current window. The loads themselves stay where they are (right before
        first use, no extra live registers); the prefetch only turns their DRAM / L2 latency into an L1 / L2 hit."""
        level = int(os.environ.get("EVAL_PREFETCH", "0"))
        if not level:
            return body
        fence_lines = ("@%p0 bra DONE;", "bar.sync 0;")
        windows = [[]]
        for ln in body:
            windows[-1].append(ln)
            if ln.strip() in fence_lines:
                windows.append([])
        pat = re.compile(r"\s*mad\.wide\.u32 (%w\d+), %stride, (\d+), (%b_\w+);")

        def taps(w):
            seen, order = set(), []
            for j, ln in enumerate(w):
                if "ld.global.nc.u32" in ln:
                    m = pat.match(w[j - 1])
                    key = (m.group(2), m.group(3))
                    if key not in seen:
                        seen.add(key)
                        order.append(key)
            return order

        out = []
        self.uses_prefetch = True
        for wi, w in enumerate(windows):
            pf = []
            if wi + 1 < len(windows):
                here = set(taps(w)) if level < 3 else set()
                pf = [k for k in taps(windows[wi + 1]) if k not in here]
            if not pf:
                out.extend(w)
                continue
            step = max(1, (len(w) // 2) // len(pf))
            k = 0
            for j, ln in enumerate(w):
                # never between a mad.wide and the load that uses it, or inside a carry chain (.cc / madc pairs)
                prev = w[j - 1] if j else ""
                if k < len(pf) and j >= k * step and ".cc." not in prev and "madc" not in ln and "ld.global.nc" not in ln:
                    col, base = pf[k]
                    out.append("    mad.wide.u32 %%pfa, %%stride, %s, %s;" % (col, base))
                    out.append("    prefetch.global.%s [%%pfa];" % ("L2" if level == 2 else "L1"))
                    k += 1
                out.append(ln)
            for col, base in pf[k:]:
                out.append("    mad.wide.u32 %%pfa, %%stride, %s, %s;" % (col, base))
                out.append("    prefetch.global.%s [%%pfa];" % ("L2" if level == 2 else "L1"))
        return out

    def _mat(self, i):
        m = self.t()
        self.emit("mov.u32 %s, %s;" % (m, self.S.nodes[i][1]))
        return m


LAUNCHER = r"""// GENERATED by tools/gen_eval_check.py - do not edit. Circuit %(name)s: host launcher of the %(nparts)d part kernels
// (gen/eval_check_%(name)s_p*.ptx, assembled by ptxas and embedded in the library as cubins).
//
// Replaces CircuitHal::eval_check for %(name)s (risc0/zkp/src/hal/mod.rs:279-289; CPU spec
// risc0/circuit/rv32im/src/prove/hal/cpu.rs:145-208; reference GPU path rv32im-sys/kernels/cuda/ffi_supra.cu:26-79).
// The constraint polynomial is the sum of %(nparts)d groups of top-level terms; part j adds its group, already divided by
// (3x)^N - 1, into `check` (part 0 stores). All per-proof constants travel in the kernel parameter block
// (constant bank): poly_mix powers, their -11 multiples, globals, mix values and the 4 divisor inverses.
#include <algorithm>
#include <mutex>
#include <string>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../ctx.h"
#include "../tables/circuit_%(name)s.h"
#include "../tables/field_tables.h"

using namespace r0;

extern "C" {
%(externs)s
}

namespace {
struct Consts {
  uint32_t pm[%(npm)d][4];
  uint32_t npm[%(npm)d][4];
  uint32_t global[%(n_global)d];
  uint32_t mix[%(n_mix)d];
  uint32_t inv_y[4];
%(coloff_decl)s};
static_assert(sizeof(Consts) == %(cst_size)d, "constant block layout differs from the generator's");

constexpr int kParts = %(nparts)d;
const char* const kPartNames[kParts] = {%(part_names)s};
cudaKernel_t g_kernels[kParts];
std::once_flag g_once;
std::string g_load_error;

// sum of the per-part partial results of one point tile: check[k][i] = sum_j partial[j][k][i]   (concurrent-tiled mode)
__global__ void k_reduce_partials(uint32_t* check, const uint32_t* partial, int nparts, size_t domain, size_t i0, size_t npts) {
  const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= 4 * npts) return;
  const size_t k = w / npts, i = i0 + w %% npts;
  uint32_t acc = 0;
  for (int j = 0; j < nparts; j++) acc = fp_add(acc, partial[((size_t)j * 4 + k) * domain + i]);
  check[k * domain + i] = acc;
}

void load_kernels() {
  const void* images[kParts] = {%(images)s};
  const char* names[kParts] = {%(names)s};
  for (int j = 0; j < kParts; j++) {
    cudaLibrary_t lib;
    // experiment hook (tools/autotune_eval_check.py): a cubin of the same name in $R0B200_CUBIN_DIR replaces the
    // embedded image, so differently assembled parts can be timed without building a library per variant
    cudaError_t e = cudaErrorFileNotFound;
    if (const char* dir = getenv("R0B200_CUBIN_DIR")) {
      const std::string path = std::string(dir) + "/" + names[j] + ".cubin";
      if (FILE* f = fopen(path.c_str(), "rb")) {
        fclose(f);
        e = cudaLibraryLoadFromFile(&lib, path.c_str(), nullptr, nullptr, 0, nullptr, nullptr, 0);
      }
    }
    if (e == cudaErrorFileNotFound) e = cudaLibraryLoadData(&lib, images[j], nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&g_kernels[j], lib, names[j]);
    if (e != cudaSuccess) {
      g_load_error = std::string("loading ") + names[j] + ": " + cudaGetErrorString(e);
      return;
    }
  }
}
}  // namespace

// check: 4 x domain words out. accum/code/data: evaluated groups (cols x domain), device. global_host: %(n_global)d words,
// mix_host: %(n_mix)d words (host). poly_mix: the drawn FpExt. po2 = log2(cycles); domain = 4 << po2.
void r0_eval_check_%(name)s(Ctx* c, uint32_t* check, const uint32_t* accum, const uint32_t* code, const uint32_t* data,
                         const uint32_t* global_host, const uint32_t* mix_host, const FpExt& poly_mix, uint32_t po2) {
  const size_t domain = size_t(4) << po2;
  R0_CHECK(po2 + 2 <= 27 && domain <= 0x80000000ull, "eval_check: po2 out of range");
  std::call_once(g_once, load_kernels);
  if (!g_load_error.empty()) throw CudaError(g_load_error);
  PhaseScope ph(c, "eval_check", (4.0 * %(ncols)d + 16.0) * (double)domain);
  static thread_local Consts k;
  // poly_mix^POLY_MIX_POWERS[k]: the table is increasing, so walk it with one running power
  {
    FpExt cur = ext_one();
    uint32_t have = 0;
    for (int j = 0; j < %(npm)d; j++) {
      const uint32_t want = %(NAME)s_POLY_MIX_POWERS[j];
      if (want < have) {
        cur = ext_one();
        have = 0;
      }
      if (want - have > 64) {
        cur = ext_mul(cur, ext_pow(poly_mix, want - have));
      } else {
        for (uint32_t s = have; s < want; s++) cur = ext_mul(cur, poly_mix);
      }
      have = want;
      for (int q = 0; q < 4; q++) {
        k.pm[j][q] = cur.c[q];
        k.npm[j][q] = fp_mul(cur.c[q], FP_NBETA);
      }
    }
  }
  memcpy(k.global, global_host, sizeof(k.global));
  memcpy(k.mix, mix_host, sizeof(k.mix));
  // (3 * w^i)^N - 1 with w = ROU_FWD[po2+2], N = 2^po2 depends only on i mod 4: (3^N) * w_4^(i mod 4) - 1
  const uint32_t three_n = fp_pow(FP_THREE, uint64_t(1) << po2);
  uint32_t cur = MONT_ONE;
  for (int r = 0; r < 4; r++) {
    k.inv_y[r] = fp_inv(fp_sub(fp_mul(three_n, cur), MONT_ONE));
    cur = fp_mul(cur, R0_ROU_FWD_MONT[2]);
  }
%(coloff_fill)s  const unsigned threads = %(threads)d;
  // Optional point tiles (EVAL_TILE_LG at generation time; default = whole domain): running all parts over a tile of
  // 2^16 points keeps its tap columns (83 MB for rv32im) in L2 for the later parts and cuts DRAM traffic ~15x, but the
  // 512-block launches it needs run at well under full occupancy: measured 26.1 ms per 2^20 points against 16.2 ms
  // untiled (2^17: 21.0, 2^15: 35.3; gpurun_out/evalcheck_variants12.log), so it is off.
  uint32_t domain32 = (uint32_t)domain;
  // Concurrent-tiled mode (R0B200_EVAL_TILED=<lg of the point tile>, off by default - see DESIGN.md 3.4 for the
  // measurement): all parts of one tile run CONCURRENTLY on a few auxiliary streams, each storing into its own partial
  // buffer (planes of `domain` words, so the kernels need no change), and a small kernel sums the partials of the tile
  // into `check`. The tile's tap columns are then touched by all parts within the same window and stay in L2, while
  // the launches in flight together still fill the GPU - which the sequential tiling below cannot do.
  static const int tiled_lg = getenv("R0B200_EVAL_TILED") ? atoi(getenv("R0B200_EVAL_TILED")) : 0;
  static const int nstreams = getenv("R0B200_EVAL_STREAMS") ? atoi(getenv("R0B200_EVAL_STREAMS")) : 8;
  static const int max_ahead = getenv("R0B200_EVAL_AHEAD") ? atoi(getenv("R0B200_EVAL_AHEAD")) : 1;
  if (tiled_lg >= 10 && domain > (size_t(1) << tiled_lg) && nstreams >= 1 && nstreams <= 32) {
    const size_t tile_pts = size_t(1) << tiled_lg;
    const size_t ntiles = domain / tile_pts;
    while ((int)c->aux_streams.size() < nstreams) {
      cudaStream_t st;
      R0_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      c->aux_streams.push_back(st);
    }
    // events: [0] = start, then per (tile parity, stream) "parts done", then per tile slot "reduced"
    const size_t nev = 1 + 2 * (size_t)nstreams + (size_t)(max_ahead + 1);
    while (c->aux_events.size() < nev) {
      cudaEvent_t ev;
      R0_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      c->aux_events.push_back(ev);
    }
    uint32_t* partial = nullptr;
    R0_CUDA(cudaMallocAsync(&partial, (size_t)kParts * 4 * domain * 4, c->stream));
    R0_CUDA(cudaEventRecord(c->aux_events[0], c->stream));
    for (int s = 0; s < nstreams; s++) R0_CUDA(cudaStreamWaitEvent(c->aux_streams[s], c->aux_events[0], 0));
    const unsigned blocks = (unsigned)((tile_pts + threads - 1) / threads);
    uint32_t first = 1u;
    for (size_t t = 0; t < ntiles; t++) {
      uint32_t base = (uint32_t)(t * tile_pts);
      // bound the drift: the part streams may run at most `max_ahead` tiles ahead of the reduction
      if (t > (size_t)max_ahead) {
        cudaEvent_t red = c->aux_events[1 + 2 * nstreams + (t - 1 - max_ahead) %% (max_ahead + 1)];
        for (int s = 0; s < nstreams; s++) R0_CUDA(cudaStreamWaitEvent(c->aux_streams[s], red, 0));
      }
      for (int j = 0; j < kParts; j++) {
        uint32_t* out = partial + (size_t)j * 4 * domain;
        void* args[] = {&out, &accum, &code, &data, &domain32, &first, &base, &k};
        R0_CUDA(cudaLaunchKernel((const void*)g_kernels[j], dim3(blocks), dim3(threads), args, 0, c->aux_streams[j %% nstreams]));
        count_launch(c);
      }
      for (int s = 0; s < nstreams; s++) {
        cudaEvent_t ev = c->aux_events[1 + (t & 1) * nstreams + s];
        R0_CUDA(cudaEventRecord(ev, c->aux_streams[s]));
        R0_CUDA(cudaStreamWaitEvent(c->stream, ev, 0));
      }
      k_reduce_partials<<<(unsigned)((4 * tile_pts + 255) / 256), 256, 0, c->stream>>>(check, partial, kParts, domain, t * tile_pts, tile_pts);
      count_launch(c);
      R0_CUDA(cudaEventRecord(c->aux_events[1 + 2 * nstreams + t %% (max_ahead + 1)], c->stream));
    }
    R0_CUDA(cudaGetLastError());
    R0_CUDA(cudaFreeAsync(partial, c->stream));
    return;
  }
  const size_t tile = std::min<size_t>(domain, size_t(1) << %(tile_lg)d);
  static const bool profile_parts = getenv("R0B200_PROFILE_PARTS") != nullptr;  // per-part timing is opt-in
  for (size_t i0 = 0; i0 < domain; i0 += tile) {
    const size_t npts = std::min(tile, domain - i0);
    const unsigned blocks = (unsigned)((npts + threads - 1) / threads);
    uint32_t base = (uint32_t)i0;
    for (int j = 0; j < kParts; j++) {
      uint32_t first = j == 0 ? 1u : 0u;
      void* args[] = {&check, &accum, &code, &data, &domain32, &first, &base, &k};
      PhaseScope part(c, profile_parts ? kPartNames[j] : nullptr);
      R0_CUDA(cudaLaunchKernel((const void*)g_kernels[j], dim3(blocks), dim3(threads), args, 0, c->stream));
      count_launch(c);
    }
  }
}
"""


def main():
    name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "rv32im"
    nparts, flatten, threads = 0, True, int(os.environ.get("EVAL_THREADS", "128"))
    for a in sys.argv:
        if a.startswith("--parts="):
            nparts = int(a.split("=")[1])
        if a == "--no-flatten":
            flatten = False
    cfg = CIRCUITS[name]
    nparts = nparts or int(os.environ.get("EVAL_PARTS", "0"))   # 0: as many parts as the unit limit asks for
    ir_path = os.path.join(ROOT, "risc0_b200", "circuits", name + ".ir.json.gz")
    if "--from-ir" in sys.argv or not os.path.exists(cfg["srcs"][0]):
        dag = load_ir(ir_path)
    else:
        fns = ir.parse_functions(cfg["srcs"])
        dag = ir.build_dag(fns, arg_names=cfg["arg_names"])
        save_ir(dag, ir_path)
    npm = 1 + max(k[1] for k in dag.nodes if k[0] == "pm")
    lay = Layout(npm, cfg["n_global"], cfg["n_mix"], max(cfg["cols"].values()))
    uses = [0] * len(dag.nodes)
    for k in dag.nodes:
        if k[0] in "+-*":
            uses[k[1]] += 1
            uses[k[2]] += 1
    parts = partition(dag, uses, nparts)
    gen_dir = os.environ.get("EVAL_OUT_DIR") or os.path.join(ROOT, "risc0_b200", "csrc", "gen")
    os.makedirs(gen_dir, exist_ok=True)
    for f in os.listdir(gen_dir):
        if f.startswith("eval_check_%s" % name):
            os.remove(os.path.join(gen_dir, f))
    externs, images, names = [], [], []
    stats = dict(circuit=name, parts=len(parts), wide_multiplies=0, reductions=0, ptx_instructions=0, tap_loads=0)
    for j, part in enumerate(parts):
        S, outs = lower(dag, part["terms"], lay)
        if flatten:
            S, outs = flatten_sums(S, outs)
        order = schedule_scalars(S, outs)   # only for the statistics printed below; emission is demand-driven
        kname = "eval_check_%s_p%d" % (name, j)
        ptx = Ptx(S, kname, lay).kernel(order, outs, threads)
        with open(os.path.join(gen_dir, kname + ".ptx"), "w") as f:
            f.write(ptx)
        from collections import Counter
        cnt = Counter(S.nodes[i][0] for i in order)
        nprod = sum(len(S.nodes[i][1]) for i in order if S.nodes[i][0] == "d")
        stats["wide_multiplies"] += nprod + cnt.get("*", 0) + 4      # dot terms, single products, output scaling
        stats["reductions"] += cnt.get("d", 0) + cnt.get("*", 0) + 4
        stats["ptx_instructions"] += sum(1 for ln in ptx.splitlines() if ln.startswith("    "))
        stats["tap_loads"] += ptx.count("ld.global.nc.u32")
        print("part %d: %d terms, scalar ops %s, dot products %d" % (j, len(part["terms"]), dict(cnt), nprod))
        externs.append("extern const unsigned char r0_cubin_%s[];" % kname)
        images.append("r0_cubin_%s" % kname)
        names.append('"%s"' % kname)
    params = dict(name=name, NAME=name.upper(), nparts=len(parts), npm=npm, n_global=cfg["n_global"], n_mix=cfg["n_mix"],
                  cst_size=lay.size, threads=threads, ncols=sum(cfg["cols"].values()),
                  tile_lg=int(os.environ.get("EVAL_TILE_LG", "30")), externs="\n".join(externs), images=", ".join(images),
                  names=", ".join(names), part_names=", ".join('"eval_check_p%d"' % j for j in range(len(parts))),
                  coloff_decl=("  uint64_t col_off[%d];  // byte offset of column q inside an evaluated group\n" % lay.ncols) if lay.ncols else "",
                  coloff_fill=("  for (int q = 0; q < %d; q++) k.col_off[q] = (uint64_t)q * domain * 4;\n" % lay.ncols) if lay.ncols else "")
    with open(os.path.join(gen_dir, "eval_check_%s.cu" % name), "w") as f:
        f.write(LAUNCHER % params)
    # per-point operation counts of the emitted kernels (bench.py's INT32 roofline reads them)
    if not os.environ.get("EVAL_OUT_DIR"):
        with open(os.path.join(ROOT, "risc0_b200", "circuits", name + ".stats.json"), "w") as f:
            json.dump(stats, f, indent=1)


if __name__ == "__main__":
    import threading
    threading.stack_size(1 << 29)   # the demand-driven emitter recurses along the DAG depth
    t = threading.Thread(target=main)
    t.start()
    t.join()


# ------------------------------------------------------------------------------------------------ reference evaluator
def evaluate_scalars(S, outs, tap, cst):
    """Pure-python evaluation of a lowered scalar DAG (Montgomery words, exact integer arithmetic) - used by the CPU
    test-suite to check the generator's lowering / flattening against the reference poly_fp without a GPU.
    tap(buf, col, back) -> Montgomery word of that tap at the point under evaluation; cst(offset) -> constant word."""
    val = [None] * len(S.nodes)
    for i, k in enumerate(S.nodes):
        o = k[0]
        if o == "i":
            v = k[1]
        elif o == "k":
            v = cst(k[1])
        elif o == "t":
            v = tap(k[1], k[2], k[3])
        elif o == "+":
            v = (val[k[1]] + val[k[2]]) % P
        elif o == "-":
            v = (val[k[1]] - val[k[2]]) % P
        elif o == "n":
            v = (-val[k[1]]) % P
        elif o == "N":
            v = P - val[k[1]]          # lazy negation: may equal P, only ever used inside a product
        elif o == "*":
            v = val[k[1]] * val[k[2]] * RINV % P
        elif o == "d":
            v = sum(val[a] * val[b] for a, b in k[1]) * RINV % P
        else:
            raise ValueError(k)
        val[i] = v
    return [val[o] for o in outs]
