"""CPU-only check of the eval_check code generator (tools/gen_eval_check.py): the committed circuit IR, lowered to
scalar ops, partitioned and flattened exactly as for PTX emission, is evaluated in exact integer arithmetic at a few
domain points and must agree with the reference's own compiled poly_fp (oracle/_ref) word for word."""
import os
import sys

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_eval_check as G  # noqa: E402

P = O.P
pytestmark = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")


def test_scalar_lowering_matches_reference_poly_fp():
    dag = G.load_ir(os.path.join(ROOT, "risc0_b200", "circuits", "rv32im.ir.json.gz"))
    npm = 1 + max(k[1] for k in dag.nodes if k[0] == "pm")
    assert npm == 458
    lay = G.Layout(npm, 90, 36)
    uses = [0] * len(dag.nodes)
    for k in dag.nodes:
        if k[0] in "+-*":
            uses[k[1]] += 1
            uses[k[2]] += 1
    parts = G.partition(dag, uses, 8)
    po2 = 5
    n, domain = 1 << po2, 4 << po2
    rng = np.random.default_rng(11)
    accum, data = O.rand_elems(rng, 103 * domain), O.rand_elems(rng, 211 * domain)
    mix, out, poly_mix = O.rand_elems(rng, 36), O.rand_elems(rng, 90), O.rand_ext(rng)
    want = O.rv32im_eval_check(accum, data, mix, out, poly_mix, po2).reshape(4, domain)

    # constant block exactly as the launcher fills it
    L = O.lib()
    tables = open(os.path.join(ROOT, "risc0_b200", "csrc", "tables", "circuit_rv32im.h")).read()
    pows = [int(x) for x in tables.split("RV32IM_POLY_MIX_POWERS[458] = {")[1].split("}")[0].replace("u", "").split(",") if x.strip()]
    pm = np.zeros((npm, 4), dtype=np.uint32)
    tmp = np.zeros(4, dtype=np.uint32)
    for j, e in enumerate(pows):
        L.orc_fpext_pow(O.ptr(tmp), O.ptr(O.u32(poly_mix)), O._u64(e))
        pm[j] = tmp
    nbeta = int(O.encode(P - 11))
    consts = {}
    for j in range(npm):
        for c in range(4):
            consts[lay.pm + 16 * j + 4 * c] = int(pm[j, c])
            consts[lay.npm + 16 * j + 4 * c] = int(L.orc_fp_mul(int(pm[j, c]), nbeta))
    for i in range(90):
        consts[lay.glob + 4 * i] = int(out[i])
    for i in range(36):
        consts[lay.mix + 4 * i] = int(mix[i])
    three_n = L.orc_fp_pow(int(O.encode(3)), O._u64(n))
    w4 = L.orc_rou_fwd(2)
    inv_y, cur = [], int(O.encode(1))
    for _ in range(4):
        inv_y.append(L.orc_fp_inv(L.orc_fp_sub(L.orc_fp_mul(three_n, cur), int(O.encode(1)))))
        cur = L.orc_fp_mul(cur, w4)

    lowered = []
    for part in parts:
        S, outs = G.lower(dag, part["terms"], lay)
        S, outs = G.flatten_sums(S, outs)
        lowered.append((S, outs))
    bufs = {"accum": accum, "data": data}
    for i in (0, 1, 2, 3, 77, domain - 1):
        def tap(buf, col, back, i=i):
            return int(bufs[buf][col * domain + ((i - 4 * back) & (domain - 1))])
        tot = [0, 0, 0, 0]
        for S, outs in lowered:
            v = G.evaluate_scalars(S, outs, tap, lambda off: consts[off])
            tot = [(a + b) % P for a, b in zip(tot, v)]
        got = [L.orc_fp_mul(int(t), int(inv_y[i & 3])) for t in tot]
        assert got == [int(x) for x in want[:, i]], i
