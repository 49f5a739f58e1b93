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


def _setup(name="rv32im", po2=5):
    """random evaluated groups + the constant block exactly as the launcher fills it, and the reference's answer"""
    cfg = G.CIRCUITS[name]
    dag = G.load_ir(os.path.join(ROOT, "risc0_b200", "circuits", name + ".ir.json.gz"))
    npm = 1 + max(k[1] for k in dag.nodes if k[0] == "pm")
    assert npm == {"rv32im": 458, "recursion": 158}[name]
    ng, nm = cfg["n_global"], cfg["n_mix"]
    lay = G.Layout(npm, ng, nm, max(cfg["cols"].values()))   # the last argument only matters for EVAL_ADDR_TABLE
    n, domain = 1 << po2, 4 << po2
    rng = np.random.default_rng(11)
    accum, data = O.rand_elems(rng, cfg["cols"]["accum"] * domain), O.rand_elems(rng, cfg["cols"]["data"] * domain)
    mix, out, poly_mix = O.rand_elems(rng, nm), O.rand_elems(rng, ng), O.rand_ext(rng)
    if name == "rv32im":
        code = np.zeros(domain, dtype=np.uint32)    # the reference's rv32im code column is all zero
        want = O.rv32im_eval_check(accum, data, mix, out, poly_mix, po2).reshape(4, domain)
    else:
        code = O.rand_elems(rng, cfg["cols"]["code"] * domain)
        want = O.recursion_eval_check(code, data, accum, mix, out, poly_mix, po2).reshape(4, domain)
    L = O.lib()
    tables = open(os.path.join(ROOT, "risc0_b200", "csrc", "tables", "circuit_%s.h" % name)).read()
    pows = [int(x) for x in tables.split("%s_POLY_MIX_POWERS[%d] = {" % (name.upper(), npm))[1].split("}")[0].replace("u", "").split(",") if x.strip()]
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
    for i in range(ng):
        consts[lay.glob + 4 * i] = int(out[i])
    for i in range(nm):
        consts[lay.mix + 4 * i] = int(mix[i])
    three_n = L.orc_fp_pow(int(O.encode(3)), O._u64(n))
    w4 = L.orc_rou_fwd(2)
    inv_y, cur = [], int(O.encode(1))
    for _ in range(4):
        inv_y.append(L.orc_fp_inv(L.orc_fp_sub(L.orc_fp_mul(three_n, cur), int(O.encode(1)))))
        cur = L.orc_fp_mul(cur, w4)
    return dict(dag=dag, lay=lay, domain=domain, accum=accum, data=data, code=code, want=want, consts=consts, inv_y=inv_y)


@pytest.mark.parametrize("name", ["rv32im", "recursion"])
def test_emitted_ptx_matches_reference_poly_fp(name):
    """the PTX text that ptxas assembles (csrc/gen/eval_check_<circuit>_p*.ptx), run one thread at a time by
    tools/ptx_interp.py, gives the reference's check polynomial word for word"""
    import glob
    import subprocess
    from ptx_interp import Kernel, Memory
    if name == "recursion" and not O.have_ref_recursion():
        pytest.skip("oracle/_ref recursion poly_fp not built")
    gen_dir = os.path.join(ROOT, "risc0_b200", "csrc", "gen")
    pattern = os.path.join(gen_dir, "eval_check_%s_p*.ptx" % name)
    files = glob.glob(pattern)
    if not files:
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_eval_check.py"), name, "--from-ir"],
                              stdout=subprocess.DEVNULL)
        files = glob.glob(pattern)
    files.sort(key=lambda f: int(f.rsplit("_p", 1)[1][:-4]))
    s = _setup(name)
    lay, domain = s["lay"], s["domain"]
    cst = bytearray(lay.size)
    for off, v in s["consts"].items():
        cst[off:off + 4] = int(v).to_bytes(4, "little")
    for r in range(4):
        cst[lay.inv_y + 4 * r:lay.inv_y + 4 * r + 4] = int(s["inv_y"][r]).to_bytes(4, "little")
    for q in range(lay.ncols):   # EVAL_ADDR_TABLE variant: byte offsets of the columns
        cst[lay.coloff + 8 * q:lay.coloff + 8 * q + 8] = (q * domain * 4).to_bytes(8, "little")
    check = np.full(4 * domain, 0xDEADBEEF, dtype=np.uint32)   # part 0 must overwrite, not accumulate
    bases = {"p_check": 1 << 40, "p_accum": 2 << 40, "p_code": 3 << 40, "p_data": 4 << 40}
    mem = Memory({bases["p_check"]: check, bases["p_accum"]: s["accum"].copy(), bases["p_code"]: s["code"].copy(),
                  bases["p_data"]: s["data"].copy()})
    points = (0, 3, 77, domain - 1)
    for j, f in enumerate(files):
        k = Kernel(open(f).read())
        assert [nm for nm, _ in k.params] == ["p_check", "p_accum", "p_code", "p_data", "p_domain", "p_first", "p_i0", "p_cst"]
        assert k.params[-1][1] == ("b8", lay.size)
        for i in points:
            params = dict(bases, p_domain=domain, p_first=1 if j == 0 else 0, p_i0=0, p_cst=bytes(cst))
            k.run(params, mem, tid=i, ctaid=0, ntid=128)
    for i in points:
        assert [int(check[c * domain + i]) for c in range(4)] == [int(x) for x in s["want"][:, i]], i
    # a thread beyond the domain leaves memory alone
    before = check.copy()
    Kernel(open(files[0]).read()).run(dict(bases, p_domain=domain, p_first=1, p_i0=0, p_cst=bytes(cst)), mem, tid=5, ctaid=1, ntid=128)
    assert np.array_equal(check, before)


def test_scalar_lowering_matches_reference_poly_fp():
    s = _setup()
    dag, lay, domain, accum, data, want, consts, inv_y = (s[k] for k in ("dag", "lay", "domain", "accum", "data", "want", "consts", "inv_y"))
    L = O.lib()
    uses = [0] * len(dag.nodes)
    for k in dag.nodes:
        if k[0] in "+-*":
            uses[k[1]] += 1
            uses[k[2]] += 1
    parts = G.partition(dag, uses, 8)

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
