"""Self-test of tools/ptx_interp.py (the CPU interpreter behind test_eval_check_generator's PTX check): carry chains,
wide multiplies, predicated branches, parameter block reads and global memory, on a hand-written kernel."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ptx_interp import Kernel, Memory  # noqa: E402

PTX = """
.version 8.6
.target sm_100a
.address_size 64
.visible .entry k(
    .param .u64 p_out, .param .u32 p_n,
    .param .align 16 .b8 p_cst[16])
{
    .reg .pred %p<2>;
    .reg .u32 %a, %b, %lo, %hi, %i, %n, %x;
    .reg .u64 %out, %w, %addr;
    ld.param.u64 %out, [p_out];
    cvta.to.global.u64 %out, %out;
    ld.param.u32 %n, [p_n];
    mov.u32 %i, %tid.x;
    setp.ge.u32 %p1, %i, %n;
    @%p1 bra DONE;
    ld.param.u32 %a, [p_cst+4];
    ld.param.u32 %b, [p_cst+8];
    mul.wide.u32 %w, %a, %b;          // (lo, hi) = a * b
    mov.b64 {%lo, %hi}, %w;
    mad.lo.cc.u32 %lo, %a, %b, %lo;   // += a * b with the carry from the low word
    madc.hi.u32 %hi, %a, %b, %hi;
    add.u32 %x, %hi, -1;
    min.u32 %x, %x, %hi;
    mad.wide.u32 %addr, %i, 8, %out;
    st.global.u32 [%addr], %lo;
    st.global.u32 [%addr+4], %x;
DONE:
    ret;
}
"""


def test_interpreter_semantics():
    k = Kernel(PTX)
    assert [n for n, _ in k.params] == ["p_out", "p_n", "p_cst"] and k.params[2][1] == ("b8", 16)
    a, b = 0xFFFFFFF1, 0xF0000003
    cst = (0).to_bytes(4, "little") + a.to_bytes(4, "little") + b.to_bytes(4, "little") + (0).to_bytes(4, "little")
    out = np.zeros(8, dtype=np.uint32)
    mem = Memory({0x1000: out})
    for tid in range(5):
        k.run(dict(p_out=0x1000, p_n=3, p_cst=cst), mem, tid=tid)
    want = 2 * a * b
    lo, hi = want & 0xFFFFFFFF, (want >> 32) & 0xFFFFFFFF
    for tid in range(3):
        assert int(out[2 * tid]) == lo and int(out[2 * tid + 1]) == min(hi, (hi - 1) & 0xFFFFFFFF)
    assert not out[6:].any()      # threads 3 and 4 took the branch


def test_interpreter_rejects_unknown_ops_and_wild_addresses():
    bad = PTX.replace("min.u32 %x, %x, %hi;", "popc.b32 %x, %hi;")
    with pytest.raises(NotImplementedError):
        Kernel(bad).run(dict(p_out=0x1000, p_n=1, p_cst=bytes(16)), Memory({0x1000: np.zeros(8, dtype=np.uint32)}))
    with pytest.raises(IndexError):
        Kernel(PTX).run(dict(p_out=0x9000, p_n=1, p_cst=bytes(16)), Memory({0x1000: np.zeros(8, dtype=np.uint32)}))
