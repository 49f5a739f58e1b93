"""GPU box: run the device witness generator on the all-instruction guest (po2 = 14) and, when it fails or differs from
the reference's compiled C++ witgen, print the mismatching cells (cycle, major / minor, column, got / want).

    [R0B200_LIB=risc0_b200/lib/libr0b200_<variant>.so] python tools/dbg_witgen.py

Used to localise the exec-kernel build problem described in DESIGN.md 3.6 (tools/witgen_const_check.sh)."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import witgen_ref as W
from risc0_b200 import B200Hal, preflight as PF
from risc0_b200.hal import WitnessGenerator, Rv32imCircuitHal, R0B200Error
from test_preflight import all_insn_guest
seg = PF.execute(all_insn_guest(), segment_po2=14)[0]
pf = PF.PreflightResults(seg, (11, 12, 13, 14))
want_glob, want_data = W.ref_generate_witness(pf)
hal = B200Hal(0, 'poseidon2')
wg = None
try:
    wg = WitnessGenerator(hal, pf)
    print('witgen ok')
    data = wg.data.view()
except R0B200Error as e:
    print('ERR', e)
    # the buffers were allocated on the instance before the failing call; dig them out of the frame
    import gc
    data = None
    for o in gc.get_objects():
        if isinstance(o, WitnessGenerator):
            data = o.data.view(); break
rows = pf.rows
if data is not None:
    data = data.reshape(-1, rows).copy(); want = np.asarray(want_data).reshape(-1, rows)
    data[data == 0xffffffff] = 0
    bad = np.argwhere(data != want)
    print('mismatching cells', len(bad))
    cyc = sorted(set(int(b[1]) for b in bad))
    print('cycles', cyc[:40])
    for c in cyc[:8]:
        cols = [int(b[0]) for b in bad if b[1] == c]
        cy = pf.cycles[c]
        print('cycle', c, 'major', int(cy['major']), 'minor', int(cy['minor']), 'cols', cols[:40])
        for col in cols[:12]:
            print('   col', col, 'got', hex(int(data[col, c])), 'want', hex(int(want[col, c])))
