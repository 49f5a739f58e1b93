"""Link risc0_b200/lib/libr0b200_<name>.so from the default build's objects with ONE replaced witgen step object
(experiment builds of the exec kernel: other launch bounds / ptxas levels), e.g.

    nvcc ... -DWG_MIN_BLOCKS=10 -c risc0_b200/csrc/witgen_step_exec.cu -o /tmp/wg10/witgen_step_exec.o
    python tools/link_witgen_variant.py wg10 /tmp/wg10/witgen_step_exec.o
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from risc0_b200 import build as B  # noqa: E402

name, repl = sys.argv[1], sys.argv[2]
objs = []
for src in B.sources():
    o = os.path.join(B.OBJ, os.path.relpath(src, B.CSRC).replace(os.sep, "_")[:-3] + ".o")
    objs.append(repl if os.path.basename(o) == os.path.basename(repl) else o)
objs.append(os.path.join(B.OBJ, "embed_cubins.o"))
lib = os.path.join(B.LIBDIR, "libr0b200_%s.so" % name)
subprocess.check_call([B.NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs)
print(lib)
