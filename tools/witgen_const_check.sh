# GPU box: the witgen suite + phase times for the default library and the exec-kernel variants named on the command line
# (libr0b200_<v>.so, tools/link_witgen_variant.py); tools/dbg_witgen.py prints the mismatching cells of a failing variant
out=gpurun_out/${TAG:-r2g}_witgen_variants.log
rm -f $out
echo "== default lib: tests/test_gpu_witgen.py" >> $out
timeout 600 python -m pytest tests/test_gpu_witgen.py -x -q 2>&1 | tail -3 >> $out
timeout 300 python tools/bench_witgen.py --po2 20 >> $out 2>&1
for v in "$@"; do
  V=$PWD/risc0_b200/lib/libr0b200_$v.so
  echo "== $v: all-instruction guest, po2=14 (tools/dbg_witgen.py)" >> $out
  R0B200_LIB=$V timeout 300 python tools/dbg_witgen.py 2>&1 | head -12 >> $out
  echo "== $v: tests/test_gpu_witgen.py" >> $out
  R0B200_LIB=$V timeout 600 python -m pytest tests/test_gpu_witgen.py -q 2>&1 | tail -6 >> $out
  R0B200_LIB=$V timeout 300 python tools/bench_witgen.py --po2 20 >> $out 2>&1
done
cat $out
