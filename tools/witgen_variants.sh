# on the GPU box: witgen / accum phase times for the default library and the variants named on the command line
out=gpurun_out/r2_witgen_variants.log
rm -f $out
python tools/bench_witgen.py --po2 20 >> $out 2>&1
for v in "$@"; do
  R0B200_LIB=$PWD/risc0_b200/lib/libr0b200_$v.so python tools/bench_witgen.py --po2 20 >> $out 2>&1
done
cat $out
