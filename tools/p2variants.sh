# on the GPU box: hash_rows occupancy / ILP variants (poseidon2.cu, R0B200_P2_VARIANT) -> gpurun_out/r2_poseidon2_variants.log
out=gpurun_out/r2_poseidon2_variants.log
rm -f $out
for v in ${VARIANTS:-0 1 2 3 4}; do
  echo -n "variant $v: " >> $out
  R0B200_P2_VARIANT=$v python tools/bench_hash.py --lg 22 --cols 64 --iters 5 >> $out 2>&1
done
cat $out
