# usage (on the GPU box): bash tools/p2modes.sh [modes...]  -> gpurun_out/r2_poseidon2_modes2.log
out=gpurun_out/r2_poseidon2_modes2.log
rm -f $out
for m in "$@"; do
  echo -n "mode $m: " >> $out
  R0B200_P2_MODE=$m python tools/bench_hash.py --lg 22 --cols 64 --iters 5 >> $out 2>&1
done
cat $out
