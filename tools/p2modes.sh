for m in 32 0 8 40 64 72 96 104 112 120 100 108 124 127 63 98 106 122 116; do
  echo -n "mode $m: " >> gpurun_out/r2_poseidon2_modes.log
  R0B200_P2_MODE=$m python tools/bench_hash.py --lg 22 --cols 64 --iters 5 >> gpurun_out/r2_poseidon2_modes.log 2>&1
done
cat gpurun_out/r2_poseidon2_modes.log
