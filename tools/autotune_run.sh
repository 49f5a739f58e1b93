# on the GPU box: per-part eval_check times of every variant directory -> gpurun_out/r2_autotune.log
out=gpurun_out/${AT_LOG:-r2_autotune.log}
rm -f $out
python tools/bench_eval_check.py --po2 ${PO2:-20} --iters 3 --tile-data >> $out 2>&1
for d in risc0_b200/lib/cubins_at/*/; do
  d=${d%/}
  R0B200_CUBIN_DIR=$PWD/$d python tools/bench_eval_check.py --po2 ${PO2:-20} --iters 3 --tile-data | sed "s|\"lib\": \"default\"|\"lib\": \"$(basename $d)\"|" >> $out 2>&1
done
cut -c1-100 $out | tail -5
