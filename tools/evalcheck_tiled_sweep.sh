# eval_check: sequential whole-domain parts (default) vs the concurrent-tiled mode, po2 = 19 (2^21 points)
out=gpurun_out/r2_evalcheck_tiled.log
rm -f $out
R0B200_PROFILE_PARTS= python tools/bench_eval_check.py --po2 19 --iters 3 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('default', d['eval_check'], d['ns_per_point'], d['check_crc32'])" >> $out 2>&1
for cfg in "16 8 1" "16 4 1" "16 12 1" "16 8 2" "15 8 1" "17 8 1" "17 12 2" "16 18 1" "16 36 1"; do
  set -- $cfg
  R0B200_EVAL_TILED=$1 R0B200_EVAL_STREAMS=$2 R0B200_EVAL_AHEAD=$3 python tools/bench_eval_check.py --po2 19 --iters 3 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tiled lg=$1 streams=$2 ahead=$3', d['eval_check'], d['ns_per_point'], d['check_crc32'])" >> $out 2>&1
done
cat $out
