# on the GPU box (1 GPU): the round's validation + profile set, every piece into gpurun_out/<tag>_*
tag=${1:-r2b}
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gputest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_gputest.log; tail -2 gpurun_out/${tag}_gputest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; cut -c1-300 gpurun_out/${tag}_bench_n1.json
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; cut -c1-300 gpurun_out/${tag}_bench_reference.json
python tools/profile_target.py > gpurun_out/${tag}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_target.py > gpurun_out/${tag}_ncu_launches.log 2>&1
[ -n "$SKIP_SECTIONS" ] || ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section Occupancy \
    --section LaunchStats --section WarpStateStats --section SchedulerStats --section InstructionStats --clock-control none \
    -k "regex:eval_check_rv32im_p|p2_hash_rows_kernel|p2_hash_fold_kernel|k_step_exec|k_step_accum|ntt_strided|ntt_fwd_contig|ntt_inv_contig|bit_reverse_tma" -c 90 \
    -o /tmp/${tag}_sections python tools/profile_target.py > gpurun_out/${tag}_ncu_sections.log 2>&1
[ -n "$SKIP_SECTIONS" ] || ncu -i /tmp/${tag}_sections.ncu-rep --page raw --csv > gpurun_out/${tag}_sections_raw.csv 2> gpurun_out/${tag}_sections_raw.err
ls -la gpurun_out/${tag}_*
