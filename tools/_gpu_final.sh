(timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_r1b.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_r1b.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; echo "bench rc=$?"
cat gpurun_out/bench_r1b.json | cut -c1-400
timeout 900 python tools/microbench.py --json gpurun_out/microbench_r1b.json > gpurun_out/microbench_r1b.log 2>&1; echo "microbench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench_r1b.log 2>&1; echo "ncu rc=$?"
