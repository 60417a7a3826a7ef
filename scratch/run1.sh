set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_gputest_b.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.log
ADB_DP_BATCH=32768 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2_bench_b32k.json 2> gpurun_out/r2_bench_b32k.log
ADB_DP_BATCH=131072 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2_bench_b128k.json 2> gpurun_out/r2_bench_b128k.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_launches_c.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/ncu_d.log 2>&1
tail -3 gpurun_out/r2_gputest_b.log
