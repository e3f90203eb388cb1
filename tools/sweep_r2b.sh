set -x
for cfg in "20 1" "10 2" "5 4" "4 5" "2 10" "1 16"; do
  set -- $cfg
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batches-per-launch $1 --inflight $2 > gpurun_out/r2b_B$1.json 2> gpurun_out/r2b_B$1.err
done
SVK_TAPE_COOP_MAX=8192 SVK_MSM_LATENCY_THREADS_MAX=50000 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batches-per-launch 5 --inflight 4 > gpurun_out/r2b_B5_thr.json 2> gpurun_out/r2b_B5_thr.err
SVK_TAPE_COOP_MAX=8192 SVK_MSM_LATENCY_THREADS_MAX=50000 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batches-per-launch 2 --inflight 10 > gpurun_out/r2b_B2_thr.json 2> gpurun_out/r2b_B2_thr.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2b_default.json 2> gpurun_out/r2b_default.err
