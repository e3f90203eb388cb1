for cfg in "10 2 8 0" "10 2 4 0" "10 2 4 1" "5 4 4 0" "5 4 4 1" "4 5 4 1" "2 10 4 1" "20 1 4 0"; do
  set -- $cfg
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batches-per-launch $1 --inflight $2 --group-size $3 --stream-priorities $4 > gpurun_out/r2f_B$1_S$2_m$3_p$4.json 2> gpurun_out/r2f_B$1_S$2_m$3_p$4.err
done
SVK_TAPE_COOP_MAX=8192 SVK_MSM_LATENCY_THREADS_MAX=50000 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batches-per-launch 5 --inflight 4 --group-size 4 --stream-priorities 1 > gpurun_out/r2f_B5_S4_m4_p1_thr.json 2> gpurun_out/r2f_thr.err
