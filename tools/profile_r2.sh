# Round-2 evidence pass (one gpurun call): tests, plain runs, then the same commands under ncu; ncu reports are summarised here
# (tools/ncu_to_json.py) because the raw reports exceed what travels back.
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r2_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; tail -2 gpurun_out/r2_smoke.txt
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_k20.json 2> gpurun_out/r2_bench_n1_k20.err
python bench.py > gpurun_out/r2_bench_n1_default.json 2> gpurun_out/r2_bench_n1_default.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null
python tools/lat_probe.py --group-size 4 > gpurun_out/r2_latency_probe.json 2>/dev/null
python tools/lat_probe.py --group-size 4 --batch 1 >> gpurun_out/r2_latency_probe.json 2>/dev/null
python tools/bench_configs.py --only msm --max-log-n 26 > gpurun_out/r2_msm_sweep.jsonl 2>/dev/null
python tools/bench_configs.py --only decide >> gpurun_out/r2_msm_sweep.jsonl 2>/dev/null
python tools/timeline_probe.py --plan 10,10 --reps 5 > gpurun_out/r2_timeline_2x10.txt 2>/dev/null
$B > gpurun_out/plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_ncu_launches.csv $B > gpurun_out/ncu1.log 2>&1
$B > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_tape|k_msm_var|k_msm_sum|k_decompress|k_group_var|k_fold_sponge" -s 60 -c 8 -o gpurun_out/r2_full $B > gpurun_out/ncu2.log 2>&1
M="python tools/bench_configs.py --only msm --min-log-n 20 --max-log-n 20"
$M > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_msm_buckets|k_msm_reduce|k_msm_combine|k_msm_prepare|k_msm_scatter" -s 15 -c 5 -o gpurun_out/r2_msm $M > gpurun_out/ncu3.log 2>&1
python tools/bench_configs.py --only decide > gpurun_out/plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_decide" -c 1 -o gpurun_out/r2_decide python tools/bench_configs.py --only decide > gpurun_out/ncu4.log 2>&1
P="python tools/lat_probe.py --iters 1"
$P > gpurun_out/plain5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_decide_coop|k_tape_coop|k_fold_sponge_dbl|k_fold_add|k_msm_var" -c 12 -o gpurun_out/r2_lat $P > gpurun_out/ncu5.log 2>&1
python tools/ncu_to_json.py gpurun_out/r2_full.ncu-rep "$B under ncu --set full --clock-control none -k regex:k_tape|k_msm_var|k_msm_sum|k_decompress|k_group_var|k_fold_sponge -s 60 -c 8 (launches of 10 batches = 40960 proofs)" > gpurun_out/r2_ncu_full_summary.json
python tools/ncu_to_json.py gpurun_out/r2_msm.ncu-rep "$M under ncu --set full -k regex:k_msm_* -s 15 -c 5 (the 2^20-point MSM)" > gpurun_out/r2_ncu_msm_summary.json
python tools/ncu_to_json.py gpurun_out/r2_decide.ncu-rep "python tools/bench_configs.py --only decide under ncu --set full -k regex:k_decide -c 1 (65536 accumulators)" > gpurun_out/r2_ncu_decide_summary.json
python tools/ncu_to_json.py gpurun_out/r2_lat.ncu-rep "$P under ncu --set full --clock-control none -k regex:k_decide_coop|k_tape_coop|k_fold_sponge_dbl|k_fold_add|k_msm_var -c 12 (one 4096-proof batch, nothing else in flight)" > gpurun_out/r2_ncu_latency_kernels_summary.json
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out
echo done
