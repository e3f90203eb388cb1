cd snark_verifier_axiom_b200/csrc
for mb in 1 8 9 10; do
  touch verify.cu
  make NVCC="/usr/local/cuda/bin/nvcc -DSVK_MSMVAR_MINBLOCKS=$mb" > /dev/null 2>&1
  grep -A2 "Function properties for _Z9k_msm_varILb1" verify.ptxas.log | tail -2
  for i in 1 2; do (cd ../..; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels_one_launch_in_flight']; print($mb, round(d['value']/1e6,3), round(d['e2e']['value']/1e6,3), 'msm_var', round(k['k_msm_var']['ms_per_launch_of_B_batches'],2), 'latB', round(d['config']['launch_latency_ms'],1))"); done
done
(cd ../..; timeout 600 python bench.py --no-cpu-baseline --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default mb=10', round(d['value']/1e6,3))")
