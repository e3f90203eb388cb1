# batched k_decide: registers per thread vs warps per sub-partition (rebuilds decide.o on the box)
cd snark_verifier_axiom_b200/csrc
for mb in 4 6 8; do
  touch decide.cu
  make NVCC="/usr/local/cuda/bin/nvcc -DSVK_DECIDE_MINBLOCKS=$mb" > /dev/null 2>&1
  grep -A2 "Function properties for _Z8k_decide" decide.ptxas.log | tail -2
  (cd ../..; python tools/bench_configs.py --only decide 2>/dev/null | tail -1)
done
touch decide.cu; make > /dev/null 2>&1
