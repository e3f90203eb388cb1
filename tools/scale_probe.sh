# N = 1, 2, 4, 8 back to back on ONE box (as the driver's scaling run does): bench lines -> gpurun_out/r2_scale_nN.json
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2_scale_n1.json 2> gpurun_out/r2_scale_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) bench.py --gpus $n --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2_scale_n$n.json 2> gpurun_out/r2_scale_n$n.err
done
python - <<'PY'
import json
base = None
for n in (1, 2, 4, 8):
    d = json.loads(open(f"gpurun_out/r2_scale_n{n}.json").read().strip().splitlines()[-1])
    base = base or d["value"]
    print(n, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 3), "eff", round(d["value"] / n / base, 3), d["config"]["single_batch_latency_ms"])
PY
