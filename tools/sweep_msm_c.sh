for c in 12 13 14 15 16; do
  SVK_MSM_C=$c python tools/bench_configs.py --only msm --max-log-n 22 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln)
    print('c=$c', d['log_n'], round(d['ms'], 3), {k: v for k, v in d['kernels_ms'].items() if v > 0.05})
"
done
