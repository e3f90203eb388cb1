#!/bin/bash
# launch plans for K = 20 batches: sizes of the concurrent calls (tools/timeline_probe.py --plan), with the k_tape schedule threshold
mkdir -p gpurun_out
out=gpurun_out/r2_plan_sweep2.txt
: > $out
for coop in 32768 8192; do
for plan in 10,10 7,7,6 5,5,5,5 4,4,4,4,4; do
  echo "SVK_TAPE_COOP_MAX=$coop" >> $out
  SVK_TAPE_COOP_MAX=$coop python tools/timeline_probe.py --plan $plan --reps 7 --no-timeline 2>&1 | grep "^# plan" >> $out
done
done
cat $out
