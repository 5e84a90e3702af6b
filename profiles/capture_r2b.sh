#!/bin/bash
# Launch lists (ncu --metrics gpu__time_duration.sum, one pass, cold caches, serialised) of one eager pass of the three
# workloads bench.py reports: sampling step (batch 32, 128^2), LoRA step (celebahq, batch 8, 256^2), training step
# (batch 64, 128^2).  Every ncu command is preceded by the identical plain command (B200_PROFILING.md).
set -u
O=gpurun_out
for mode in sampling lora train; do
  C="python profiles/step_eager.py $mode"
  $C > $O/r2b_${mode}_plain.log 2>&1 && timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum \
      --clock-control none --csv --log-file $O/r2b_launches_${mode}.csv $C > $O/r2b_${mode}_ncu.log 2>&1
  echo "$mode rc=$?"; cat $O/r2b_${mode}_plain.log | tail -1
done
ls -la $O | head -30
