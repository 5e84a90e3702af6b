#!/bin/bash
# Final captures of round 2 (1 GPU, under gpurun): (1) `ncu --set full` of every epilogue variant of the shipped CTA-pair
# halo conv (one instantiation per variant), raw metrics exported as CSV on the box; (2) launch lists of one eager
# training step, one sampling step and one LoRA step.  Every ncu command follows the identical plain command.
set -u
O=gpurun_out
K="python profiles/bench_kernels.py"
$K epi --first 1 --iters 1 --warmup 0 > $O/r2c_epi_plain.log 2>&1 && timeout 400 ncu --set full --clock-control none \
    -k regex:conv_halo_pair -c 7 -o $O/r2c_epi -f $K epi --first 1 --iters 1 --warmup 0 > $O/r2c_epi_ncu.log 2>&1
echo "epi rc=$?"
[ -f $O/r2c_epi.ncu-rep ] && ncu -i $O/r2c_epi.ncu-rep --page raw --csv > $O/r2c_epi_raw.csv 2>/dev/null && rm -f $O/r2c_epi.ncu-rep
for mode in train sampling lora; do
  C="python profiles/step_eager.py $mode"
  $C > $O/r2c_${mode}_plain.log 2>&1 && timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum \
      --clock-control none --csv --log-file $O/r2c_launches_${mode}.csv $C > $O/r2c_${mode}_ncu.log 2>&1
  echo "$mode rc=$?"; tail -1 $O/r2c_${mode}_plain.log
done
ls -la $O | head -30
