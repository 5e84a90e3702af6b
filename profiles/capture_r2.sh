#!/bin/bash
# ncu captures of the round-2 build (run under gpurun, 1 GPU).  Every ncu command is preceded by the identical plain
# command (B200_PROFILING.md); outputs land in gpurun_out/ and are summarised into profiles/ by summarize_ncu.py.
# Kept small on purpose: `--set full` replays every kernel ~40 times and a report with sources is 1-2 MB per launch
# (the first version of this script captured 90 launches per group, ran into its time limit and overflowed the 64 MiB
# that gpurun copies back).
set -u
O=gpurun_out
K="python profiles/bench_kernels.py"
run() {  # name, ncu args..., -- command...
  local name=$1; shift
  local nargs=(); while [ "$1" != "--" ]; do nargs+=("$1"); shift; done; shift
  "$@" > $O/${name}_plain.log 2>&1 && timeout 400 ncu "${nargs[@]}" "$@" > $O/${name}_ncu.log 2>&1
  echo "$name rc=$?"
}
FULL="--set full --clock-control none"
# (1) the dominant kernel as shipped: CTA-pair halo conv, 128->128 with every epilogue variant (plain ... gn-bwd fusion)
run r2_epi $FULL --import-source on -k regex:conv_halo_pair -c 7 -o $O/r2_epi -f -- $K epi --first 1 --iters 1 --warmup 0
# (2) every conv shape once (halo kernel down to 32x32, generic implicit GEMM below / 1x1, split-K instances)
run r2_convs $FULL -k regex:"conv_gemm_kernel|conv_halo" -c 24 -o $O/r2_convs -f -- $K conv --iters 1 --warmup 0
# (3) weight-gradient kernels, every shape once
run r2_wgrad $FULL -k regex:wgrad -c 24 -o $O/r2_wgrad -f -- $K wgrad --iters 1 --warmup 0
# (4) GroupNorm streaming kernels at 128x128
run r2_gn $FULL -k regex:"gn_" -c 16 -o $O/r2_gn -f -- $K gn --first 2 --iters 1 --warmup 0
# (5) launch list of the eager training step (shares)
B="python bench.py --steps 1 --warmup 3 --no-graph --no-sampling --no-lora --no-cpu-baseline"
$B > $O/r2_launch_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4400 -c 1300 --csv --log-file $O/r2_launches.csv $B > $O/r2_launch_ncu.log 2>&1
echo "launches rc=$?"
ls -la $O/ | head -40; du -sh $O
