#!/bin/bash
# ncu captures of the round-2 build (run under gpurun, 1 GPU).  Every ncu command is preceded by the identical plain
# command (B200_PROFILING.md); outputs land in gpurun_out/ and are summarised into profiles/ by summarize_ncu.py.
set -u
O=gpurun_out
K="python profiles/bench_kernels.py"
run() {  # name, ncu args..., -- command...
  local name=$1; shift
  local nargs=(); while [ "$1" != "--" ]; do nargs+=("$1"); shift; done; shift
  "$@" > $O/${name}_plain.log 2>&1 && ncu "${nargs[@]}" "$@" > $O/${name}_ncu.log 2>&1
  echo "$name rc=$?"
}
# (1) the dominant kernel as shipped: CTA-pair halo conv, 128->128 and 256->128 3x3 @128^2, batch 64
run r2_pair --set full --clock-control none --import-source on -k regex:conv_halo_pair -s 3 -c 2 -o $O/r2_pair -f -- $K conv --first 2 --iters 1
# (2) generic implicit-GEMM instances and the halo kernel at low resolution: every conv shape once
run r2_convs --set full --clock-control none --import-source on -k regex:"conv_gemm_kernel|conv_halo" -c 90 -o $O/r2_convs -f -- $K conv --iters 1
# (3) weight-gradient kernels, every shape once
run r2_wgrad --set full --clock-control none --import-source on -k regex:wgrad -c 90 -o $O/r2_wgrad -f -- $K wgrad --iters 1
# (4) epilogue variants of the pair kernel (fusion costs)
run r2_epi --set full --clock-control none --import-source on -k regex:conv_halo_pair -c 40 -o $O/r2_epi -f -- $K epi --first 1 --iters 1
# (5) GroupNorm streaming kernels
run r2_gn --set full --clock-control none --import-source on -k regex:"gn_" -c 60 -o $O/r2_gn -f -- $K gn --first 2 --iters 1
# (6) launch list of the eager training step (shares)
B="python bench.py --steps 2 --warmup 3 --no-graph --no-sampling --no-lora --no-cpu-baseline"
$B > $O/r2_launch_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 4400 -c 1300 --csv --log-file $O/r2_launches.csv $B > $O/r2_launch_ncu.log 2>&1
echo "launches rc=$?"
ls -la $O/*.ncu-rep
