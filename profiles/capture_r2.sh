#!/bin/bash
# ncu captures of the round-2 build (run under gpurun, 1 GPU).  Every ncu command is preceded by the identical plain
# command (B200_PROFILING.md).  `--set full` reports are ~1.3 MB per launch and gpurun copies back at most 64 MiB, so
# every report is exported to its raw-metrics CSV ON THE BOX and deleted; only the small epilogue report (with sources,
# for the per-instruction stall page) travels.  Summaries: profiles/summarize_ncu.py.
set -u
O=gpurun_out
K="python profiles/bench_kernels.py"
run() {  # name, keep-report(0/1), ncu args..., -- command...
  local name=$1; shift
  local keep=$1; shift
  local nargs=(); while [ "$1" != "--" ]; do nargs+=("$1"); shift; done; shift
  "$@" > $O/${name}_plain.log 2>&1 && timeout 400 ncu "${nargs[@]}" -o $O/$name -f "$@" > $O/${name}_ncu.log 2>&1
  echo "$name rc=$?"
  if [ -f $O/$name.ncu-rep ]; then
    ncu -i $O/$name.ncu-rep --page raw --csv > $O/${name}_raw.csv 2>/dev/null
    [ "$keep" = "1" ] || rm -f $O/$name.ncu-rep
  fi
}
FULL="--set full --clock-control none"
# (1) the dominant kernel as shipped: CTA-pair halo conv 128->128 @128^2, batch 64, every epilogue variant (raw metrics)
run r2_epi 0 $FULL -k regex:conv_halo_pair -c 7 -- $K epi --first 1 --iters 1 --warmup 0
# (1b) sources + per-instruction stalls for three of them: plain, bias+res+csum (launch 6), gn-bwd fusion (launch 7)
run r2_epi_src 1 $FULL --import-source on -k regex:conv_halo_pair -s 5 -c 2 -- $K epi --first 1 --iters 1 --warmup 0
# (2) every conv shape once (halo kernel down to 32x32, generic implicit GEMM below / 1x1, split-K instances)
run r2_convs 0 $FULL -k regex:"conv_gemm_kernel|conv_halo" -c 24 -- $K conv --iters 1 --warmup 0
# (3) weight-gradient kernels, every shape once
run r2_wgrad 0 $FULL -k regex:wgrad -c 24 -- $K wgrad --iters 1 --warmup 0
# (4) GroupNorm streaming kernels at 128x128
run r2_gn 0 $FULL -k regex:"gn_" -c 16 -- $K gn --first 2 --iters 1 --warmup 0
# (5) launch list of the eager training step (shares)
B="python bench.py --steps 1 --warmup 3 --no-graph --no-sampling --no-lora --no-cpu-baseline"
$B > $O/r2_launch_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4400 -c 1300 --csv --log-file $O/r2_launches.csv $B > $O/r2_launch_ncu.log 2>&1
echo "launches rc=$?"
ls -la $O/ | head -40; du -sh $O
