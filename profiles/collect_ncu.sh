#!/bin/bash
# Collects the ncu evidence kept under profiles/ (run on the GPU box through gpurun; outputs land in gpurun_out/).
#   1. launch list of one randomized-PCA fit at 2M x 1024 (the c2 shape with fewer rows, so ncu's serialised
#      cold-cache replays stay short) -> per-kernel share of the step
#   2. `--set full` captures of the tcgen05 kernels (tc_xb fast / precise, tc_atb precise) and of the one-pass
#      FastICA kernel -> DRAM bytes per launch, tensor-pipe activity, registers, grid
set -x
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --rows 2000000 --steps 1 --warmup 1 --no-e2e --no-cpu"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c2_2Mrows.csv $B > $OUT/launches_c2.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:tc_gemm -c 11 -f -o $OUT/tc_full $B > $OUT/tc_full.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:ica_fused -c 2 -f -o $OUT/ica_full python bench.py --config c3 --steps 1 --warmup 1 --no-e2e --no-cpu > $OUT/ica_full.log 2>&1
# gpurun copies back at most 64 MiB: keep the raw metric pages, drop the reports
ncu -i $OUT/tc_full.ncu-rep --page raw --csv > $OUT/tc_full_raw.csv 2>/dev/null
ncu -i $OUT/ica_full.ncu-rep --page raw --csv > $OUT/ica_full_raw.csv 2>/dev/null
rm -f $OUT/tc_full.ncu-rep $OUT/ica_full.ncu-rep
ls -la $OUT
