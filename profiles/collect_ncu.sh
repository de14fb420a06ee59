#!/bin/bash
# Collects the ncu evidence kept under profiles/ (run on the GPU box through gpurun; outputs land in gpurun_out/,
# profiles/summarize_ncu.py turns them into the tracked summaries).
#   1. launch list of one randomized-PCA fit at 2M x 1024 (the c2 shape with fewer rows, so ncu's serialised
#      cold-cache replays stay short) -> per-kernel share of the step
#   2. `--set full` captures of the tcgen05 kernels (tc_xb fast / precise, tc_atb precise) and of the one-pass
#      FastICA kernel -> DRAM bytes per launch, tensor-pipe activity, registers, grid
#   3. `--set full` of the FP64 tensor-path kernels of exact PCA (first Gram launch, first pass-2 GEMM) at 2M x 512
#   4. launch list of the block Jacobi engine on a 2048 x 2048 SVD
set -x
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --rows 2000000 --steps 1 --warmup 1 --no-e2e --no-cpu"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c2_2Mrows.csv $B > $OUT/launches_c2.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:tc_gemm -c 10 -f -o $OUT/tc_full $B > $OUT/tc_full.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:ica_fused -c 2 -f -o $OUT/ica_full python bench.py --config c3 --steps 1 --warmup 1 --no-e2e --no-cpu > $OUT/ica_full.log 2>&1
C="python bench.py --config c4s --steps 1 --warmup 0 --no-e2e --no-cpu"
timeout 600 ncu --set full --clock-control none -k regex:atb_dmma -s 0 -c 1 -f -o $OUT/dmma_gram $C > $OUT/dmma_gram.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:gemm_nn_dmma -s 21 -c 1 -f -o $OUT/dmma_gemm $C > $OUT/dmma_gemm.log 2>&1
PETAL_JACOBI_BLOCK_MIN=256 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bj_|jacobi_sym" -c 6000 --csv --log-file $OUT/launches_block_jacobi.csv python -m pytest -q tests/test_gpu_configs.py -k "block_jacobi_svd and 2048" > $OUT/launches_bj.log 2>&1
# gpurun copies back at most 64 MiB: keep the raw metric pages, drop the reports
for f in tc_full ica_full dmma_gram dmma_gemm; do
  ncu -i $OUT/$f.ncu-rep --page raw --csv > $OUT/${f}_raw.csv 2>/dev/null
  rm -f $OUT/$f.ncu-rep
done
ls -la $OUT
