#!/bin/bash
# ncu captures of the kernels added late in r02 (run through gpurun; raw metric pages land in gpurun_out/):
#   ica_defl_pass (deflation FastICA one-pass kernel) and the tc_xb launches of inverse_transform.
set -x
OUT=gpurun_out
mkdir -p $OUT
timeout 300 ncu --set full --clock-control none -k regex:defl_pass -s 2 -c 2 -f -o $OUT/defl_full python profiles/tools/new_kernels_probe.py defl > $OUT/defl_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 4 -c 2 -f -o $OUT/inv_full python profiles/tools/new_kernels_probe.py inv > $OUT/inv_full.log 2>&1
for f in defl_full inv_full; do
  ncu -i $OUT/$f.ncu-rep --page raw --csv > $OUT/${f}_raw.csv 2>/dev/null
done
ncu -i $OUT/inv_full.ncu-rep --page source --csv > $OUT/inv_full_source.csv 2>/dev/null
rm -f $OUT/defl_full.ncu-rep $OUT/inv_full.ncu-rep
ls -la $OUT | tail -8
