#!/bin/bash
# The validation sequence behind the r02 artefacts in profiles/ (run on a B200 box through gpurun from the repo root:
#   gpurun --timeout 2000 -- 'bash profiles/tools/run_validation.sh' ; outputs land in gpurun_out/).
#   1. the whole GPU test suite and the smoke entry
#   2. bench lines: c2 (default), c3, c1, c4 (with its e2e from a 65.5 GB pinned host matrix), c5 and c2q7
#   3. c2 forced out of core (--host-staging 2), with and without the Gram-mode power iterations
#   4. compute-sanitizer memcheck / racecheck over the kernels added late in r02
# Multi-GPU (gpurun --gpus 2): torchrun ... tests/dist_gpu_check.py ; torchrun ... bench.py --gpus 2 [--config c5]
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
( time timeout 500 python -m pytest tests -q -m gpu ) > $O/val_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/val_pytest.log
( time timeout 120 python __graft_entry__.py smoke ) > $O/val_smoke.log 2>&1; echo "smoke rc=$?"
timeout 400 python bench.py > $O/val_bench_c2.json 2> $O/val_bench_c2.err; echo "c2 rc=$?"
timeout 200 python bench.py --config c3 > $O/val_bench_c3.json 2> $O/val_bench_c3.err; echo "c3 rc=$?"
timeout 200 python bench.py --config c1 > $O/val_bench_c1.json 2> $O/val_bench_c1.err; echo "c1 rc=$?"
timeout 400 python bench.py --config c4 --steps 1 --warmup 1 > $O/val_bench_c4.json 2> $O/val_bench_c4.err; echo "c4 rc=$?"
timeout 300 python bench.py --config c5 --no-cpu --no-e2e > $O/val_bench_c5.json 2> $O/val_bench_c5.err; echo "c5 rc=$?"
timeout 300 python bench.py --config c2q7 --no-cpu > $O/val_bench_c2q7.json 2> $O/val_bench_c2q7.err; echo "c2q7 rc=$?"
timeout 300 python bench.py --no-cpu --steps 2 --warmup 1 --host-staging 2 > $O/val_bench_c2_ooc.json 2> $O/val_bench_c2_ooc.err; echo "ooc rc=$?"
timeout 300 python bench.py --no-cpu --steps 1 --warmup 1 --host-staging 2 --no-host-gram > $O/val_bench_c2_ooc_plain.json 2> $O/val_bench_c2_ooc_plain.err; echo "ooc plain rc=$?"
( time timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_deflation.py tests/test_gpu_streaming.py -q -m gpu \
  -k "golden or (deflation_fit and 20000) or (instantiations and (20 or 64 or 33 or 130)) or (rpca_f32_host and 25076) or (pca_host and 20557) or (gram_mode and 30000) or (fastica_host and float64)" ) > $O/val_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $O/val_memcheck.log
( time timeout 200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_deflation.py -q -m gpu -k "golden or (instantiations and (float32-64 or float64-33))" ) > $O/val_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 $O/val_racecheck.log
