#!/bin/bash
# Round-1 last validation call: candidate defaults HDN_TC_FASTX=2 (two-source warp-per-chunk transform) and
# HDN_POOL_FAST=1 (vector max-pool backward), precision "mixed"; steps by priority, each bounded.
set +e
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 )); }
step() { echo "$1 rc=$2 t=$(el)s" >> gpurun_out/status_c.txt; }
: > gpurun_out/status_c.txt
export HDN_TC_FASTX=2 HDN_POOL_FAST=1
PT="python -m pytest -q --tb=line -p no:cacheprovider"
timeout 200 $PT tests -m gpu > gpurun_out/c1_full_fastx2_pool1.txt 2>&1; step c1_full $?
timeout 120 python tests/grad_errors.py mixed bf16x3 bf16 > gpurun_out/c2_grad_errors.txt 2> gpurun_out/c2_err.txt; step c2_grad_errors $?
timeout 120 python bench.py --steps 2 --warmup 3 --no-cpu --precision mixed > gpurun_out/c3_bench_mixed.json 2> gpurun_out/c3_err.txt; step c3_bench_mixed $?
timeout 120 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/c4_bench_bf16.json 2> gpurun_out/c4_err.txt; step c4_bench_bf16 $?
HDN_POOL_FAST=0 timeout 60 $PT tests/test_gpu_tc.py -k "two_src or skip" > gpurun_out/c5_two_src_only.txt 2>&1; step c5_two_src $?
HDN_TC_FASTX=1 timeout 90 $PT tests/test_gpu_models.py -k "fp32 and (hybrid or unet2d)" > gpurun_out/c6_pool_only.txt 2>&1; step c6_pool $?
timeout 90 python __graft_entry__.py smoke > gpurun_out/c7_smoke.txt 2>&1; step c7_smoke $?
cat gpurun_out/status_c.txt
