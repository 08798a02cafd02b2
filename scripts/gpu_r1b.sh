#!/bin/bash
# Round-1 late validation call (tight GPU budget): bf16x3 precision + warp-per-chunk transform (HDN_TC_FASTX).
# Steps are ordered by priority; every step is bounded by its own timeout and logs into gpurun_out/.
set +e
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 )); }
step() { echo "$1 rc=$2 t=$(el)s" >> gpurun_out/status.txt; }
: > gpurun_out/status.txt
python -c "import torch; print(torch.cuda.get_device_name(0), torch.cuda.device_count())" > gpurun_out/env.txt 2>&1
step import $?
PT="python -m pytest -q --tb=line -p no:cacheprovider"
HDN_TC_FASTX=1 timeout 300 $PT tests/test_gpu_tc.py > gpurun_out/a_tc_fastx1.txt 2>&1; step a_tc_fastx1 $?
HDN_TC_FASTX=1 timeout 240 $PT tests/test_gpu_models.py -k "bf16" --tb=short > gpurun_out/b_models_fastx1.txt 2>&1; step b_models_fastx1 $?
HDN_TC_FASTX=1 timeout 240 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/c_bench_bf16_fastx1.json 2> gpurun_out/c_err.txt; step c_bench_bf16_fastx1 $?
HDN_TC_FASTX=1 timeout 240 python bench.py --steps 2 --warmup 3 --no-cpu --precision bf16x3 > gpurun_out/d_bench_x3_fastx1.json 2> gpurun_out/d_err.txt; step d_bench_x3_fastx1 $?
HDN_TC_FASTX=0 timeout 200 $PT tests/test_gpu_tc.py -k "bf16x3" > gpurun_out/e_tc_x3_fastx0.txt 2>&1; step e_tc_x3_fastx0 $?
HDN_TC_FASTX=0 timeout 200 $PT tests/test_gpu_models.py -k "bf16x3" --tb=short > gpurun_out/f_models_x3_fastx0.txt 2>&1; step f_models_x3_fastx0 $?
HDN_TC_FASTX=0 timeout 240 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/g_bench_bf16_fastx0.json 2> gpurun_out/g_err.txt; step g_bench_bf16_fastx0 $?
HDN_TC_FASTX=1 timeout 300 $PT tests -m gpu > gpurun_out/h_full_fastx1.txt 2>&1; step h_full_fastx1 $?
tail -3 gpurun_out/a_tc_fastx1.txt gpurun_out/b_models_fastx1.txt
cat gpurun_out/status.txt
