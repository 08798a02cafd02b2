#!/bin/bash
# Round-2 end-of-round validation on one GPU: every GPU test, smoke(), the default bench line (headline, with the CPU arm), the
# reference arm, and the other BASELINE configurations (c2, c3, c5).  Outputs are copied into profiles/ by hand afterwards.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method thread > gpurun_out/r02_pytest_gpu_all.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -3 gpurun_out/r02_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_c4_default.json 2> gpurun_out/r02_bench_c4_default_err.txt; cut -c1-200 gpurun_out/r02_bench_c4_default.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_c4_reference.json 2> gpurun_out/r02_bench_c4_reference_err.txt; cut -c1-200 gpurun_out/r02_bench_c4_reference.json
for cfg in c2 c3; do
  timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 > gpurun_out/r02_bench_$cfg.json 2> gpurun_out/r02_bench_${cfg}_err.txt; cut -c1-200 gpurun_out/r02_bench_$cfg.json
done
timeout 900 python bench.py --config c5 --steps 1 --warmup 3 --no-cpu > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5_err.txt; cut -c1-200 gpurun_out/r02_bench_c5.json
timeout 120 scripts/micro/mma_rate 148 quick > gpurun_out/r02_mma_rate_mn.txt 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_final_clocks.txt
du -sk gpurun_out > gpurun_out/r02_final_status.txt
