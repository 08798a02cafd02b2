#!/bin/bash
# round-end style validation: all GPU tests, smoke, headline bench, ncu launch list + DRAM traffic of the top kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread > gpurun_out/pytest_gpu_all.log 2>&1; tail -3 gpurun_out/pytest_gpu_all.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 600 python bench.py --batch 1 --steps 3 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_512x48_b1.json; cut -c1-160 gpurun_out/bench_512x48_b1.json
timeout 600 python bench.py --batch 2 --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_512x48_b2.json; cut -c1-160 gpurun_out/bench_512x48_b2.json
for cw in "fianl_conv wgrad conv_wgrad_tc_kernel" "3dconv_up4 wgrad conv_wgrad_tc_kernel" "3dconv_up4 fprop conv_tc_kernel" "3dconv_up4 dgrad conv_tc_kernel"; do
  set -- $cw
  timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:$3 -s 2 -c 1 --csv --log-file gpurun_out/traffic_$1_$2.csv python scripts/prof_conv.py $1 $2 1 > /dev/null 2>&1
  tail -4 gpurun_out/traffic_$1_$2.csv | cut -d, -f5,13-15 | cut -c1-200
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1215 -c 1215 --csv --log-file gpurun_out/launches_512x48_b1_final.csv \
   python bench.py --batch 1 --steps 1 --warmup 1 --no-cpu --profile-steps 1 > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-120
