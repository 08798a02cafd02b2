#!/bin/bash
# Round 2, call B: the whole GPU suite (new parity tests incl. full-shape and headline-size layers), the per-kernel
# split of the tc2 weight gradient (ncu launch list), and one default bench line.
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity2.py 2>&1 | tail -15 > gpurun_out/r2b_pytest_gpu_old.txt
timeout 3000 python -m pytest tests/test_gpu_parity2.py -q -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r2b_pytest_parity2.txt
for c in fianl_conv 3dconv_up4; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_$c.csv \
     python scripts/prof_conv.py $c wgrad 3 1 > gpurun_out/r2b_ncu_$c.log 2>&1
done
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench_default.json 2> gpurun_out/r2b_bench_err.txt
echo done > gpurun_out/r2b_status.txt
