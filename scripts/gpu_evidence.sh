#!/bin/bash
# Round-2 evidence at HEAD (one GPU): launch list of one headline `mixed` step, `ncu --set full` of the dominant kernels at the
# headline shape (2 slabs x 48 slices), section captures of the element-wise kernels inside a real step, clocks.
# Reports -> gpurun_out/*.ncu-rep; digested into profiles/ by scripts/ncu_digest.py / launch_summary.py on the build box.
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# (1) every launch of the second of two headline steps (batch 2, 512x512x48, mixed, dropout on)
timeout 1500 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_step.csv python scripts/one_step.py 2 48 512 > gpurun_out/r02_one_step.log 2>&1
# (2) dominant kernels, --set full, headline shape
# The reports are digested on the box (raw metric page as csv, source page gzipped for the kernels under study) and then
# deleted: eight `--set full` reports with imported sources exceed what gpurun copies back (64 MiB).
cap() {  # name pass precision kernel-regex skip tag [src]
  HDN_PROF_FULL=1 timeout 600 $NCU --set full --import-source on -k regex:$4 -s $5 -c 1 -f -o gpurun_out/r02_ncu_$6 python scripts/prof_conv.py $1 $2 1 $3 > gpurun_out/r02_ncu_$6.log 2>&1
  tail -2 gpurun_out/r02_ncu_$6.log
  ncu -i gpurun_out/r02_ncu_$6.ncu-rep --page raw --csv > gpurun_out/r02_ncu_$6.raw.csv 2>> gpurun_out/r02_ncu_$6.log
  if [ -n "$7" ]; then ncu -i gpurun_out/r02_ncu_$6.ncu-rep --page source --csv 2>> gpurun_out/r02_ncu_$6.log | gzip -9 > gpurun_out/r02_ncu_$6.source.csv.gz; fi
  rm -f gpurun_out/r02_ncu_$6.ncu-rep
}
cap 3dconv_up4 fprop 2 conv_tc_kernel 2 fprop_x3_3dconv_up4 src
cap fianl_conv dgrad 2 conv_tc_kernel 2 dgrad_x3_fianl_conv
cap fianl_conv wgrad 1 conv_wgrad_tc2_kernel 2 wgrad_tc2_fianl_conv
cap fianl_conv wgrad 1 act_pack_bf16_kernel 4 act_pack_fianl_conv
cap dense2_x1 dgrad 2 conv_tc_kernel 2 dgrad_x3_dense2_x1 src
cap dense2_x1 fprop 2 conv_tc_kernel 2 fprop_x3_dense2_x1 src
cap dense2_x2 fprop 2 conv_tc_kernel 2 fprop_x3_dense2_x2
# (3) the other kernels inside a real (reduced-batch: 1 slab x 16 slices) step: speed-of-light + memory sections
timeout 1200 $NCU --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy \
   -k regex:'pool|bn_|wce_|sgd_|dropout|triplets|cat4|colstats|colsum|zero_window|pack_weights|small_n|conv_wgrad_tc_kernel' \
   -c 120 -f -o gpurun_out/r02_ncu_elementwise python scripts/one_step.py 1 16 512 > gpurun_out/r02_ncu_elementwise.log 2>&1
ncu -i gpurun_out/r02_ncu_elementwise.ncu-rep --page raw --csv > gpurun_out/r02_ncu_elementwise.raw.csv 2>> gpurun_out/r02_ncu_elementwise.log
rm -f gpurun_out/r02_ncu_elementwise.ncu-rep
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/r02_evidence_clocks.txt
du -sk gpurun_out > gpurun_out/r02_evidence_files.txt; ls -la gpurun_out >> gpurun_out/r02_evidence_files.txt 2>&1
echo done > gpurun_out/r02_evidence_status.txt
