#!/bin/bash
# A/B builds of libhdn.so (selected at run time with HDN_LIB=<path>): compile-time variants of conv_tc.cu, and the kernel files
# of an earlier commit inside today's library, so that two forms can be timed side by side in one GPU call.
#   scripts/build_variants.sh <name> [nvcc -D flags...]            e.g.  nofold -DHDN_NO_FOLD
#   scripts/build_variants.sh <name> @<commit>                     conv_tc.cu / conv_tc2_wgrad.cu / tc_common.cuh of <commit>
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=h-denseunet_b200/variants; mkdir -p $out
src=h-denseunet_b200/csrc
files="api.cu conv_simt.cu conv_tc.cu conv_tc_wgrad.cu conv_tc2_wgrad.cu elementwise.cu postproc.cu augment.cu"
flags="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared"
if [[ "$1" == @* ]]; then
  c=${1#@}; tmp=$(mktemp -d); mkdir -p $tmp/h-denseunet_b200/csrc $tmp/include
  cp $src/*.cu $src/*.cuh $tmp/h-denseunet_b200/csrc/; cp include/hdn.h $tmp/include/
  for f in conv_tc.cu conv_tc2_wgrad.cu tc_common.cuh; do git show $c:$src/$f > $tmp/h-denseunet_b200/csrc/$f; done
  # switches that did not exist at that commit
  printf '\nvoid hdn_tc_sw128_set(int) {}\nvoid hdn_tc_x3fold_set(int) {}\n' >> $tmp/h-denseunet_b200/csrc/conv_tc.cu
  (cd $tmp/h-denseunet_b200/csrc && nvcc $flags -o /root/repo/$out/libhdn_$name.so $files -lcuda)
  rm -rf $tmp
else
  (cd $src && nvcc $flags "$@" -o ../variants/libhdn_$name.so $files -lcuda)
fi
ls -la $out/libhdn_$name.so
