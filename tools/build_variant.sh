#!/bin/bash
# A/B build of the library with extra nvcc flags: tools/build_variant.sh NAME "-DPCG_WSM_LAYOUT=28 ..."  -> ab/lib_NAME.so
# Only inst_cstr.cu differs between headline-kernel variants, so the other objects are copied from the main build.
set -e
NAME=$1; EXTRA=$2
cd "$(dirname "$0")/../pcgym_b200/csrc"
mkdir -p ../../ab build_$NAME
for o in build/*.o; do b=$(basename $o); [ "$b" = inst_cstr.o ] || cp -u $o build_$NAME/$b; done
touch -c build_$NAME/*.o
make -s OUT=../../ab/lib_$NAME.so BUILD=build_$NAME EXTRA="$EXTRA" -j8
grep -A2 "wsm_rollout_kernelINS_5CstrWENS_5Tsit5ELi1ELi1ELb0ELb1ELb0E" build_$NAME/inst_cstr.ptxas.log | grep -E "registers|spill" | head -4
