#!/bin/bash
# A/B build with extra nvcc flags where only ONE translation unit differs: tools/build_variant_x.sh NAME inst_extraction "-DPCG_LANES_MINB=6"
set -e
NAME=$1; UNIT=$2; EXTRA=$3
cd "$(dirname "$0")/../pcgym_b200/csrc"
mkdir -p ../../ab build_$NAME
for o in build/*.o; do b=$(basename $o); [ "$b" = $UNIT.o ] || cp -u $o build_$NAME/$b; done
touch -c build_$NAME/*.o
make -s OUT=../../ab/lib_$NAME.so BUILD=build_$NAME EXTRA="$EXTRA" -j8
grep -A2 "lanes_rollout_kernelINS_11ExtractionTILi5EEENS_5Tsit5ELi1E" build_$NAME/$UNIT.ptxas.log | grep -E "registers|spill" | head -4
