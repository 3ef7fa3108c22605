#!/bin/bash
# Builds tuning variants of libebncuda.so (threads per block x vehicle-kernel blocks per SM) into
# enhancedbayesiannetworks.jl_b200/tune/. Usage: tools/build_variants.sh "256,4 128,6 ..."
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $ROOT/enhancedbayesiannetworks.jl_b200/tune
for v in $1; do
  T=${v%,*}; B=${v#*,}
  D=/tmp/ebn_variant_${T}_${B}${SUFFIX}
  rm -rf $D && mkdir -p $D/pkg/csrc $D/include && cp $ROOT/enhancedbayesiannetworks.jl_b200/csrc/*.cu* $ROOT/enhancedbayesiannetworks.jl_b200/csrc/*.h $ROOT/enhancedbayesiannetworks.jl_b200/csrc/*.inc $ROOT/enhancedbayesiannetworks.jl_b200/csrc/Makefile $ROOT/enhancedbayesiannetworks.jl_b200/csrc/embed_sources.py $ROOT/enhancedbayesiannetworks.jl_b200/csrc/build_id.py $D/pkg/csrc/ && cp $ROOT/include/ebncuda.h $D/include/
  (cd $D/pkg/csrc && make -j8 EXTRA="-DEBN_THREADS=$T -DEBN_VEHICLE_BLOCKS=$B $EXTRA_FLAGS" OUT=$ROOT/enhancedbayesiannetworks.jl_b200/tune/libebncuda_${T}_${B}${SUFFIX}.so > /dev/null 2>&1) &
done
wait
ls -la $ROOT/enhancedbayesiannetworks.jl_b200/tune/
