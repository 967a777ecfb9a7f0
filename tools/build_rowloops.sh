#!/bin/sh
# Builds the row-loop formulation study into tools/variants/ (git-ignored):
#   rowloop_variants.cu : -DVAR=<0..3,9,100..103> -DMM=<M>            shipped order vs reordered / pinned variants
#   rowloop_orders.cu   : -DCV=<0..4> -DGV=<0..4> -DPIN=<0|1> -DMM=<M>  contraction order x gradient order x pinned or not
# Run every binary on a B200 (each prints one line: algorithmic FMA rate of 128 / clk / SM); tools/sass_rf_model.py scores
# the same loops offline from `cuobjdump -sass`.
set -e
cd "$(dirname "$0")"
mkdir -p variants
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3"
for v in 0 1 2 3 9 100 101 102 103; do for m in 5 4; do $NV -DVAR=$v -DMM=$m -o variants/rl_${v}_${m} rowloop_variants.cu & done; wait; done
for cv in 0 1 2 3 4; do for gv in 0 1 2 3 4; do for pin in 1 0; do for m in 5 4; do
  $NV -DCV=$cv -DGV=$gv -DPIN=$pin -DMM=$m -o variants/rl2_${cv}_${gv}_${pin}_${m} rowloop_orders.cu &
done; done; wait; done; done
$NV -o variants/rowstep_bench rowstep_bench.cu
