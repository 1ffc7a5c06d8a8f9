#!/bin/bash
# GPU session Z: two lanes per permutation (k_leaf_pair) against one thread per leaf, small trees and the G1 shape; digests must agree.
mkdir -p gpurun_out
( for spec in "13 9800" "14 9800" "14 2664" "15 4096" "16 2816" "17 1680" "12 1024"; do echo "== logL ncols = $spec"; ./tools/microbench/pb_cur $spec | grep -E "bs=128 minb=1 |pair"; done ) > gpurun_out/r2z_poseidon_pair.txt 2>&1
cat gpurun_out/r2z_poseidon_pair.txt
