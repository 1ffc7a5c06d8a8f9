#!/bin/bash
# GPU session Y: leaf-hash launch-bound variants in the small-tree regime (Fq12: 2^14 leaves x 9 800 columns) against the G1 shape.
mkdir -p gpurun_out
( for spec in "14 9800" "15 4096" "16 2816" "17 1680" "13 9800"; do echo "== logL ncols = $spec"; ./tools/microbench/pb_cur $spec; done ) > gpurun_out/r2y_poseidon_small_trees.txt 2>&1
cat gpurun_out/r2y_poseidon_small_trees.txt
