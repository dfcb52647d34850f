#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'jaccard_featurize_kernel|tanimoto_bits_kernel' -s 2 -c 2 \
    -o gpurun_out/r01_similarity_ncu -f python tools/bench_similarity.py > gpurun_out/ncu_similarity.log 2>&1
echo "ncu exit $?"
