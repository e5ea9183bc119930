#!/bin/bash
# A/B of a library variant (MC_LIB_VARIANT build) against the product library on one GPU
set -u
mkdir -p gpurun_out
V=${1:-cta4}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
python bench.py --steps 50 --warmup 5 --scaling-only > gpurun_out/ab_bench_product.json 2> gpurun_out/ab_bench_product.err
MESHCLUST_B200_LIB=$PWD/meshclust_b200/libmeshclust_b200_$V.so python bench.py --steps 50 --warmup 5 --scaling-only > gpurun_out/ab_bench_$V.json 2> gpurun_out/ab_bench_$V.err
