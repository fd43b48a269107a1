#!/bin/bash
# usage (repo root, GPU box): tools/profile_r2.sh <tag>
#   1. ncu --set full of the two kernels BASELINE.json names (stand-alone loop of bench.py: tools/kernels.py)
#   2. ncu launch list (gpu__time_duration.sum only) of one full-size bench step
tag=${1:-r02}
ncu --set full --import-source on --clock-control none -k regex:gemm_split_kernel -c 1 --launch-skip 3 \
    -o gpurun_out/${tag}_gemm python tools/kernels.py > gpurun_out/${tag}_gemm.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:scatter_fwd -c 1 --launch-skip 2 \
    -o gpurun_out/${tag}_scatter python tools/kernels.py > gpurun_out/${tag}_scatter.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_launches.log 2>&1
