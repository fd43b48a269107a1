#!/bin/bash
# usage: tools/profile.sh  (run from the repo root) <tag> <batch> <unroll>   -> gpurun_out/launches_<tag>.csv (ncu launch list of one bench step)
tag=$1; b=$2; t=$3
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --batch $b --unroll $t --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$tag.log 2>&1
