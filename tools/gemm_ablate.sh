#!/bin/bash
# usage: tools/gemm_ablate.sh "<debug values>" "<shapes>" "<bn:mc modes>"   (DSB_GEMM_DEBUG ablation matrix, dev only)
for d in ${1:-0 1 2 4 3 5 6}; do DSB_GEMM_DEBUG=$d timeout 120 python tools/gemmab.py 10 ${2:-ffn1,ffn2} ${3:-0:1,256:4} 2>&1 | tail -${4:-4}; done
