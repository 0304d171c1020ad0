#!/bin/bash
# developer tool (GPU box): hand-back point sweep.  usage: bash tools/run_yield_sweep.sh <tag> <lib> <step> ...
tag=$1; lib=$2; shift 2
for p in "$@"; do echo "step $p"; EMC_YIELD_STEP=$p EMC_LIB=$PWD/$lib timeout 300 python tools/ab_one.py 2>&1 | tail -1; done | tee gpurun_out/${tag}_sweep.log
