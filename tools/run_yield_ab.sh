#!/bin/bash
# developer tool (GPU box): lane hand-back on / off for some builds.  usage: bash tools/run_yield_ab.sh <tag> <lib> ...
tag=$1; shift
nvidia-smi -L | head -1
for lib in "$@"; do for o in '{}' '{"lane_yield": false}'; do EMC_AB_OPTS="$o" EMC_LIB=$PWD/$lib timeout 300 python tools/ab_one.py 2>&1 | tail -1; done; done | tee gpurun_out/${tag}_ab.log
