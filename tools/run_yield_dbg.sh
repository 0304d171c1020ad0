#!/bin/bash
# developer tool (GPU box): timeline of the lane hand-back on the C3 / planar 100 k batches (build with -DEMC_YIELD_DEBUG)
tag=$1; lib=$2
for o in '{}' '{"lane_yield": false}'; do echo "== $o"; EMC_YIELD_DEBUG=1 EMC_AB_OPTS="$o" EMC_LIB=$PWD/$lib timeout 300 python tools/ab_one.py 2>&1 | grep "emc-dbg" | tail -18; done | tee gpurun_out/${tag}_dbg.log
