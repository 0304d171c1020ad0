#!/bin/bash
# developer tool (GPU box): A/B list of builds, then one full ncu capture of the in-tree flight kernel on the bench workload
# usage: bash tools/run_prof_ab.sh <tag> lib1 lib2 ...
tag=$1; shift
nvidia-smi -L | head -1
tools/ab.sh "$@" 2>&1 | tee gpurun_out/${tag}_ab.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:emc_flight -s 3 -c 1 -f -o gpurun_out/${tag}_flight python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_ncu_full.log 2>&1; echo "full rc=$?"
