#!/bin/bash
# developer tool (GPU box): reference arm, ncu launch list of the bench, one full capture of the flight kernel, hand-back point sweep
# usage: tools/run_checkpoint2.sh <tag>   -> gpurun_out/<tag>_*
tag=${1:-ckpt}
nvidia-smi -L | head -1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:emc_flight -s 3 -c 1 -f -o gpurun_out/${tag}_flight python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_ncu_full.log 2>&1; echo "full rc=$?"
for p in 600 800 1200; do echo "step $p"; EMC_YIELD_STEP=$p EMC_LIB=$PWD/erpl_monte_carlo_sim_b200/libemc.so timeout 300 python tools/ab_one.py 2>&1 | tail -1; done | tee gpurun_out/${tag}_sweep.log
ls -la gpurun_out | grep ${tag}
