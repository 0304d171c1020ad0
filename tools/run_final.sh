#!/bin/bash
# developer tool (GPU box): the round-end sequence — gpu tests, smoke, bench (both arms), launch list, one full capture
tag=${1:-final}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -c 200 gpurun_out/${tag}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:emc_flight -s 3 -c 1 -f -o gpurun_out/${tag}_flight python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_ncu_full.log 2>&1; echo "full rc=$?"
