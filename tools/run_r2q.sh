nvidia-smi -L | head -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:emc_flight -s 3 -c 1 -f -o gpurun_out/r2q_flight python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2q_ncu_full.log 2>&1; echo "flight rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:emc_strict -s 3 -c 1 -f -o gpurun_out/r2q_strict python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2q_ncu_strict.log 2>&1; echo "strict rc=$?"
