# round-2 checkpoint: gpu tests, the bench line (both arms), launch list, one full capture of the flight kernel
nvidia-smi -L | head -2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 900 python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2k_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2k_bench_ref.json 2> gpurun_out/r2k_bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2k_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2k_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:emc_flight -s 3 -c 1 -f -o gpurun_out/r2k_flight python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2k_ncu_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out | head -20
