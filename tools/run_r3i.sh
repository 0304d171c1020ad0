nvidia-smi -L | head -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
N_C3=50000 N_PLANAR=3000 timeout 900 python tools/parity_sweep.py > gpurun_out/r3i_parity_sweep.jsonl 2>&1; cut -c1-1100 gpurun_out/r3i_parity_sweep.jsonl
tools/ab.sh erpl_monte_carlo_sim_b200/libemc.so 2>&1 | tee gpurun_out/r3i_ab.log
