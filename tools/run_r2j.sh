timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
N_C3=50000 N_PLANAR=3000 timeout 600 python tools/parity_sweep.py > gpurun_out/r2j_parity_sweep.jsonl 2>&1; cat gpurun_out/r2j_parity_sweep.jsonl | cut -c1-900
tools/ab.sh erpl_monte_carlo_sim_b200/libemc.so 2>&1 | tail -1
