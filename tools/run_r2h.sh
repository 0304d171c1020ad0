timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
N_C3=50000 N_PLANAR=3000 timeout 600 python tools/parity_sweep.py > gpurun_out/r2h_parity_sweep.jsonl 2>&1; cat gpurun_out/r2h_parity_sweep.jsonl | cut -c1-1500
tools/ab.sh erpl_monte_carlo_sim_b200/libemc.so > gpurun_out/r2h_ab.log 2>&1
EMC_AB_OPTS='{"strict_tail":false}' tools/ab.sh erpl_monte_carlo_sim_b200/libemc.so >> gpurun_out/r2h_ab.log 2>&1
cat gpurun_out/r2h_ab.log
