timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "compaction or scheduling or variants or launch" 2>&1 | tail -4
tools/ab.sh erpl_monte_carlo_sim_b200/libemc.so > gpurun_out/r2d_ab.log 2>&1
EMC_AB_OPTS='{"compaction":false}' tools/ab.sh erpl_monte_carlo_sim_b200/libemc.so >> gpurun_out/r2d_ab.log 2>&1
cat gpurun_out/r2d_ab.log
