nvidia-smi -L | head -1
timeout 300 tools/ab.sh erpl_monte_carlo_sim_b200/libemc.so variants/libemc_diet5.so 2>&1 | tee gpurun_out/r2s_ab.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
