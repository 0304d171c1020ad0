nvidia-smi -L | head -1
tools/ab.sh variants/libemc_base.so variants/libemc_diet1.so 2>&1 | tee gpurun_out/r2l_ab.log
EMC_AB_OPTS='{"block_threads":256,"blocks_per_sm":2}' tools/ab.sh erpl_monte_carlo_sim_b200/libemc.so 2>&1 | tee -a gpurun_out/r2l_ab.log
EMC_AB_OPTS='{"blocks_per_sm":4}' tools/ab.sh erpl_monte_carlo_sim_b200/libemc.so 2>&1 | tee -a gpurun_out/r2l_ab.log
