tools/ab.sh erpl_monte_carlo_sim_b200/libemc.so variants/libemc_r1phys.so > gpurun_out/r2c_ab.log 2>&1
EMC_AB_OPTS='{"block_threads":256,"blocks_per_sm":2}' tools/ab.sh erpl_monte_carlo_sim_b200/libemc.so >> gpurun_out/r2c_ab.log 2>&1
EMC_AB_OPTS='{"block_threads":128,"blocks_per_sm":4}' tools/ab.sh erpl_monte_carlo_sim_b200/libemc.so >> gpurun_out/r2c_ab.log 2>&1
cat gpurun_out/r2c_ab.log
