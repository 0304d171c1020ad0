nvidia-smi -L | head -1
tools/ab.sh variants/libemc_diet1.so variants/libemc_diet2.so variants/libemc_diet2n.so variants/libemc_b416.so 2>&1 | tee gpurun_out/r2m_ab.log
