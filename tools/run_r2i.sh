DUMP_DIFFER=gpurun_out/r2i_differ N_C3=50000 N_PLANAR=0 timeout 600 python tools/parity_sweep.py 2>&1 | cut -c1-700
EMC_LIB=$PWD/variants/libemc_strict_early.so N_C3=50000 N_PLANAR=0 timeout 600 python tools/parity_sweep.py 2>&1 | cut -c1-700
EMC_LIB=$PWD/variants/libemc_strict_early.so tools/ab.sh variants/libemc_strict_early.so 2>&1 | tail -1
