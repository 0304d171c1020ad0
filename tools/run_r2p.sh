nvidia-smi -L | head -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
N_C3=50000 N_PLANAR=3000 timeout 900 python tools/parity_sweep.py > gpurun_out/r2p_parity_sweep.jsonl 2>&1; cut -c1-1200 gpurun_out/r2p_parity_sweep.jsonl
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2p_bench.err
