timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "compaction" 2>&1 | tail -3
timeout 600 python tools/api_probe.py 2>&1 | grep -v "^$" | cut -c1-170 > gpurun_out/r2f_api_probe.log; head -120 gpurun_out/r2f_api_probe.log
