set -x
timeout 900 python -m pytest tests/test_gpu_dropin.py -q -m gpu -x -s -k "single_decoder or cov" 2>&1 | tail -40 > gpurun_out/r2k_pytest_dropin.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "config3_final" 2>&1 | tail -60 > gpurun_out/r2k_pytest_config3.log
