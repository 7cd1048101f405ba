set -x
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q -m gpu -x -s -k "single_decoder" 2>&1 | grep -v "^$" | tail -14 > gpurun_out/s16_pytest_single.log
bash scratch/run17.sh s16
