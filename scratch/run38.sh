set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "config3_final and f16x3f" 2>&1 | grep "^\[f16x3f\]\|   curve\|against the ref\|passed\|failed" > gpurun_out/s29_golden_mix.log
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q -m gpu -s -k "single_decoder and f16x3f" 2>&1 | grep "single decoder\|passed\|failed" > gpurun_out/s29_single_mix.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "full_config1 or full_pair_list" 2>&1 | tail -3 > gpurun_out/s29_long.log
