timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "config5_shape or multi_curve_windows or write_only_inside" 2>&1 | tail -6 > gpurun_out/s33_pytest.log
