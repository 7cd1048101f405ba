timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -k "more_than_two or edge_shapes or errors_are_loud or six_mc" 2>&1 | tail -12 > gpurun_out/s15_pytest_m.log
