set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -k "more_than_two or edge_shapes" 2>&1 | tail -12 > gpurun_out/s14_pytest_m.log
bash scratch/run17.sh s14
