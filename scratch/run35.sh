set -x
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "edge_shapes and (f16x3-130 or f16-257 or f16x3-3-2)" > gpurun_out/s27_memcheck.log 2>&1
echo "rc=$?" >> gpurun_out/s27_memcheck.log
