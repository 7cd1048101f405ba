set -x
python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/r2_pytest1.log
for i in 1 2 3 4 5; do python -m pytest tests -q -m gpu -k "test_full_pair_list or test_split_row_lists" 2>&1 | tail -3 >> gpurun_out/r2_pytest_repeat.log; done
python bench.py --steps 10 --warmup 3 --precision f16 --no-cpu > gpurun_out/r2_bench_f16_a.json 2> gpurun_out/r2_bench_f16_a.err
python bench.py --steps 10 --warmup 3 --precision f16x3 --no-cpu > gpurun_out/r2_bench_f16x3_a.json 2> gpurun_out/r2_bench_f16x3_a.err
python scratch/full_job.py 1000 f16,f16x3,tf32,fp32 > gpurun_out/r2_full_job.log 2>&1
