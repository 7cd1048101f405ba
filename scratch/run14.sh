set -x
timeout 300 python scratch/stats_tc.py scratch/variants_build/stats.so f16x3 > gpurun_out/r2n_stats_f16x3.log 2>&1
timeout 300 python scratch/stats_tc.py scratch/variants_build/stats.so f16 > gpurun_out/r2n_stats_f16.log 2>&1
