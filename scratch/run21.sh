timeout 120 python scratch/stats_tc.py scratch/variants_build/stats.so f16 > gpurun_out/s10_stats_f16.log 2>&1
