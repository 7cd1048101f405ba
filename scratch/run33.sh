timeout 120 python scratch/stats_tc5.py scratch/variants_build/stats.so f16 > gpurun_out/s22_stats5_f16.log 2>&1
