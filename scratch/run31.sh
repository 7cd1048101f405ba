timeout 120 python scratch/stats_tc5.py scratch/variants_build/stats.so f16 > gpurun_out/s20_stats5_f16.log 2>&1
timeout 120 python scratch/stats_tc.py scratch/variants_build/stats.so f16x3 > gpurun_out/s20_stats3_f16x3.log 2>&1
timeout 120 python scratch/stats_tc.py scratch/variants_build/stats.so f16 > gpurun_out/s20_stats3_f16.log 2>&1
