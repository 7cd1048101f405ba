set -x
tag=$1
bash scratch/run17.sh $tag
timeout 120 python scratch/stats_tc.py scratch/variants_build/stats.so f16x3 > gpurun_out/${tag}_stats_f16x3.log 2>&1
timeout 120 python scratch/stats_tc.py scratch/variants_build/stats.so f16 > gpurun_out/${tag}_stats_f16.log 2>&1
