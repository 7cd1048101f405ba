timeout 120 python scratch/tmem_lat.py > gpurun_out/s18_tmem_lat.log 2>&1
