# what the driver runs first on the GPU box
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2v_pytest_full.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1
