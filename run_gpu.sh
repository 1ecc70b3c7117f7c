mkdir -p gpurun_out
timeout 1700 python tools/ms_sweep.py 16 > gpurun_out/ms_sweep.log 2>&1
cat gpurun_out/ms_sweep.log
