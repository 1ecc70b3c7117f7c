mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 200 --warmup 20 > gpurun_out/bench2.log 2> gpurun_out/bench2.err; echo "bench rc=$?" >> gpurun_out/bench2.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench2_ref.log 2>&1
tail -6 gpurun_out/pytest_gpu.log; cat gpurun_out/bench2.log; tail -3 gpurun_out/bench2.err; cat gpurun_out/bench2_ref.log
