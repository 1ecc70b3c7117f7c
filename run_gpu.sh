mkdir -p gpurun_out
cd tests && timeout 900 python -m pytest . -m gpu -q -x -k "greedy or tta" > ../gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> ../gpurun_out/pytest_gpu.log; cd ..
tail -25 gpurun_out/pytest_gpu.log
