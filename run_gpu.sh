mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "loss or e2e" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
: > gpurun_out/sweep.log
python tools/loss_sweep.py >> gpurun_out/sweep.log 2>&1
for f in tools/variants/*.so; do CELLULUS_B200_LIB=$PWD/$f timeout 120 python tools/loss_sweep.py >> gpurun_out/sweep.log 2>&1; done
timeout 300 python bench.py --skip-cpu --skip-detect > gpurun_out/bench_redux.json 2> gpurun_out/bench_redux.err
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/sweep.log; python -c "
import json;d=json.load(open('gpurun_out/bench_redux.json'));print('ms',d['ms_per_step'],'frac',d['roofline']['frac'],'planar',d['planar']['ms_per_step'],d['planar']['roofline']['frac'])"
