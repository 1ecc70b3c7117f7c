mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --skip-cpu > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench rc=$?" >> gpurun_out/bench3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_detect.csv python tools/profile_target.py detect > gpurun_out/ncu_launches.log 2>&1
tail -15 gpurun_out/pytest_gpu.log; python -c "
import json;d=json.load(open('gpurun_out/bench3.json'));print('loss ms',d['ms_per_step'],'frac',d['roofline']['frac'],'planar',d['planar']['ms_per_step'],'detect',d['detect']['value'],d['detect']['ms_per_volume'],d['detect']['e2e'])"; tail -2 gpurun_out/bench3.err
