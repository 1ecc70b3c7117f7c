mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/ngpu.txt
: > gpurun_out/scale_bench.jsonl; : > gpurun_out/scale_blockwise.jsonl; : > gpurun_out/scale.err
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 200 --warmup 20 --skip-cpu --skip-detect >> gpurun_out/scale_bench.jsonl 2>> gpurun_out/scale.err
    timeout 600 python tools/blockwise_bench.py both 1.0 >> gpurun_out/scale_blockwise.jsonl 2>> gpurun_out/scale.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 200 --warmup 20 --skip-cpu --skip-detect >> gpurun_out/scale_bench.jsonl 2>> gpurun_out/scale.err
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) tools/blockwise_bench.py both 1.0 >> gpurun_out/scale_blockwise.jsonl 2>> gpurun_out/scale.err
  fi
done
cat gpurun_out/ngpu.txt; python -c "
import json
for l in open('gpurun_out/scale_bench.jsonl'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('bench N',d['n_gpus'],'value %.4g'%d['value'],'ms',round(d['ms_per_step'],5),'e2e %.4g'%d['e2e']['value'])
"; cat gpurun_out/scale_blockwise.jsonl; tail -3 gpurun_out/scale.err
