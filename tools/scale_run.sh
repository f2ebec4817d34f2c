#!/bin/bash
# 1/2/4/8-GPU scaling lines of bench.py on one box (run under gpurun --gpus 8)
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) \
     bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
done
for n in 1 2 4 8; do tail -1 gpurun_out/scale_$n.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['sampler_exchange']['ms_per_step'])"; done
