#!/bin/bash
# bench.py at N = 1 .. $1 GPUs of one box (run under gpurun --gpus N); prints the scaling table
MAXN=${1:-8}
STEPS=${2:-20}
mkdir -p gpurun_out
for n in ${NLIST:-1 2 4 8}; do
  [ $n -gt $MAXN ] && break
  if [ $n -eq 1 ]; then
    python bench.py --gpus 1 --steps $STEPS --warmup 5 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) \
       bench.py --gpus $n --steps $STEPS --warmup 5 --no-cpu-baseline $BENCH_EXTRA > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  fi
  tail -c 600 gpurun_out/scale_$n.err
done
python - <<'PY'
import json, glob
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open("gpurun_out/scale_%d.json" % n).read().strip().splitlines()[-1])
    except Exception as e:
        continue
    if base is None:
        base = d["value"]
    print("N=%d value %.4g evals/s  %.3f ms/step  eff %.3f  e2e %.4g (%.3f ms)  allgather %.3f ms  check=%s" % (
        d["n_gpus"], d["value"], d["ms_per_step"], d["value"] / base / d["n_gpus"], d["e2e"]["value"],
        d["e2e"]["ms_per_step"], d["allgather_loglik"]["ms_per_step"], d["parity_check"]))
PY
