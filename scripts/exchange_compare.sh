#!/bin/bash
# usage (on the GPU box):  bash scripts/exchange_compare.sh N exchange [exchange ...]
# Runs bench.py at N GPUs once per --exchange mode; JSON lines land in gpurun_out/x_<mode>_<N>.json.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
N=$1; shift
for ex in "$@"; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps 50 --warmup 5 --exchange $ex --no-cpu --no-e2e > gpurun_out/x_${ex}_$N.json 2> gpurun_out/x_${ex}_$N.err
  echo "rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/x_${ex}_$N.json').read().strip().splitlines()[-1])
    print('$ex', d['n_gpus'], round(d['value'], 1), round(d['ms_per_step'], 4))
except Exception as e:
    print('$ex failed', e)
PY
done
