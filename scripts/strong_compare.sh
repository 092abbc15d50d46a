#!/bin/bash
# usage (on the GPU box):  bash scripts/strong_compare.sh N exchange [exchange ...]
# BASELINE configs[4] (8 x 33.5M-entity scenes over N GPUs) once per --exchange mode.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
N=$1; shift
for ex in "$@"; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 \
    bench.py --gpus $N --workload scenes8x32m --steps 20 --warmup 3 --exchange $ex > gpurun_out/z_${ex}_$N.json 2> gpurun_out/z_${ex}_$N.err
  echo "rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/z_${ex}_$N.json').read().strip().splitlines()[-1])
    print('strong', '$ex', d['n_gpus'], round(d['value'], 1), round(d['ms_per_step'], 4), d['config']['exchange'])
except Exception as e:
    print('$ex failed', e)
PY
  tail -3 gpurun_out/z_${ex}_$N.err
done
