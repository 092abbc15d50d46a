#!/bin/bash
# usage (on the GPU box):  bash scripts/validate_gpu1.sh TAG
# One-GPU round-end check: GPU tests, smoke, the default bench line, then (only after those exited 0,
# and never as a bench value) the ncu launch list of a short bench run.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-check}
set -o pipefail
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 || exit 1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" || exit 2
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || { tail -20 gpurun_out/bench_$TAG.err; exit 3; }
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print('bench', round(d['value'], 1), d['unit'], 'ms/step', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 4),
      'frame frac', round(d['roofline']['frame']['frac'], 4), 'e2e', round(d['e2e']['value'], 1), 'cpu', round(d['cpu_baseline']['value'], 1))
PY
timeout 300 python bench.py --workload flat1m --no-cpu --no-e2e > gpurun_out/bench_${TAG}_flat1m.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"
# one ncu --set full capture each of the dominant kernel (8th scene_level launch = level 7) and of emit
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scene_level_kernel --launch-skip 7 --launch-count 1 \
  -o gpurun_out/prof_${TAG}_level7 -f python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu level7 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:emit_visible_kernel --launch-skip 1 --launch-count 1 \
  -o gpurun_out/prof_${TAG}_emit -f python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e >> gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu emit rc=$?"
