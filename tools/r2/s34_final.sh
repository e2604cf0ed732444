#!/bin/bash
# round 2, final validation on one B200: what the driver runs at round end (GPU tests, smoke, bench both arms) on the final commit
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s34.log) 2>&1
echo "=== pytest gpu (all)"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== reference arm"; timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 2>/dev/null | tail -1 | cut -c1-700
echo "=== bench (default flags)"; S=$(date +%s); timeout 1200 python bench.py 2>/dev/null | tail -1 > gpurun_out/r2_s34_bench_default.json; echo "wall $(( $(date +%s) - S )) s"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_s34_bench_default.json').read())
print({k:(v if not isinstance(v,dict) else {a:b for a,b in v.items() if a!='per_kernel'}) for k,v in d.items() if k!='config'})
PY
echo "=== done"
