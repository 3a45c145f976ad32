#!/bin/bash
# A/B helper for gpurun: runs bench.py once per argument with `VAR=value` pairs exported, prints the headline numbers.
#   bash tools/abtest.sh "VTMGPU_E2E_DENSE_RECORDS=1" "VTMGPU_E2E_DENSE_RECORDS=0"
for v in "$@"; do
  echo "== $v"
  env $v python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-stress --e2e-steps ${E2E_STEPS:-3} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernel_ms'], d['e2e'])"
done
