#!/bin/bash
# A/B of library builds on the proof / bn254 workloads: swaps the in-package library on the GPU box (a scratch copy of the repo)
for v in "$@"; do
  cp tools/ubench/lib_$v.so bbs_sign_b200/libbbs_b200.so
  echo "== $v"
  python bench.py --workload proof --n 131072 --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('proof', d['kernels_ms'])"
  python bench.py --workload proof --n 32768 --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('proof32k', d['kernels_ms'])"
  python bench.py --workload bn254 --n 131072 --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bn254', d['kernels_ms'])"
done
