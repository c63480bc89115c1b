#!/bin/bash
# usage: tools/bench_variants_bn.sh lib1.so ...   (tuning helper: BN254 config-3 workload, different CUDA builds)
for lib in "$@"; do
  BBS_B200_LIB=$PWD/$lib python bench.py --workload bn254 --n 131072 --steps 2 --warmup 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', round(d['value']), d['kernels_ms'])" || echo "$lib FAILED"
done
