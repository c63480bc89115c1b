#!/bin/bash
# usage: tools/bench_variants.sh lib1.so lib2.so ...   (tuning helper: same bench, different CUDA builds)
for lib in "$@"; do
  BBS_B200_LIB=$PWD/$lib python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', round(d['value']), d['kernels_ms'])" || echo "$lib FAILED"
done
