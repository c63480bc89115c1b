#!/bin/bash
# ncu evidence for profiles/ (B200_PROFILING.md recipe); run on the GPU box from the repo root AFTER the plain commands
# have exited 0:   bash tools/profile_round.sh r02
# 1. launch list of the default bench command (cold-cache, serialised: compare shares, not absolutes)
# 2. one `--set full` capture per hot kernel (source-mapped: the library is built with -lineinfo)
set -u
TAG=${1:-r02}
OUT=gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/plain_$TAG.log 2>&1 || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches_${TAG}_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_launches_$TAG.log 2>&1
full() {  # name, kernel regex, bench arguments...
  local name=$1 rx=$2; shift 2
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$rx" -c 1 -f -o $OUT/prof_${name}_$TAG \
      python bench.py --steps 1 --warmup 3 --no-cpu "$@" > $OUT/ncu_${name}_$TAG.log 2>&1
  ncu -i $OUT/prof_${name}_$TAG.ncu-rep --page raw --csv > $OUT/raw_${name}_$TAG.csv 2>/dev/null
}
full coop 'pairing_coop_kernel' --extras ''
full verify_g1 'verify_g1_split_kernel' --extras ''
full proof_g1 "proof_g1_split_kernel" --workload proof --n 65536
full sign 'sign_kernel' --workload sign --n 524288
full rlc_prep 'rlc_prep_kernel' --workload rlc --n 524288
ls -la $OUT | tail -20
