set -u
OUT=gpurun_out
TAG=${1:-r02b}
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:pairing_coop_kernel" -c 1 -f -o $OUT/prof_coop_$TAG python bench.py --steps 1 --warmup 3 --no-cpu --extras '' > $OUT/ncu_coop_$TAG.log 2>&1
ncu -i $OUT/prof_coop_$TAG.ncu-rep --page raw --csv > $OUT/raw_coop_$TAG.csv 2>/dev/null
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:verify_g1_kernel" -c 1 -f -o $OUT/prof_verify_g1_$TAG python bench.py --steps 1 --warmup 3 --no-cpu --extras '' > $OUT/ncu_verify_g1_$TAG.log 2>&1
ncu -i $OUT/prof_verify_g1_$TAG.ncu-rep --page raw --csv > $OUT/raw_verify_g1_$TAG.csv 2>/dev/null
tail -2 $OUT/ncu_coop_$TAG.log
