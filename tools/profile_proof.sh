set -u
OUT=gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:proof_g1_item" -c 1 -f -o $OUT/prof_proof_g1_r02 python bench.py --steps 1 --warmup 3 --no-cpu --workload proof --n 65536 > $OUT/ncu_proof_g1_r02.log 2>&1
ncu -i $OUT/prof_proof_g1_r02.ncu-rep --page raw --csv > $OUT/raw_proof_g1_r02.csv 2>/dev/null
tail -3 $OUT/ncu_proof_g1_r02.log
