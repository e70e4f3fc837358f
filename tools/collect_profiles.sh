#!/bin/bash
# Round artefacts: full bench line, ncu launch lists (bench command, training step) and one --set full capture of
# the hash-encode kernels.  Every ncu pass runs only after the same command exited 0 without ncu.
# usage: tools/collect_profiles.sh <tag>      (run from the repo root on the GPU box)
set -u
TAG=${1:-r1d}
OUT=gpurun_out
mkdir -p $OUT
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err || { echo "bench failed"; tail -5 $OUT/bench_$TAG.err; }
CMD="python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline"
if $CMD > $OUT/plain_$TAG.log 2>&1; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_$TAG.log 2>&1
fi
export N_RAND=8192 STEPS=1 WARMUP=1
if python tools/prof_train.py > $OUT/plain_train_$TAG.log 2>&1; then
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
      --clock-control none -c 400 --csv --log-file $OUT/launches_train8192_$TAG.csv python tools/prof_train.py > $OUT/ncu_train_$TAG.log 2>&1
fi
if python tools/prof_hash.py > $OUT/plain_hash_$TAG.log 2>&1; then
  ncu --set full --clock-control none --import-source on -k regex:"hash_fwd|hash_bwd|sort" -c 12 -o $OUT/prof_hash_$TAG -f python tools/prof_hash.py > $OUT/ncu_hash_$TAG.log 2>&1
fi
if python tools/prof_mlp.py > $OUT/plain_mlp_$TAG.log 2>&1; then
  ncu --set full --clock-control none --import-source on -k regex:"mlp_tc" -s 3 -c 3 -o $OUT/prof_mlp_$TAG -f python tools/prof_mlp.py > $OUT/ncu_mlp_$TAG.log 2>&1
fi
python tools/measure_tensor_peak.py > $OUT/tensor_peaks_$TAG.json 2>&1
python tools/time_mlp.py > $OUT/time_mlp_$TAG.log 2>&1
ls -la $OUT | tail -20
