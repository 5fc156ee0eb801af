#!/bin/bash
# One GPU visit: parity tests, bench (graph + eager), ncu launch list, ncu --set full of the HBM-bound kernels.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q -s > $OUT/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_gpu_tests.log
tail -3 $OUT/${TAG}_gpu_tests.log
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
cat $OUT/${TAG}_bench.json | cut -c1-1500
tail -5 $OUT/${TAG}_bench.err
timeout 600 python bench.py --no-graph --steps 10 --warmup 3 --no-cpu-baseline --no-reference-cuda > $OUT/${TAG}_bench_eager.json 2>> $OUT/${TAG}_bench.err
if [ "$2" != "noncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/${TAG}_launches.csv python tools/profile_step.py > $OUT/${TAG}_ncu_launches.log 2>&1
for k in gcn_input gcn_output smpl_skin; do
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -c 1 -f \
      -o $OUT/${TAG}_full_$k python tools/profile_step.py > $OUT/${TAG}_ncu_$k.log 2>&1
done
fi
ls -la $OUT | tail -20
