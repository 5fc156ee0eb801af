#!/bin/bash
# GPU visit after the transposed K1: full GPU test suite, bench (both arms' own line only), launch list, ncu --set full of K1.
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -x -q -m gpu -s > $OUT/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/${TAG}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_1gpu.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-600 $OUT/${TAG}_bench_1gpu.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/${TAG}_launches.csv python tools/profile_step.py > $OUT/${TAG}_ncu_launches.log 2>&1; echo "launch list rc=$?"
python tools/ncu_summary.py shares $OUT/${TAG}_launches.csv > $OUT/${TAG}_launch_shares.txt; head -12 $OUT/${TAG}_launch_shares.txt
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gcn_hidden_umma -s 8 -c 2 -f \
    -o $OUT/${TAG}_full_gcn_hidden python tools/profile_step.py > $OUT/${TAG}_ncu_gcn.log 2>&1; echo "ncu gcn rc=$?"
python tools/ncu_summary.py full $OUT/${TAG}_full_gcn_hidden.ncu-rep > $OUT/${TAG}_ncu_full_gcn_hidden_umma.json
ls -la $OUT | tail -12
