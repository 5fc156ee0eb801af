#!/bin/bash
# One GPU visit (round 2): reference noise floor, ncu launch list of a whole pass, ncu --set full of the encoder kernels.
# Usage (from the repo root, under gpurun): bash tools/gpu_round2.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python tools/reference_noise_floor.py > $OUT/${TAG}_reference_noise_floor.json 2> $OUT/${TAG}_noise_floor.err; echo "noise floor rc=$?"
cat $OUT/${TAG}_reference_noise_floor.json | head -60
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/${TAG}_launches.csv python tools/profile_step.py > $OUT/${TAG}_ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_gemm -c 53 -f \
    -o $OUT/${TAG}_full_conv_gemm python tools/profile_resnet.py > $OUT/${TAG}_ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:linear_umma -c 8 -f \
    -o $OUT/${TAG}_full_linear_umma python tools/profile_step.py > $OUT/${TAG}_ncu_linear.log 2>&1; echo "ncu linear rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gcn_hidden_umma -s 8 -c 2 -f \
    -o $OUT/${TAG}_full_gcn_hidden python tools/profile_step.py > $OUT/${TAG}_ncu_gcn.log 2>&1; echo "ncu gcn rc=$?"
ls -la $OUT | tail -12
