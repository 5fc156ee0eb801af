#!/bin/bash
# One GPU visit (round 2): ncu launch list of a whole pass and of the image encoder, ncu --set full of the tensor-core
# kernels (summarised ON the box: the .ncu-rep files of 50+ launches exceed gpurun's 64 MiB return limit).
# Usage (from the repo root, under gpurun): bash tools/gpu_round2.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/${TAG}_launches.csv python tools/profile_step.py > $OUT/${TAG}_ncu_launches.log 2>&1; echo "launch list rc=$?"
python tools/ncu_summary.py shares $OUT/${TAG}_launches.csv > $OUT/${TAG}_launch_shares.txt; cat $OUT/${TAG}_launch_shares.txt | head -20
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
    --clock-control none --profile-from-start off --csv --log-file $OUT/${TAG}_resnet_launches_ncu.csv python tools/profile_resnet.py > /dev/null 2>&1; echo "resnet list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_gemm -c 53 -f \
    -o /tmp/${TAG}_full_conv_gemm python tools/profile_resnet.py > $OUT/${TAG}_ncu_conv.log 2>&1; echo "ncu conv rc=$?"
python tools/ncu_summary.py full /tmp/${TAG}_full_conv_gemm.ncu-rep > $OUT/${TAG}_ncu_full_conv_gemm.json
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:linear_umma -c 8 -f \
    -o /tmp/${TAG}_full_linear_umma python tools/profile_step.py > $OUT/${TAG}_ncu_linear.log 2>&1; echo "ncu linear rc=$?"
python tools/ncu_summary.py full /tmp/${TAG}_full_linear_umma.ncu-rep > $OUT/${TAG}_ncu_full_linear_umma.json
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gcn_hidden_umma -s 8 -c 2 -f \
    -o $OUT/${TAG}_full_gcn_hidden python tools/profile_step.py > $OUT/${TAG}_ncu_gcn.log 2>&1; echo "ncu gcn rc=$?"
python tools/ncu_summary.py full $OUT/${TAG}_full_gcn_hidden.ncu-rep > $OUT/${TAG}_ncu_full_gcn_hidden_umma.json
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gcn_input|gcn_output|smpl_skin_tiled" -c 3 -f \
    -o $OUT/${TAG}_full_simt python tools/profile_step.py > $OUT/${TAG}_ncu_simt.log 2>&1; echo "ncu simt rc=$?"
python tools/ncu_summary.py full $OUT/${TAG}_full_simt.ncu-rep > $OUT/${TAG}_ncu_full_simt_kernels.json
ls -la $OUT | tail -16
