#!/bin/bash
# Round-2 ncu evidence, one gpurun call (one GPU). Every program first runs to exit 0 WITHOUT ncu.
#   1. launch list (gpu__time_duration.sum) of a 2-step configs[1] request  -> gpurun_out/r02_launches.csv
#   2. ncu --set full of the four DiT GEMM kinds + joint attention at M = 1920 and M = 640 -> summaries
set -u
mkdir -p gpurun_out
export ECHO_PROFILE_STEPS=2
python tools/profile_step.py > gpurun_out/ncu_plain.log 2>&1 || { echo "profile_step failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv \
    python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_summary.txt 2>&1
gzip -f gpurun_out/r02_launches.csv
for M in 1920 640; do
  M=$M python tools/ncu_gemm.py > gpurun_out/ncu_gemm_plain_$M.log 2>&1 || { echo "ncu_gemm M=$M failed"; exit 1; }
  M=$M ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|attn_tc" -s 10 -c 5 -f -o gpurun_out/r02_m$M \
      python tools/ncu_gemm.py > gpurun_out/ncu_gemm_$M.log 2>&1
  python tools/ncu_summary.py gpurun_out/r02_m$M.ncu-rep qkvg w13 wo w2 attention > gpurun_out/r02_ncu_m$M.txt 2>&1
done
python tools/ncu_traffic.py gpurun_out/r02_m1920.ncu-rep gpurun_out/r02_m640.ncu-rep "$(cat .git_head 2>/dev/null || echo r02)" > gpurun_out/r02_ncu_traffic.json 2>&1
ls -la gpurun_out/*.ncu-rep
tail -20 gpurun_out/r02_launches_summary.txt
cat gpurun_out/r02_ncu_m640.txt | head -12
