python tools/ncu_gemm.py > gpurun_out/ncu_plain.log 2>&1 && ncu --set full --import-source on --clock-control none -k regex:gemm_tc -s 8 -c 4 -o gpurun_out/prof_gemm_v6 python tools/ncu_gemm.py > gpurun_out/ncu6.log 2>&1
tail -5 gpurun_out/ncu6.log; ls -la gpurun_out/*.ncu-rep
