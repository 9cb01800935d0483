mkdir -p gpurun_out
CS_B200_VERBOSE=2 ./profiles/probes/setup_probe 3163 3163 2 > gpurun_out/r2c_setup_probe_3163.txt 2>&1; echo "probe rc=$?"; grep -v "setup/device\] L[1-9]" gpurun_out/r2c_setup_probe_3163.txt | head -150
CS_B200_VERBOSE=1 ./profiles/probes/setup_probe 1000 1000 2 > gpurun_out/r2c_setup_probe_1000.txt 2>&1; grep -E "rep|n=" gpurun_out/r2c_setup_probe_1000.txt
timeout 600 ncu --set full --clock-control none -k regex:k_stencil -c 8 -o gpurun_out/r2c_stencil_spmm python profiles/run_profile.py --rows 3163 --what spmm --reps 2 > gpurun_out/r2c_ncu1.log 2>&1; echo "ncu1 rc=$?"; tail -3 gpurun_out/r2c_ncu1.log
timeout 600 ncu --set full --clock-control none -k regex:k_stencil -c 14 -o gpurun_out/r2c_stencil_cg python profiles/run_profile.py --rows 3163 --what cg --precond amg --reps 1 > gpurun_out/r2c_ncu2.log 2>&1; echo "ncu2 rc=$?"; tail -3 gpurun_out/r2c_ncu2.log
CS_B200_NO_STENCIL=1 timeout 600 ncu --set full --clock-control none -k regex:k_spmm_win -c 24 -o gpurun_out/r2c_win_cg python profiles/run_profile.py --rows 3163 --what cg --precond amg --reps 1 > gpurun_out/r2c_ncu3.log 2>&1; echo "ncu3 rc=$?"; tail -3 gpurun_out/r2c_ncu3.log
ls -la gpurun_out/*.ncu-rep
