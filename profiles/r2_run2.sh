mkdir -p gpurun_out
./profiles/probes/alloc_probe > gpurun_out/r2b_alloc_probe.txt 2>&1; cat gpurun_out/r2b_alloc_probe.txt
timeout 900 python -m pytest tests/test_device_setup.py tests/test_reference_kats.py -m gpu -q > gpurun_out/r2b_test_setup.log 2>&1; echo "setup+kat tests rc=$?"; tail -4 gpurun_out/r2b_test_setup.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_raster_assembly.py -m gpu -q > gpurun_out/r2b_test_parity.log 2>&1; echo "parity tests rc=$?"; tail -4 gpurun_out/r2b_test_parity.log
timeout 1500 python -m pytest tests/test_gpu_scale.py -m gpu -q --durations=5 > gpurun_out/r2b_test_scale.log 2>&1; echo "scale tests rc=$?"; tail -12 gpurun_out/r2b_test_scale.log
CS_B200_VERBOSE=1 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2b_bench.json; grep "cs_b200 setup" gpurun_out/r2b_bench.err | head -70
CS_B200_VERBOSE=1 timeout 600 python bench.py --config c2 --steps 5 --warmup 3 --skip-cpu --skip-direct --skip-spmv1e7 > gpurun_out/r2b_bench_c2.json 2> gpurun_out/r2b_bench_c2.err; echo "bench c2 rc=$?"
tail -c 300 gpurun_out/r2b_bench_c2.json; grep "cs_b200 setup\]" gpurun_out/r2b_bench_c2.err | head -20
