mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 600 python -m pytest tests/test_device_setup.py -m gpu -x -q > gpurun_out/r2a_test_setup.log 2>&1; echo "setup tests rc=$?"
tail -15 gpurun_out/r2a_test_setup.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_device_setup.py > gpurun_out/r2a_test_all.log 2>&1; echo "all gpu tests rc=$?"
tail -8 gpurun_out/r2a_test_all.log
CS_B200_VERBOSE=1 timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2a_bench.json; grep "cs_b200 setup" gpurun_out/r2a_bench.err | head -60
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 1 --warmup 1 --pairs 16 --skip-cpu --skip-e2e --skip-spmv1e7 --skip-direct --loop plain > gpurun_out/r2a_ncu_bench.log 2>&1; echo "ncu rc=$?"
