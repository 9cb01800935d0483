mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/r2e_test_multigpu.log 2>&1; echo "multi-gpu test rc=$?"; tail -25 gpurun_out/r2e_test_multigpu.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 2 --skip-spmv1e7 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/r2e_bench_n2.err | cut -c1-300
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/r2e_bench_n2.json").read().strip().splitlines()[-1])
    print("N=2 value", round(l["value"], 1), "ms/step", round(l["ms_per_step"], 2), "e2e", l["e2e"] and round(l["e2e"]["value"], 1), "setup", l["setup"], "iters", l["detail"]["iterations_per_rank_sum_max_count"])
except Exception as e:
    print("ERR", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 --cpu-sample 4 --rows 600 --cols 600 --pairs 8 > gpurun_out/r2e_bench_ref_n2.json 2> gpurun_out/r2e_bench_ref_n2.err; echo "reference arm under torchrun rc=$?"; tail -c 300 gpurun_out/r2e_bench_ref_n2.json
