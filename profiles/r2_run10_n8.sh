mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 --skip-spmv1e7 > gpurun_out/r2j_bench_n8.json 2> gpurun_out/r2j_bench_n8.err; echo "bench n8 rc=$?"; tail -3 gpurun_out/r2j_bench_n8.err | cut -c1-300
timeout 900 $TR --master-port 29532 bench.py --gpus 8 --rows 4000 --cols 4000 --pairs 1000 --steps 2 --warmup 1 --skip-e2e --skip-spmv1e7 > gpurun_out/r2j_bench_c4_n8.json 2> gpurun_out/r2j_bench_c4_n8.err; echo "bench C4 n8 rc=$?"; tail -3 gpurun_out/r2j_bench_c4_n8.err | cut -c1-300
timeout 600 $TR --master-port 29533 profiles/run_network.py --device-resident --reps 3 --check 2 > gpurun_out/r2j_network_c5_n8.json 2> gpurun_out/r2j_network_c5_n8.err; echo "C5 n8 rc=$?"; tail -3 gpurun_out/r2j_network_c5_n8.err | cut -c1-300
python - <<'PY'
import json
for f in ("r2j_bench_n8", "r2j_bench_c4_n8"):
    try:
        l = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(l["value"], 1), "ms/step", round(l["ms_per_step"], 2), "e2e", l["e2e"] and round(l["e2e"]["value"], 1), "setup", {k: v for k, v in l["setup"].items() if k != "note"}, "iters", l["detail"]["iterations_per_rank_sum_max_count"])
    except Exception as e:
        print(f, "ERR", e)
try:
    print(open("gpurun_out/r2j_network_c5_n8.json").read().strip().splitlines()[-1][:900])
except Exception as e:
    print("C5 ERR", e)
PY
