mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --rows 4000 --cols 4000 --pairs 1000 --steps 1 --warmup 1 --skip-e2e --skip-spmv1e7 --skip-cpu --skip-direct > gpurun_out/r2k_bench_c4_n1.json 2> gpurun_out/r2k_bench_c4_n1.err; echo "C4 n1 rc=$?"; tail -3 gpurun_out/r2k_bench_c4_n1.err | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 --skip-cpu --skip-direct > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("r2k_bench_c4_n1", "r2k_bench"):
    try:
        l = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(l["value"], 1), "ms/step", round(l["ms_per_step"], 2), "e2e", l["e2e"] and round(l["e2e"]["value"], 1), "setup", {k: v for k, v in l["setup"].items() if k != "note"}, "traffic", l["roofline"]["traffic"], "frac", l["roofline"]["frac"])
    except Exception as e:
        print(f, "ERR", e)
PY
