mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2n_smoke.log | cut -c1-300
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2n_bench.err | cut -c1-300
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r2n_bench.json").read().strip().splitlines()[-1])
print("value", round(l["value"], 1), "ms/step", round(l["ms_per_step"], 2), "e2e", round(l["e2e"]["value"], 1), "roof", round(l["roofline"]["frac"], 3), "parity", l["parity"]["max_rel_dev_of_R"], "cpu", l["cpu_baseline"]["value"], "setup", {k: v for k, v in l["setup"].items() if k != "note"})
PY
