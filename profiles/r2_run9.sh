mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/r2i_tests_all.log 2>&1; echo "all gpu tests rc=$?"; tail -8 gpurun_out/r2i_tests_all.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1; echo "smoke rc=$?"; cat gpurun_out/r2i_smoke.log | tail -3
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/r2i_bench.err
( time timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2i_bench_ref.json 2> gpurun_out/r2i_bench_ref.err; echo "reference rc=$?"; tail -4 gpurun_out/r2i_bench_ref.err
timeout 600 python bench.py --config c2 --steps 20 --warmup 5 > gpurun_out/r2i_bench_c2.json 2> gpurun_out/r2i_bench_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json
for f in ("r2i_bench", "r2i_bench_ref", "r2i_bench_c2"):
    try:
        l = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(l["value"], 2), "ms/step", round(l["ms_per_step"], 2), "steps", l["steps"], "e2e", l["e2e"] and round(l["e2e"]["value"], 2),
              "roof", l.get("roofline") and round(l["roofline"]["frac"], 3), "setup", l.get("setup"), "parity", l.get("parity") and l["parity"]["max_rel_dev_of_R"],
              "cpu", l.get("cpu_baseline") and (l["cpu_baseline"]["value"], l["cpu_baseline"]["cores"]))
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r2i_launches.csv python bench.py --steps 1 --warmup 1 --pairs 16 --skip-cpu --skip-e2e --skip-spmv1e7 --skip-direct --loop plain > gpurun_out/r2i_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"k_stencil|k_cg_update|k_spmm_win" -c 40 -o /tmp/r2i_full python profiles/run_profile.py --rows 3163 --what cg8 --precond amg --reps 1 > gpurun_out/r2i_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/r2i_full.ncu-rep --page raw --csv > gpurun_out/r2i_full_raw.csv 2>/dev/null; ls -la gpurun_out/r2i_full_raw.csv
timeout 600 ncu --set full --clock-control none -k regex:"k_stencil" -c 6 -o /tmp/r2i_spmm python profiles/run_profile.py --rows 3163 --what spmm --reps 1 > gpurun_out/r2i_ncu_spmm.log 2>&1
ncu -i /tmp/r2i_spmm.ncu-rep --page raw --csv > gpurun_out/r2i_spmm_raw.csv 2>/dev/null; ls -la gpurun_out/r2i_spmm_raw.csv
