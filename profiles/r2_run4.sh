mkdir -p gpurun_out
python profiles/probes/py_setup_probe.py > gpurun_out/r2d_py_probe_notorch.txt 2>&1; cat gpurun_out/r2d_py_probe_notorch.txt | grep -E "rep|modules"
python profiles/probes/py_setup_probe.py --torch > gpurun_out/r2d_py_probe_torch.txt 2>&1; cat gpurun_out/r2d_py_probe_torch.txt | grep -E "rep|modules"
timeout 900 python -m pytest tests/test_device_setup.py -m gpu -q -k "stencil or hierarchy or iterations" > gpurun_out/r2d_test_stencil.log 2>&1; echo "stencil tests rc=$?"; tail -3 gpurun_out/r2d_test_stencil.log | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r2d_bench.err
timeout 600 python bench.py --config c2 --steps 10 --warmup 3 --skip-direct > gpurun_out/r2d_bench_c2.json 2> gpurun_out/r2d_bench_c2.err; echo "bench c2 rc=$?"
python - <<'PY'
import json
for f in ("r2d_bench", "r2d_bench_c2"):
    try:
        l = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(l["value"], 1), "ms/step", round(l["ms_per_step"], 2), "e2e", l["e2e"] and round(l["e2e"]["value"], 1), "roof", l["roofline"] and round(l["roofline"]["frac"], 3),
              "setup", l["setup"]["create_s"], l["setup"]["create_first_in_process_s"], "spmv", l.get("spmv_1e7") and {k: (round(v["ms"], 4), round(v["frac"], 3), round(v["actual_frac"], 3)) for k, v in l["spmv_1e7"].items() if isinstance(v, dict)},
              "parity", l.get("parity") and l["parity"]["max_rel_dev_of_R"], "pcg_iter", l["detail"].get("pcg_iter_k8_ms"), l["detail"].get("pcg_iter_k1_ms"), "cpu", l.get("cpu_baseline") and l["cpu_baseline"]["value"])
    except Exception as e:
        print(f, "ERR", e)
PY
# ncu: launch list of one solve (plain launches) + full sets of the top kernels, exported to CSV on the box
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 1 --warmup 1 --pairs 16 --skip-cpu --skip-e2e --skip-spmv1e7 --skip-direct --loop plain > gpurun_out/r2d_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"k_stencil|k_cg_update|k_spmm_win" -c 60 -o /tmp/r2d_full python profiles/run_profile.py --rows 3163 --what cg8 --precond amg > gpurun_out/r2d_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/r2d_full.ncu-rep --page raw --csv > gpurun_out/r2d_full_raw.csv 2>/dev/null; ls -la gpurun_out/r2d_full_raw.csv /tmp/r2d_full.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:"k_stencil" -c 6 -o /tmp/r2d_spmm python profiles/run_profile.py --rows 3163 --what spmm --reps 1 > gpurun_out/r2d_ncu_spmm.log 2>&1
ncu -i /tmp/r2d_spmm.ncu-rep --page raw --csv > gpurun_out/r2d_spmm_raw.csv 2>/dev/null; ls -la gpurun_out/r2d_spmm_raw.csv
