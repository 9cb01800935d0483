mkdir -p gpurun_out
./profiles/probes/alloc_probe > gpurun_out/r2b_alloc_probe.txt 2>&1; cat gpurun_out/r2b_alloc_probe.txt
timeout 900 python -m pytest tests/test_device_setup.py tests/test_reference_kats.py -m gpu -q > gpurun_out/r2b_test_setup.log 2>&1; echo "setup+kat tests rc=$?"; tail -25 gpurun_out/r2b_test_setup.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_raster_assembly.py -m gpu -q > gpurun_out/r2b_test_parity.log 2>&1; echo "parity tests rc=$?"; tail -6 gpurun_out/r2b_test_parity.log | cut -c1-300
timeout 1500 python -m pytest tests/test_gpu_scale.py -m gpu -q --durations=5 > gpurun_out/r2b_test_scale.log 2>&1; echo "scale tests rc=$?"; tail -14 gpurun_out/r2b_test_scale.log | cut -c1-300
CS_B200_VERBOSE=1 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2b_bench.json; grep "cs_b200 setup" gpurun_out/r2b_bench.err | head -90
CS_B200_NO_STENCIL=1 timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-direct --skip-e2e > gpurun_out/r2b_bench_nostencil.json 2> gpurun_out/r2b_bench_nostencil.err; echo "bench nostencil rc=$?"
CS_B200_VERBOSE=1 timeout 600 python bench.py --config c2 --steps 5 --warmup 3 --skip-cpu --skip-direct --skip-spmv1e7 > gpurun_out/r2b_bench_c2.json 2> gpurun_out/r2b_bench_c2.err; echo "bench c2 rc=$?"
tail -c 300 gpurun_out/r2b_bench_c2.json; grep "cs_b200 setup\]" gpurun_out/r2b_bench_c2.err | head -20
python - <<'PY'
import json
for f in ("r2b_bench", "r2b_bench_nostencil", "r2b_bench_c2"):
    try:
        l = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(l["value"], 1), "ms/step", round(l["ms_per_step"], 2), "roof", l["roofline"] and round(l["roofline"]["frac"], 3),
              "setup", l["setup"]["create_s"], "spmv", l.get("spmv_1e7") and {k: round(v["frac"], 3) for k, v in l["spmv_1e7"].items() if isinstance(v, dict)},
              "parity", l.get("parity") and l["parity"]["max_rel_dev_of_R"], "pcg_iter", l["detail"].get("pcg_iter_k8_ms"))
    except Exception as e:
        print(f, "ERR", e)
PY
