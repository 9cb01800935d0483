mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_device_setup.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2h_tests.log | cut -c1-400
B="python bench.py --steps 3 --warmup 2 --skip-cpu --skip-e2e --skip-direct --skip-spmv1e7"
for v in default noimplicit; do
  case $v in
    default) E="";;
    noimplicit) E="CS_B200_NO_IMPLICIT_X0=1";;
  esac
  env $E timeout 600 $B > gpurun_out/r2h_bench_$v.json 2> gpurun_out/r2h_bench_$v.err; echo "bench $v rc=$?"
done
env timeout 600 python bench.py --config c2 --steps 10 --warmup 3 --skip-cpu --skip-direct --skip-spmv1e7 > gpurun_out/r2h_bench_c2.json 2> gpurun_out/r2h_bench_c2.err
python - <<'PY'
import json
for v in ("default", "noimplicit", "c2"):
    try:
        l = json.loads(open(f"gpurun_out/r2h_bench_{v}.json").read().strip().splitlines()[-1])
        print(v, "value", round(l["value"], 1), "ms/step", round(l["ms_per_step"], 2), "pcg_iter k8/k1", round(l["detail"]["pcg_iter_k8_ms"], 3), round(l["detail"]["pcg_iter_k1_ms"], 3), "roof", round(l["roofline"]["frac"], 3), "iters", l["detail"]["iterations_rank0"][:8], "R0", l["detail"]["R_first"][0])
        print("    ", {k: (v_["launches"], round(v_["avg_ms"], 4), round(v_["frac"], 3)) for k, v_ in l["roofline"]["by_kernel"].items()})
    except Exception as e:
        print(v, "ERR", e)
PY
