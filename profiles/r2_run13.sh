mkdir -p gpurun_out
B="python bench.py --pairs 32 --steps 3 --warmup 2 --skip-cpu --skip-e2e --skip-direct"
for v in pjocc4 allocc4; do
  case $v in
    pjocc4) E="";;
    allocc4) E="CS_B200_STENCIL_OCC4=1";;
  esac
  env $E timeout 400 $B > gpurun_out/r2m_bench_$v.json 2> gpurun_out/r2m_bench_$v.err; echo "bench $v rc=$?"
done
BEST=$(python - <<'PY'
import json
def val(v):
    try:
        return json.loads(open(f"gpurun_out/r2m_bench_{v}.json").read().strip().splitlines()[-1])["value"]
    except Exception:
        return 0.0
a, b = val("pjocc4"), val("allocc4")
print("CS_B200_STENCIL_OCC4=1" if b > a * 1.005 else "CS_B200_STENCIL_OCC4=0")
PY
)
echo "best: $BEST"
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2m_tests_default.log 2>&1; echo "tests(default) rc=$?"; tail -3 gpurun_out/r2m_tests_default.log | cut -c1-300
env CS_B200_STENCIL_OCC4=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_device_setup.py -m gpu -q -x > gpurun_out/r2m_tests_occ4.log 2>&1; echo "tests(stencil occ4) rc=$?"; tail -3 gpurun_out/r2m_tests_occ4.log | cut -c1-300
env $BEST timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2m_smoke.log | cut -c1-300
env $BEST timeout 900 python bench.py --steps 5 --warmup 3 --skip-cpu --skip-direct > gpurun_out/r2m_bench_headline.json 2> gpurun_out/r2m_bench_headline.err; echo "headline rc=$?"
env $BEST timeout 600 python bench.py --config c2 --steps 10 --warmup 3 --skip-cpu --skip-direct --skip-spmv1e7 > gpurun_out/r2m_bench_c2.json 2> gpurun_out/r2m_bench_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json
for v in ("pjocc4", "allocc4", "headline", "c2"):
    try:
        l = json.loads(open(f"gpurun_out/r2m_bench_{v}.json").read().strip().splitlines()[-1])
        sp = l.get("spmv_1e7") or {}
        print(v, "value", round(l["value"], 1), "ms/step", round(l["ms_per_step"], 2), "e2e", l["e2e"] and round(l["e2e"]["value"], 1), "pcg_iter k8/k1", round(l["detail"]["pcg_iter_k8_ms"], 3), round(l["detail"]["pcg_iter_k1_ms"], 3), "roof", round(l["roofline"]["frac"], 3), "spmv k1/k8", sp and (round(sp["k1"]["ms"], 4), round(sp["k8"]["ms"], 4)), "R0", l["detail"]["R_first"][0])
        print("    ", {k: (v_["launches"], round(v_["avg_ms"], 4), round(v_["frac"], 3)) for k, v_ in l["roofline"]["by_kernel"].items()})
    except Exception as e:
        print(v, "ERR", e)
PY
