mkdir -p gpurun_out
B="python bench.py --pairs 32 --steps 3 --warmup 2 --skip-cpu --skip-e2e --skip-direct"
for v in base occ4 s1 s2 s2occ4; do
  case $v in
    base) E="";;
    occ4) E="CS_B200_PJ_OCC4=1";;
    s1) E="CS_B200_STENCIL_VARIANT=1";;
    s2) E="CS_B200_STENCIL_VARIANT=2";;
    s2occ4) E="CS_B200_STENCIL_VARIANT=2 CS_B200_PJ_OCC4=1";;
  esac
  env $E timeout 400 $B > gpurun_out/r2l_bench_$v.json 2> gpurun_out/r2l_bench_$v.err; echo "bench $v rc=$?"
done
env CS_B200_STENCIL_VARIANT=2 CS_B200_PJ_OCC4=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2l_tests_s2occ4.log 2>&1; echo "tests(s2occ4) rc=$?"; tail -3 gpurun_out/r2l_tests_s2occ4.log | cut -c1-300
env CS_B200_STENCIL_VARIANT=1 timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2l_smoke_s1.log 2>&1; echo "smoke(s1) rc=$?"; tail -2 gpurun_out/r2l_smoke_s1.log | cut -c1-300
python - <<'PY'
import json
for v in ("base", "occ4", "s1", "s2", "s2occ4"):
    try:
        l = json.loads(open(f"gpurun_out/r2l_bench_{v}.json").read().strip().splitlines()[-1])
        print(v, "value", round(l["value"], 1), "ms/step", round(l["ms_per_step"], 2), "pcg_iter k8/k1", round(l["detail"]["pcg_iter_k8_ms"], 3), round(l["detail"]["pcg_iter_k1_ms"], 3), "roof", round(l["roofline"]["frac"], 3), "spmv k1/k8", round(l["spmv_1e7"]["k1"]["ms"], 4), round(l["spmv_1e7"]["k8"]["ms"], 4), "iters", l["detail"]["iterations_rank0"][:6], "R0", l["detail"]["R_first"][0])
        print("    ", {k: (v_["launches"], round(v_["avg_ms"], 4), round(v_["frac"], 3)) for k, v_ in l["roofline"]["by_kernel"].items()})
    except Exception as e:
        print(v, "ERR", e)
PY
