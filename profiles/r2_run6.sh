mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_device_setup.py tests/test_gpu_parity.py tests/test_raster_assembly.py tests/test_out_files.py tests/test_reference_kats.py -m gpu -q -x > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2f_tests.log | cut -c1-400
B="python bench.py --steps 3 --warmup 2 --skip-cpu --skip-e2e --skip-direct --skip-spmv1e7"
for v in default nofuse rplain nofuse_pplain nofuse_prplain; do
  case $v in
    default) E="";;
    nofuse) E="CS_B200_NO_FUSED_PROLONG=1";;
    rplain) E="CS_B200_WIN_MASK=7";;
    nofuse_pplain) E="CS_B200_NO_FUSED_PROLONG=1 CS_B200_WIN_MASK=11";;
    nofuse_prplain) E="CS_B200_NO_FUSED_PROLONG=1 CS_B200_WIN_MASK=3";;
  esac
  env $E timeout 600 $B > gpurun_out/r2f_bench_$v.json 2> gpurun_out/r2f_bench_$v.err; echo "bench $v rc=$?"
done
python - <<'PY'
import json
for v in ("default", "nofuse", "rplain", "nofuse_pplain", "nofuse_prplain"):
    try:
        l = json.loads(open(f"gpurun_out/r2f_bench_{v}.json").read().strip().splitlines()[-1])
        print(v, "value", round(l["value"], 1), "ms/step", round(l["ms_per_step"], 2), "pcg_iter k8/k1", round(l["detail"]["pcg_iter_k8_ms"], 3), round(l["detail"]["pcg_iter_k1_ms"], 3), "roof", round(l["roofline"]["frac"], 3), "iters", l["detail"]["iterations_rank0"][:8], "R0", l["detail"]["R_first"][0])
    except Exception as e:
        print(v, "ERR", e)
PY
