timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_conv.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -3
timeout 500 python bench.py --steps 6 --warmup 3 > gpurun_out/r2l_rl.json 2> gpurun_out/r2l_rl.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2l_rl.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"])
r = d["roofline"]
print(r["frac"], r.get("frac_tensor_work"), r["us_per_launch"])
print({k: (round(v["frac"], 3), round(v.get("us_per_launch", 0), 1)) for k, v in r["kernels"].items()})
PY
