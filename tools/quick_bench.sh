#!/bin/bash
# short bench run; prints the headline numbers of every leg
mkdir -p gpurun_out
timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline $BENCH_ARGS > gpurun_out/bench_tmp.log 2>&1
tail -1 gpurun_out/bench_tmp.log > gpurun_out/bench_tmp.json
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_tmp.json"))
except Exception as e:
    print("bench failed:", e); print(open("gpurun_out/bench_tmp.log").read()[-3000:]); raise SystemExit(1)
print("infer ms", round(d["ms_per_step"], 3), "value", f'{d["value"]:.4g}', "e2e", f'{d["e2e"]["value"]:.4g}', "conv frac", round(d["roofline"]["frac"], 3), "fcomb ms", round(d["roofline_fcomb"]["kernel_ms_per_step"], 3), d["clocks"])
print("sweep", d.get("mc_sweep_px_samples_per_s")); print("single", d.get("single_image")); print("augment", d.get("augment"))
t = d.get("train")
if t:
    print("train ms", round(t["ms_per_step"], 3), "python-launched", round(t["ms_per_step_python_launched"], 3), "img/s", round(t["value"], 1), "e2e", round(t["e2e"]["value"], 1))
    print({k: round(v["ms_per_step"], 3) for k, v in t["kernels"].items()})
    print("src", t["source_train"]["value"] if t["source_train"] else None)
    for k, v in (t.get("joint_fixmatch") or {}).items():
        print(k, "graph ms", round(v["ms_per_step"], 3), "python ms", round(v["ms_per_step_python_launched"], 3), "img/s", round(v["value"], 1))
PY
