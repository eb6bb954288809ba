#!/bin/bash
# Final-state evidence in one gpurun call (C8W compiled in as the default): GPU suite, default bench,
# BASELINE configs 1/2/3/5, Generator timeline.  Stage-wise outputs under gpurun_out/.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t0=$(date +%s)
timeout 300 python -m pytest tests -m gpu -q -s -p no:cacheprovider --durations=8 > gpurun_out/r02_pytest_gpu_final_c8w.log 2>&1
echo "pytest rc=$? after $(( $(date +%s) - t0 )) s" | tee -a gpurun_out/r02_pytest_gpu_final_c8w.log
timeout 150 python bench.py > gpurun_out/r02_bench_c8w_default.json 2> gpurun_out/r02_bench_c8w_default.err
echo "bench rc=$? after $(( $(date +%s) - t0 )) s"
timeout 120 python tools/run_configs.py c8w > gpurun_out/r02_configs_c8w.json 2> gpurun_out/r02_configs_c8w.err
echo "configs rc=$? after $(( $(date +%s) - t0 )) s"
timeout 60 python tools/g_timeline.py --precision c8w --out gpurun_out/r02_g_timeline_c8w.md > /dev/null 2> gpurun_out/r02_g_timeline_c8w.err
echo "timeline rc=$? after $(( $(date +%s) - t0 )) s"
tail -2 gpurun_out/r02_pytest_gpu_final_c8w.log
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_c8w_default.json").read().strip().splitlines()[-1])
    print("default (%s): %.2f ms/step, %.0f frames/s, e2e %.0f, clocks %s" % (d["engine"]["precision_mode"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["clocks"]))
    for o in d["other_precision"]:
        print("  other:", o["precision"], "lean" if o.get("lean") else "", "%.2f ms/step" % o["ms_per_step"])
except Exception as e:
    print("bench line unreadable:", e)
try:
    print(open("gpurun_out/r02_configs_c8w.json").read()[-900:])
except Exception as e:
    print(e)
P
