#!/bin/bash
# One gpurun call validating the C8W precision mode as the candidate default:
#   1. the whole GPU suite with MCGVC_PRECISION=c8w as the library default (what the suite runs if the
#      compiled-in default is flipped), every c8w referee included, without -x;
#   2. the default bench in c8w (its other_precision block re-times c8 and the other modes on the SAME box);
#   3. smoke() under the same default.
# Everything lands in gpurun_out/ stage by stage, so a call cut short still leaves what finished.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t0=$(date +%s)
MCGVC_PRECISION=c8w timeout 420 python -m pytest tests -m gpu -q -s -p no:cacheprovider --durations=12 > gpurun_out/r02_pytest_gpu_c8w_default.log 2>&1
echo "pytest rc=$? after $(( $(date +%s) - t0 )) s" | tee -a gpurun_out/r02_pytest_gpu_c8w_default.log
timeout 200 python bench.py --precision c8w --fast-steps 8 > gpurun_out/r02_bench_c8w.json 2> gpurun_out/r02_bench_c8w.err
echo "bench rc=$? after $(( $(date +%s) - t0 )) s"
MCGVC_PRECISION=c8w timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_c8w.log 2>&1
echo "smoke rc=$? after $(( $(date +%s) - t0 )) s"
tail -3 gpurun_out/r02_pytest_gpu_c8w_default.log; tail -2 gpurun_out/r02_smoke_c8w.log
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_c8w.json").read().strip().splitlines()[-1])
    print("c8w: %.2f ms/step, %.0f frames/s, e2e %.0f, clocks %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["clocks"]))
    for o in d["other_precision"]:
        print("  other:", o["precision"], "lean" if o.get("lean") else "", "%.2f ms/step" % o["ms_per_step"])
    print("  G fwd+bwd:", {k: round(v["ms"], 2) for k, v in d["generator_fwd_bwd"].items() if isinstance(v, dict)})
except Exception as e:
    print("bench line unreadable:", e)
P
