# one box, one GPU: GPU test suite, smoke, bench records of every BASELINE config (profiles/r3_bench_*), reference arm
set -x
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/rF_gputests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/rF_gputests.log
timeout 900 python bench.py > gpurun_out/rF_bench_c3.json 2> gpurun_out/rF_bench_c3.err; echo rc=$?
timeout 600 python bench.py --impl reference > gpurun_out/rF_bench_c3_reference.json 2> gpurun_out/rF_bench_c3_reference.err; echo rc=$?
for c in 2 1 5; do timeout 900 python bench.py --config $c --no-train > gpurun_out/rF_bench_c$c.json 2> gpurun_out/rF_bench_c$c.err; echo "config $c rc=$?"; done
python - <<PY
import json
for c in ["c3","c2","c1","c5","c3_reference"]:
    try:
        d=json.load(open("gpurun_out/rF_bench_%s.json"%c)); print(c, round(d["value"]), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), (d.get("roofline") or {}).get("frac"), (d.get("roofline") or {}).get("net_tensor_frac_end_to_end"), d.get("clocks"), (d.get("arena_match") or {}).get("seconds"))
    except Exception as ex: print(c, "ERR", ex)
PY
