# one box, one GPU: GPU test suite, bench records of every BASELINE config, reference arm, step timeline, ncu launch list + full capture
set -x
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r4_gputests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r4_gputests.log
timeout 900 python bench.py > gpurun_out/r4_bench_c3.json 2> gpurun_out/r4_bench_c3.err; echo rc=$?
timeout 600 python bench.py --impl reference > gpurun_out/r4_bench_c3_reference.json 2> gpurun_out/r4_bench_c3_reference.err; echo rc=$?
for c in 2 1 4 5; do timeout 900 python bench.py --config $c > gpurun_out/r4_bench_c$c.json 2> gpurun_out/r4_bench_c$c.err; echo "config $c rc=$?"; done
AZP_B200_LIB=build/libazp_b200_bringup.so AZP_TIMELINE=1 timeout 300 python tools/step_timeline.py --moves 3 > gpurun_out/r4_timeline.txt 2>&1; echo rc=$?
AZP_FUSE_DESCENT=0 AZP_B200_LIB=build/libazp_b200_bringup.so AZP_TIMELINE=1 timeout 300 python tools/step_timeline.py --moves 3 > gpurun_out/r4_timeline_unfused.txt 2>&1; echo rc=$?
SM="python bench.py --steps 1 --warmup 3 --sims 24 --preroll 40 --no-cpu-baseline --no-vl --no-train"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r4_launches.csv $SM > gpurun_out/r4_ncu1.log 2>&1; echo rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv1_taps|conv_umma|fc_umma|descend_conv1" -s 160 -c 10 -o gpurun_out/r4_full -f $SM > gpurun_out/r4_ncu2.log 2>&1; echo rc=$?
python - <<PY
import json
for c in ["c3","c2","c1","c4","c5","c3_reference"]:
    try:
        d=json.load(open("gpurun_out/r4_bench_%s.json"%c)); print(c, round(d["value"]), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), (d.get("roofline") or {}).get("frac"), d.get("clocks"))
    except Exception as ex: print(c, "ERR", ex)
PY
