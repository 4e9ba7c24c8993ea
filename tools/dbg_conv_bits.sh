B="timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-vl --no-train --preroll 16"
for d in 0 1 2 4 3 7; do
AZP_DYN_TILES=0 AZP_B200_LIB=build/libazp_b200_bringup.so AZP_DEBUG_CONV=$d $B > gpurun_out/r3n_dbg$d.json 2> gpurun_out/r3n_dbg$d.err; echo "debug=$d rc=$?"
done
python - <<PY
import json
for n in [0,1,2,4,3,7]:
    try:
        d=json.load(open("gpurun_out/r3n_dbg%d.json"%n)); o=d["roofline"]["other_kernels_us_per_launch"]; print("debug", n, "conv2 %.1f conv3 %.1f conv1 %.1f ms/step %.1f" % (o["conv2"], o["dominant"], o["conv1"], d["ms_per_step"]))
    except Exception as ex: print(n, "ERR", ex)
PY
