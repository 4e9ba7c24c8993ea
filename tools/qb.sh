#!/bin/bash
# quick bench: prints "<label> sims/s ms_per_step conv3_ms" ; usage: tools/qb.sh label [bench args...]
label=$1; shift
out=$(timeout 300 python bench.py --dtype bf16 --steps 3 --warmup 3 --no-cpu-baseline --no-vl --no-train --preroll 0 "$@" 2>&1 | tail -1)
python - "$label" "$out" <<'PY'
import sys, json
try:
    d = json.loads(sys.argv[2]); k = d["roofline"]["kernel_ms_per_launch"]; print(sys.argv[1], round(d["value"]), round(d["ms_per_step"], 1), round(k * 1000, 1) if k else None, "us conv3", "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[1], "FAILED", sys.argv[2][-300:])
PY
