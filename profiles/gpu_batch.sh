mkdir -p gpurun_out
B="timeout 400 python bench.py --steps 6 --warmup 3 --no-strong --no-gpu-reference --no-cpu-baseline"
run() { name=$1; shift; "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err; python - <<PY
import json
for l in open("gpurun_out/$name.log"):
    if l.startswith("{"):
        d = json.loads(l); print("$name", round(d["value"], 1), "fps", round(d["ms_per_step"], 1), "ms e2e", round(d["e2e"]["value"], 1), d.get("single_clip"), d["clocks"])
PY
grep "Error\|error" gpurun_out/$name.err | cut -c1-600; }
run p0_c1 $B
export PP_SCAN_PRIORITY=1
run p1_c1 $B
run p1_c2 $B --clips-in-flight 2
run p1_c3 $B --clips-in-flight 3
