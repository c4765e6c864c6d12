mkdir -p gpurun_out
B="timeout 400 python bench.py --steps 6 --warmup 3 --no-strong --no-gpu-reference --no-cpu-baseline"
run() { name=$1; shift; "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err; python - <<PY
import json
for l in open("gpurun_out/$name.log"):
    if l.startswith("{"):
        d = json.loads(l); print("$name", round(d["value"], 1), "fps", round(d["ms_per_step"], 1), "ms e2e", round(d["e2e"]["value"], 1), d.get("single_clip"))
PY
grep "graph-timed\|Error\|error" gpurun_out/$name.err | cut -c1-600; }
run b_hybrid $B
export CUDA_DEVICE_MAX_CONNECTIONS=32
run b_conn32 $B
run b_conn32_c2 $B --clips-in-flight 2
run b_conn32_w5 $B --windows-in-flight 5
run b_conn32_c3 $B --clips-in-flight 3
