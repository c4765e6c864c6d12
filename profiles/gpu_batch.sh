mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_modules.py -x -q -m gpu -s -k "small_cin or raft_bi or shipping_defaults" 2>&1 | grep -v "^$" | tail -12
B="timeout 400 python bench.py --steps 6 --warmup 3 --no-strong --no-gpu-reference --no-cpu-baseline"
run() { name=$1; shift; "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err; python - <<PY
import json
for l in open("gpurun_out/$name.log"):
    if l.startswith("{"):
        d = json.loads(l); print("$name", round(d["value"], 1), "fps", round(d["ms_per_step"], 1), "ms e2e", round(d["e2e"]["value"], 1), d.get("single_clip"), d["clocks"])
PY
grep "autotune plans\|Error\|error" gpurun_out/$name.err | cut -c1-1500; }
run s_c1 $B
PP_SMALL_CIN=0 run s0_c1 $B
run s_c2 $B --clips-in-flight 2
run s_c3 $B --clips-in-flight 3
