mkdir -p gpurun_out
timeout 300 python bench.py --steps 6 --warmup 3 --no-strong --no-gpu-reference --no-cpu-baseline --clips-in-flight 2 > gpurun_out/bench_c2fl.log 2> gpurun_out/bench_c2fl.err; python - <<PY
import json
for l in open("gpurun_out/bench_c2fl.log"):
    if l.startswith("{"):
        d = json.loads(l); print(round(d["value"], 1), "fps", round(d["ms_per_step"], 1), "ms e2e", round(d["e2e"]["value"], 1), d.get("single_clip"), d["clocks"], d["config"]["clips_in_flight"])
PY
tail -2 gpurun_out/bench_c2fl.err | cut -c1-300
