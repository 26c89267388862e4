#!/bin/bash
# Every bench line of the round on one GPU: the five BASELINE configurations through bench.py (B200 arm with its CPU
# baseline leg, then the reference arm), appended to gpurun_out/r02_bench_lines.jsonl (copied to profiles/ afterwards).
out=${OUT:-gpurun_out/r02_bench_lines.jsonl}
: > "$out"
for wl in c4 c3 c2 c5 c1; do
  python bench.py --workload $wl 2> gpurun_out/bench_$wl.err | grep '^{' >> "$out"
  python bench.py --impl reference --workload $wl 2> gpurun_out/ref_$wl.err | grep '^{' >> "$out"
done
python bench.py --workload c4 --waves 8 --no-cpu-baseline 2> gpurun_out/bench_c4w8.err | grep '^{' >> "$out"
OUT=$out python - <<'PY'
import json
import os
for l in open(os.environ.get("OUT", "gpurun_out/r02_bench_lines.jsonl")):
    d = json.loads(l)
    print(d.get("impl", "b200"), d["config"]["workload"][:40], "value %.4g" % d["value"], "ms/step %.4g" % d["ms_per_step"],
          "e2e %.4g" % d["e2e"]["value"], "frac", round(d.get("roofline", {}).get("frac") or 0, 3),
          "cpu", d.get("cpu_baseline", {}).get("value"))
PY
