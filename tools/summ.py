"""print the key figures of bench.py JSON lines read from stdin or files: tools/summ.py a.json b.json"""
import json
import sys

for path in sys.argv[1:] or ["-"]:
    txt = sys.stdin.read() if path == "-" else open(path).read()
    for line in txt.splitlines():
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        rf = d.get("roofline") or {}
        print(path, "value", round(d.get("value", 0), 1), "ms/step", round(d.get("ms_per_step", 0), 2), "pcg_ms", round(d.get("pcg_ms", 0), 2),
              "iters", d.get("pcg_iters"), "spmv_ms", rf.get("avg_launch_ms"), "n", rf.get("launches_timed"), "iso_ms", rf.get("isolated_ms"),
              "launches", d.get("gpu_launches"), "mhz", (d.get("clocks") or {}).get("sm_mhz"), "e2e", (d.get("e2e") or {}).get("value"))
