"""Print the few numbers of a bench.py JSON line that matter while iterating."""
import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    print("value %.0f graphs/s  %.3f ms/step (eager %.3f)  e2e %.0f (%.3f ms)  launches/step %s  roof %.4f (%.3f ms)" % (
        d["value"], d["ms_per_step"], d.get("ms_per_step_eager", 0), d["e2e"]["value"], d["e2e"].get("ms_per_step", 0),
        d.get("gpu_launches_per_step"), (d.get("roofline") or {}).get("frac", 0), (d.get("roofline") or {}).get("ms_per_launch", 0)))
