"""Dev tool: one-line summary of bench.py JSON lines read from stdin (ignores non-JSON lines)."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
for line in sys.stdin:
    if line.startswith("{"):
        d = json.loads(line)
        s = d.get("stage_ms_per_step", {})
        print(tag, "value %.1f e2e %.1f ms/step %.3f stageA %.3f stageB %.3f convs %.3f" % (
            d["value"], d["e2e"]["value"], d["ms_per_step"], s.get("stage_a_warp", 0), s.get("stage_b_net", 0),
            s.get("profiled_pass", {}).get("convs", 0)))
