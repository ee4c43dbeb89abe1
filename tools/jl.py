#!/usr/bin/env python
"""Print selected fields of the last JSON line on stdin: tools/jl.py label"""
import json, sys
lab = sys.argv[1] if len(sys.argv) > 1 else ""
d = [json.loads(l) for l in sys.stdin if l.startswith("{")][-1]
r = d.get("roofline") or {}
print(lab, "%.3fM solves/s" % (d["value"] / 1e6), "%.2f ms/step" % d["ms_per_step"], "roofline %.3f" % (r.get("frac") or 0),
      "conv %.4f" % d["config"].get("converged_fraction", 0), "|", d["config"].get("collective", "")[:50],
      "| e2e", (d.get("e2e") or {}).get("value"))
