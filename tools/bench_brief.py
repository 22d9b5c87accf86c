#!/usr/bin/env python
"""One line per bench.py JSON line on stdin (A/B runs): value, e2e, sequential."""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
for x in sys.stdin:
    if not x.startswith("{"): continue
    l = json.loads(x)
    print(tag, "value", round(l["value"]), "ms/step", round(l["ms_per_step"], 3), "e2e", round(l["e2e"]["value"]),
          "pageable", l["e2e"].get("pageable_value"), "alone ms", round(l["sequential"]["ms_per_step"], 2),
          "e2e alone ms", l["sequential"].get("e2e_ms_per_step"), "drain", l["per_rank"]["drain_ms"])
