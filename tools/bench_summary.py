#!/usr/bin/env python
"""One screen of a bench.py JSON line (stdin): headline, stages, e2e, the configs block."""
import json
import sys

d = json.loads(sys.stdin.read())
r = lambda v: round(v, 2) if isinstance(v, (int, float)) else v
print("c2 gpus", d["n_gpus"], "ms/step", r(d["ms_per_step"]), {k: r(v) for k, v in d["stage_ms"].items()},
      "e2e", r(d["e2e"]["ms_per_step"]), "pageable", r(d["e2e_pageable"]["ms_per_step"]), "cold", r(d["e2e_cold"]["ms"]),
      "parity", d["parity"]["ok"], "sha", d["table_sha256"][:12], "clocks", d["clocks"])
for k, v in d.get("configs", {}).items():
    if "error" in v:
        print(k, "ERROR", v["error"])
    elif k == "c3":
        print(k, {m: (r(v[m]["value"]), v[m]["table_sha256"][:12]) for m in ("default", "fast_version")})
    else:
        print(k, "ms/step", r(v["ms_per_step"]), {a: r(b) for a, b in v["stage_ms"].items()}, "e2e", r(v["e2e"]["ms_per_step"]),
              "sha", v["table_sha256"][:12])
