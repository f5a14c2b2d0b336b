#!/usr/bin/env python
"""Compact view of a bench.py JSON line: python profiles/show_bench.py file.json"""
import json, sys
for path in sys.argv[1:]:
    for line in open(path):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        print(path, "| value", round(d.get("value", 0), 1), d.get("unit"), "| ms/step", round(d.get("ms_per_step", 0), 3),
              "| e2e", round(d.get("e2e", {}).get("value", 0), 1), "| launches", d.get("gpu_launches"), "| clocks", d.get("clocks"))
        r = d.get("roofline", {})
        print("  roofline:", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k not in ("tensor", "hbm", "peak_source")})
        for k in ("tensor", "hbm"):
            if k in r:
                print("   ", k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in r[k].items()})
        st = d.get("stages") or d.get("stage_ms_per_step")
        if st:
            for k, v in st.items():
                print("   ", k, v)
        if "cpu_baseline" in d:
            print("  cpu:", d["cpu_baseline"])
