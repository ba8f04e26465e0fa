#!/usr/bin/env python
"""Markdown tables from bench.py JSON lines:  python tools/results_table.py profiles/r02_bench_*.json"""
import json
import sys


def line(path):
    rows = [l for l in open(path).read().splitlines() if l.startswith("{")]
    return json.loads(rows[-1]) if rows else None


def f(v, unit=1e6, nd=2):
    return "—" if v is None else ("%.*f" % (nd, v / unit))


print("| workload | N | ms/step (cold) | value (cold) M/s | resident plan M/s | warm M/s | moving M/s | e2e M/s | CPU all cores k/s (cores) | CPU 1 thread k/s | phases alone (ms) |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for p in sys.argv[1:]:
    d = line(p)
    if not d or d.get("impl") == "reference":
        continue
    cb = d.get("cpu_baseline") or {}
    ph = d.get("phase_ms_per_step", {})
    phs = ", ".join("%s %.2f" % (k, ph[k]) for k in ("flow", "los", "index", "cohesion", "velocity", "fields_cold_wall", "tick_alone") if k in ph)
    mv = d.get("value_moving") or {}
    print("| %s | %d | %.2f | %s | %s | %s | %s | %s | %s | %s | %s |" % (
        d["config"]["workload"].split(":")[0], d["n_gpus"], d["ms_per_step"], f(d["value"]), f(d.get("value_resident_plan")),
        f(d.get("value_warm")), f(mv.get("value")), f(d["e2e"]["value"]),
        ("%s (%s)" % (f(cb.get("value"), 1e3, 1), cb.get("cores"))) if cb else "—",
        f(cb.get("value_1thread"), 1e3, 2) if cb else "—", phs))
