#!/usr/bin/env python
"""One line per bench JSON: tools/benchsum.py gpurun_out/<tag>_bench_*.json"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(f, "ERR", e)
        continue
    r = d.get("roofline") or {}
    e = d.get("e2e") or {}
    c = d.get("clocks") or {}
    print(f"{f.split('/')[-1]:34s} n={d.get('n_gpus')} value {d['value']:.3e} step {d.get('ms_per_step') or 0:.4f} ms | e2e {e.get('ms_per_step') or 0:.4f} ms {e.get('value', 0):.3e} | "
          f"kernel {r.get('kernel_ms') or 0:.4f} ms x{r.get('launches_timed')} frac {r.get('frac') or 0:.3f} | launches {d.get('gpu_launches')} | matches {d.get('matches_per_step')} | sm {c.get('sm_mhz')} {c.get('reasons')}")
