#!/usr/bin/env python
"""One-screen digest of a bench.py JSON line: step time, e2e, and ms/step + roofline fraction per kernel family."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e = d.get("e2e") or {}
print(f"ms/step {d.get('ms_per_step'):.3f}  e2e {e.get('ms_per_step', float('nan')):.3f}  launches {d.get('gpu_launches')}  ffma {d.get('ffma_fallbacks_in_tc')}  clocks {d.get('clocks', {}).get('sm_mhz')}")
rows = [d.get("roofline")] + list(d.get("roofline_secondary") or [])
for r in rows:
    if not r: continue
    frac = r.get("frac")
    print(f"  {r.get('kernel'):12s} {r.get('ms_per_step', 0):7.3f} ms  n={r.get('launches_per_step', '-')}  frac={frac if frac is None else round(frac, 3)}")
