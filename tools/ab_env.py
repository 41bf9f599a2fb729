#!/usr/bin/env python
"""A/B a runtime switch of the library on the GPU: tools/ab_env.py VAR=val0,val1 [bench args...]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
var, vals = sys.argv[1].split("=")
for rep in range(2):
    for v in vals.split(","):
        for extra in (["--dtype", "f32"], ["--dtype", "f64"], ["--workload", "config3"], ["--workload", "weak16k", "--steps", "100"]):
            env = dict(os.environ, **{var: v})
            out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu"] + extra + sys.argv[2:],
                                 capture_output=True, text=True, env=env)
            try:
                d = json.loads(out.stdout.strip().splitlines()[-1])
                print(f"{var}={v} {' '.join(extra):34s} {d['value']:.2f} GLUPS frac {d['roofline']['frac']:.4f} e2e {d['e2e']['value']:.2f}", flush=True)
            except Exception:
                print(var, v, extra, "FAILED", out.stderr[-400:], flush=True)
