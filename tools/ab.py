#!/usr/bin/env python
"""A/B timing of library variants: runs bench.py (device leg only) once per library given on the
command line (FCPP_LIB) and prints plans/s and the kernel times.
    python tools/ab.py field_coverage_path_planning_b200/libfcpp.so field_coverage_path_planning_b200/variants/libfcpp_x.so"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for lib in sys.argv[1:]:
    env = dict(os.environ, FCPP_LIB=os.path.join(ROOT, lib))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "20", "--warmup", "3", "--no-cpu"],
                         env=env, capture_output=True, text=True)
    try:
        j = json.loads(out.stdout.strip().splitlines()[-1])
        print(lib, round(j["value"]), round(j["e2e"]["value"]), j["roofline"]["kernel_ms"], flush=True)
    except Exception:
        print(lib, "FAILED", out.stdout[-500:], out.stderr[-1500:], flush=True)
