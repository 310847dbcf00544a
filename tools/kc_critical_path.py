#!/usr/bin/env python
"""Timed critical path of the data-flow captured KC CounteRGAN plan (pcg_b200.dataflow.Program.timed_critical_path):
which operators bound the step once launch latency is hidden by the parallel branches.  python tools/kc_critical_path.py [kc|moons_cf]"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "kc"
B = bench.OTHER[kind]["batch"]
step, _, _ = bench._other_setup(kind, B, torch.device("cuda"))
for i in range(3):
    step(i)
torch.cuda.synchronize()
plan = next(c.cell_contents for c in step.__closure__ if hasattr(c.cell_contents, "run"))
prog = plan.run.program
total, path, allus = prog.timed_critical_path()
print(f"{kind}: {len(prog.ops)} operators, {sum(1 for _ in path)} on the timed critical path = {total:.1f} us "
      f"(all operators one after another: {allus:.1f} us)")
agg = collections.OrderedDict()
for _, name, us in path:
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {us:8.1f} us  {n:3d} x {name}")
print("in order:", " ".join(f"{name}:{us:.0f}" for _, name, us in path))
