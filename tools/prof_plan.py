"""Per-launcher device time of one step of a Python-composed plan (eager, CUDA events on the launching stream).
    python tools/prof_plan.py dcgan|kc|moons_cf|cgan_moons|simple_moons|cwgan"""
import ctypes, json, sys
sys.path.insert(0, '.')
import torch
import bench
import pcg_b200
from pcg_b200 import _lib
kind = sys.argv[1]
B = bench.OTHER[kind]["batch"]
L = _lib.load()
import pcg_b200.dcgan.plan as DP, pcg_b200.tabular.layers as TL, pcg_b200.moons.gan as MG
native_step, _, _ = bench._other_setup(kind, B, torch.device("cuda"))
native_step(0); native_step(1); torch.cuda.synchronize()
# find the plan object through the closure and force eager execution
plan = [c.cell_contents for c in native_step.__closure__ if hasattr(c.cell_contents, "step")][0]
if hasattr(plan, "use_graph"): plan.use_graph = False
if hasattr(getattr(plan, "run", None), "use_graph"): plan.run.use_graph = False
_lib.check(L.pcg_profile_begin())
n = 3
for i in range(n): native_step(i)
buf = ctypes.create_string_buffer(1 << 16)
_lib.check(L.pcg_profile_end(buf, ctypes.c_size_t(len(buf))))
prof = json.loads(buf.value.decode())
tot = sum(v["ms"] for v in prof.values())
print(f"{kind}: {tot / n:.3f} ms/step summed over launchers")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {v['ms'] / n:8.4f} ms  {v['launches'] // n:4d} launches  {k}")
