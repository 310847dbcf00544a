"""Informational comparator: the UNMODIFIED reference train_countergan (baseline/_ref) under
PyTorch eager on ONE B200 - cuDNN / cuBLAS, the reference's own Python loop including its per-sample mask construction -
at batch 512, (a) fp32 with TF32 disabled, (b) TF32 enabled, (c) bf16 autocast.  Prints one JSON line per variant."""
import json
import sys
sys.path.insert(0, '.')
import torch
import bench

for name, tf32, ac in (("fp32 (TF32 off)", False, None), ("fp32 storage, TF32 on", True, None), ("bf16 autocast", True, torch.bfloat16)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    r = bench.reference_step_rate(20, 5, device="cuda", autocast=ac, budget_s=60.0)
    if r is None:
        print(json.dumps({"unavailable": "baseline/_ref not staged"}))
        break
    rate, ms, _, n = r
    print(json.dumps({"impl": "reference modules + reference trainer, torch eager on B200", "variant": name, "batch": 512,
                      "value": rate, "unit": "samples/s", "ms_per_step": ms, "timed_steps": n}), flush=True)
