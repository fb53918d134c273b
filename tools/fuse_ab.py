#!/usr/bin/env python
"""A/B the opt-in fused layer-1 + w_v#0 kernel (option fuse_l1) against the two separate kernels on one box:
per-stage times (ms per 1024 windows) and a bitwise comparison of the probabilities.  Usage: python tools/fuse_ab.py"""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
from genomad_b200 import engine
from gpu_diag import make_windows
n = 1024
clf = engine.Classifier(None, device=0, max_batch=n)
a = torch.from_numpy(make_windows(n)).cuda()
out = torch.empty((n, 3), dtype=torch.float32, device="cuda")
ref = None
for rep in range(2):
  for f in (0, 1):
    clf.set_option("fuse_l1", f)
    for _ in range(2): clf.predict_ascii(a, out)
    clf.set_option("profile_stages", 1)
    for _ in range(5): clf.predict_ascii(a, out)
    torch.cuda.synchronize()
    acc = {}
    for name, ms in clf.stage_times(): acc.setdefault(name, []).append(ms)
    clf.set_option("profile_stages", 0)
    res = out.cpu().numpy().copy(); ref = res if ref is None else ref
    print(f"fuse_l1={f}: " + "  ".join(f"{k}={sum(v)/len(v):.3f}" for k, v in acc.items()) + f"  total={sum(sum(v)/len(v) for v in acc.values()):.3f}  bitwise_same={bool((res==ref).all())}", flush=True)
