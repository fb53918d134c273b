#!/usr/bin/env python
"""
GPU bring-up diagnostic: runs the CUDA path stage by stage against the CPU oracle and prints every
difference (does not stop at the first failure).  Each risky configuration runs in its own
subprocess because a device-side trap poisons the CUDA context.

    python tools/gpu_diag.py                 # all configurations
    python tools/gpu_diag.py --one tc 0 synthetic 4
"""
import argparse
import json
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def make_windows(n, seed=11):
    rng = np.random.default_rng(seed)
    a = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, (n, 6000))].copy()
    if n > 1:
        a[1, 3000:3400] = ord("N")
        a[1, 10] = ord("R")
    if n > 2:
        a[2, 2600:] = ord("N")
    if n > 3:
        a[3, :] = ord("A")
    return a


def maxdiff(name, got, ref, tol):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    d = np.abs(got - ref)
    i = np.unravel_index(d.argmax(), d.shape)
    tol = tol * max(1.0, float(np.abs(ref).max()))   # tolerance relative to the tensor's scale
    ok = bool(d.max() <= tol)
    print(f"  [{'ok' if ok else 'FAIL'}] {name}: max|d|={d.max():.3e} at {i} got={got[i]:.6g} ref={ref[i]:.6g} "
          f"(|ref|max={np.abs(ref).max():.3g}, tol={tol:g})", flush=True)
    return ok


def run_one(impl, mode, wkind, n):
    import torch
    from genomad_b200 import engine, weights as W
    from oracle import igloo_model as M, tokenizer as T

    w = W.load_weights()
    if wkind == "synthetic":
        w = M.synthetic_igloo_weights(w)
    a = make_windows(n)
    tok = T.tokenize_windows(a)
    probs_ref, inter = M.forward(tok, w, torch.float32, return_intermediates=True)
    probs64 = M.forward(tok, w, torch.float64)

    clf = engine.Classifier(w, device=0, max_batch=max(n, 8))
    clf.set_option("conv_impl", 1 if impl == "ref" else 0)
    da = torch.from_numpy(a).cuda()
    ok = True

    t_gpu = clf.encode(da).cpu().numpy()
    eq = np.array_equal(t_gpu, tok)
    print(f"  [{'ok' if eq else 'FAIL'}] encode bit-exact: {eq} (mismatches: {(t_gpu != tok).sum()})", flush=True)
    ok &= eq

    clf.set_option("debug_stop", 1)
    clf.predict_ascii(da); torch.cuda.synchronize()
    ok &= maxdiff("y1 (embed conv1)", clf.debug_fetch("buf0", n).cpu().numpy(), inter["y1"].numpy(), 2e-6)
    ok &= maxdiff("mpi0 (gather on y1)", clf.debug_fetch("mpi0", n).cpu().numpy(), inter["ig0"]["mpi"].numpy(),
                  1e-30 if wkind == "shipped" else 2e-4)

    clf.set_option("debug_stop", 2)
    clf.predict_ascii(da); torch.cuda.synchronize()
    ok &= maxdiff("y2 (conv2)", clf.debug_fetch("buf1", n).cpu().numpy(), inter["y2"].numpy(), 6e-5)
    ok &= maxdiff("q0 (w_v0 + maxpool on y1)", clf.debug_fetch("q0", n).cpu().numpy(), inter["ig0"]["q"].numpy(), 3e-6)

    clf.set_option("debug_stop", 3)
    clf.predict_ascii(da); torch.cuda.synchronize()
    ok &= maxdiff("y3 (conv3)", clf.debug_fetch("buf0", n).cpu().numpy(), inter["y3"].numpy(), 2e-5)

    clf.set_option("debug_stop", 0)
    p = clf.predict_ascii(da).cpu().numpy()
    ok &= maxdiff("q1", clf.debug_fetch("q1", n).cpu().numpy(), inter["ig1"]["q"].numpy(), 8e-6)
    ok &= maxdiff("mpi1", clf.debug_fetch("mpi1", n).cpu().numpy(), inter["ig1"]["mpi"].numpy(),
                  1e-30 if wkind == "shipped" else 2e-4)
    ok &= maxdiff("h0 (attention out)", clf.debug_fetch("h0", n).cpu().numpy(), inter["h0"].numpy(), 1e-5)
    ok &= maxdiff("probs vs fp32 oracle", p, probs_ref, 1e-4)
    ok &= maxdiff("probs vs fp64 oracle", p, probs64, 1e-4)
    p_tok = clf.predict_tokens(torch.from_numpy(tok.view(np.int16)).cuda().view(torch.uint16)).cpu().numpy()
    ok &= maxdiff("probs(tokens) vs probs(ascii)", p_tok, p, 0.0)
    p_host = clf.classify_host(a)
    ok &= maxdiff("classify_host vs device path", p_host, p, 0.0)
    print(f"  launches so far: {clf.kernel_launches}")
    return ok


def bench_one(impl, mode, n):
    import torch
    from genomad_b200 import engine
    clf = engine.Classifier(None, device=0, max_batch=n)
    clf.set_option("conv_impl", 1 if impl == "ref" else 0)
    a = torch.from_numpy(make_windows(n)).cuda()
    out = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    for _ in range(2):
        clf.predict_ascii(a, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        clf.predict_ascii(a, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"  bench impl={impl} n={n}: {ms:.3f} ms/step -> {n / ms * 1e3:.0f} windows/s", flush=True)
    clf.set_option("profile_stages", 1)
    clf.predict_ascii(a, out); torch.cuda.synchronize()
    for name, t in clf.stage_times():
        print(f"     {name:14s} {t:9.3f} ms  ({t / n * 1e3:8.3f} us/window)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--one", nargs=4, metavar=("IMPL", "MODE", "WEIGHTS", "N"))
    ap.add_argument("--bench", nargs=3, metavar=("IMPL", "MODE", "N"))
    args = ap.parse_args()
    if args.one:
        impl, mode, wkind, n = args.one
        ok = run_one(impl, int(mode), wkind, int(n))
        sys.exit(0 if ok else 3)
    if args.bench:
        impl, mode, n = args.bench
        bench_one(impl, int(mode), int(n))
        return
    results = {}
    configs = [("ref", 0, "synthetic", 4), ("tc", 0, "synthetic", 4), ("tc", 0, "shipped", 4)]
    for impl, mode, wkind, n in configs:
        print(f"=== impl={impl} desc_base_mode={mode} weights={wkind} n={n}", flush=True)
        t0 = time.time()
        r = subprocess.run([sys.executable, __file__, "--one", impl, str(mode), wkind, str(n)], timeout=600)
        results[f"{impl}/{mode}/{wkind}"] = r.returncode
        print(f"=== exit {r.returncode} in {time.time() - t0:.1f}s", flush=True)
    good_mode = 0 if results.get("tc/0/synthetic") == 0 else None
    print("RESULTS", json.dumps(results), "good_mode", good_mode, flush=True)
    for impl, n in (("ref", 64), ("tc", 1024)):
        if impl == "tc" and good_mode is None:
            continue
        print(f"=== bench impl={impl}", flush=True)
        subprocess.run([sys.executable, __file__, "--bench", impl, str(good_mode or 0), str(n)], timeout=600)


if __name__ == "__main__":
    main()
