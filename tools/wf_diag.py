#!/usr/bin/env python
"""Staged vs direct-gather warp-fuse kernel, bit comparison in separate processes (the kernel choice is latched per process
by CLASFV_WARP_FUSE_DIRECT).  Repeats every (dtype, edge_hops) case and also compares each kernel with ITSELF across
processes, to tell a rounding difference between the kernels from run-to-run non-determinism of one of them.

    python tools/wf_diag.py [repeats]
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "tests", "test_gpu_parity.py")).read()
body = src[src.index('_WF_SCRIPT = r"""') + len('_WF_SCRIPT = r"""'):]
open("/tmp/wf.py", "w").write(body[:body.index('"""')].format(root=ROOT))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_parity  # noqa: E402  (the inputs are generated once, here, and shared by every process)
test_gpu_parity._wf_inputs("/tmp/wf_inputs.npz")
repeats = int(sys.argv[1]) if len(sys.argv) > 1 else 2


def run(name, dtype, edge, tag):
    e = dict(os.environ)
    e.pop("CLASFV_WARP_FUSE_DIRECT", None)
    if name == "direct":
        e["CLASFV_WARP_FUSE_DIRECT"] = "1"
    path = f"/tmp/wf_{name}_{tag}.npz"
    subprocess.run([sys.executable, "/tmp/wf.py", dtype, edge, path, "/tmp/wf_inputs.npz"], check=True, env=e)
    return np.load(path)["acc"]


def diff(a, b):
    d = a.view(np.uint32) != b.view(np.uint32)
    idx = np.argwhere(d)
    s = f"{int(d.sum())} of {d.size} differ, max abs {float(np.abs(a - b).max()):.3g}"
    if len(idx):
        s += f"; frames {np.unique(idx[:, 0])[:12]} (n={len(np.unique(idx[:, 0]))}) rows {np.unique(idx[:, 2])[:60]} cols {np.unique(idx[:, 3])[:70]}"
    return s


for dtype, edge in (("bf16", "0"), ("bf16", "1"), ("fp32", "0"), ("fp32", "1")):
    staged = [run("staged", dtype, edge, i) for i in range(repeats)]
    direct = [run("direct", dtype, edge, i) for i in range(repeats)]
    print(f"{dtype} edge_hops={edge}", flush=True)
    print("  staged vs direct       :", diff(staged[0], direct[0]), flush=True)
    for i in range(1, repeats):
        print(f"  staged run 0 vs run {i}  :", diff(staged[0], staged[i]))
        print(f"  direct run 0 vs run {i}  :", diff(direct[0], direct[i]), flush=True)
