import sys, os, subprocess, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
src = open(os.path.join(ROOT, "tests", "test_gpu_parity.py")).read()
body = src[src.index('_WF_SCRIPT = r"""') + len('_WF_SCRIPT = r"""'):]
body = body[:body.index('"""')].format(root=ROOT)
open("/tmp/wf.py", "w").write(body)
for dtype, edge in (("bf16", "0"),):
    out = {}
    for name, env in (("staged", {}), ("direct", {"CLASFV_WARP_FUSE_DIRECT": "1"})):
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, "/tmp/wf.py", dtype, edge, f"/tmp/{name}.npz"], check=True, env=e)
        out[name] = np.load(f"/tmp/{name}.npz")
    a, b = out["staged"]["acc"], out["direct"]["acc"]
    d = a.view(np.uint32) != b.view(np.uint32)
    idx = np.argwhere(d)
    print(dtype, "edge", edge, "differing", int(d.sum()), "of", d.size, "max abs", float(np.abs(a - b).max()),
          "cnt equal", bool(np.array_equal(out["staged"]["cnt"], out["direct"]["cnt"])), flush=True)
    if len(idx):
        print(" frames", np.unique(idx[:, 0])[:20], "classes", np.unique(idx[:, 1]), "rows", np.unique(idx[:, 2])[:12], "cols", np.unique(idx[:, 3])[:12])
        for f, c, y, x in idx[:6]:
            print("  ", f, c, y, x, a[f, c, y, x], b[f, c, y, x])
