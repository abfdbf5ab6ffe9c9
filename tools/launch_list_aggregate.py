#!/usr/bin/env python
"""Aggregate the ncu launch list of tools/launch_list.sh into per-kernel totals of ONE bench step.

    python tools/launch_list_aggregate.py gpurun_out/launches_dram_TAG.csv profiles/r01_final_dram_traffic_per_step.json

A step of the dense-video schedule starts with the video-level `stem_conv_kernel` launch and ends with the
`warp_fuse` launch; the capture window (-s / -c of launch_list.sh) rarely starts on a step boundary, so the step is stitched
from the window: the launches from the first stem_conv_kernel to the end of its step if that step is complete, else the tail of
the cut step + the head of the next one (identical launches in identical order step after step).
bench.py reads the JSON for `roofline.traffic`.
"""
import csv
import json
import sys


def short(name):
    n = name.split("(")[0].split("::")[-1]
    if n.startswith("conv_umma_kernel"):          # <1> single CTA, <2> CTA pairs: one trunk kernel for bench.py's roofline
        return "conv_umma_kernel"
    return n[:28]


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ci = {h: i for i, h in enumerate(hdr)}
    launches = {}
    for r in rows:
        d = launches.setdefault(int(r[ci["ID"]]), {"name": r[ci["Kernel Name"]], "grid": r[ci["Grid Size"]]})
        v = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]].lower()
        m = r[ci["Metric Name"]]
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        else:
            v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
        d[m] = v
    seq = [launches[i] for i in sorted(launches)]
    is_stem = [i for i, l in enumerate(seq) if "stem_conv_kernel" in l["name"]]
    is_fuse = [i for i, l in enumerate(seq) if "warp_fuse" in l["name"]]
    step = None
    for s in is_stem:                                       # a complete step inside the window
        e = next((f for f in is_fuse if f > s), None)
        nxt = next((t for t in is_stem if t > s), len(seq) + 1)
        if e is not None and e < nxt:
            step = seq[s:e + 1]
            how = f"launches {s}..{e} of the capture window"
            break
    if step is None:                                        # stitch: head of the last step + the rest from the cut step before it
        s1 = is_stem[-1]
        head = seq[s1:]
        f0 = max(f for f in is_fuse if f < s1)
        key = lambda l: (l["name"], l.get("grid"))
        for length in range(max(f0 + 1, len(head)), f0 + 1 + len(head) + 1):
            first_pos = length - 1 - f0                     # step position of window index 0
            if first_pos > len(head):
                break
            if all(key(seq[pos - first_pos]) == key(head[pos]) for pos in range(first_pos, len(head))):
                step = head + seq[len(head) - first_pos:f0 + 1]
                how = f"stitched: launches {s1}..{len(seq) - 1} + {len(head) - first_pos}..{f0} of the capture window"
                break
        if step is None:
            raise SystemExit("no complete step in the capture window and the two partial steps do not line up; widen -c")
    per = {}
    for l in step:
        k = per.setdefault(short(l["name"]), {"launches": 0, "us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
        k["launches"] += 1
        k["us"] += l.get("gpu__time_duration.sum", 0.0)
        k["dram_read_bytes"] += l.get("dram__bytes_read.sum", 0.0)
        k["dram_write_bytes"] += l.get("dram__bytes_write.sum", 0.0)
    per = dict(sorted(per.items(), key=lambda kv: -kv[1]["us"]))
    out = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum of one bench step "
                     f"({len(step)} launches; tools/launch_list.sh, {how})",
           "per_kernel": per,
           "total_us_serialised": sum(k["us"] for k in per.values()),
           "total_dram_bytes_per_step": sum(k["dram_read_bytes"] + k["dram_write_bytes"] for k in per.values())}
    json.dump(out, open(dst, "w"), indent=1)
    tot = out["total_us_serialised"]
    for n, k in per.items():
        print(f"{n:30s} {k['launches']:4d} {k['us']:10.1f} us {100 * k['us'] / tot:5.1f} %  rd {k['dram_read_bytes'] / 1e9:7.3f} GB  wr {k['dram_write_bytes'] / 1e9:7.3f} GB")
    print(f"total {len(step)} launches {tot:.1f} us, DRAM {out['total_dram_bytes_per_step'] / 1e9:.2f} GB")


if __name__ == "__main__":
    main()
