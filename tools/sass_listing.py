#!/usr/bin/env python
"""cuobjdump -sass listings of the tcgen05 / bulk-copy kernels for profiles/ (runs in the build container, no GPU needed).

    python tools/sass_listing.py profiles/r02
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "fully-automated-multi-heartbeat-echocardiography-video-segmentation-and-motion-tracking_b200", "csrc", "libclasfv_b200.so")
WANT = ["UTCHMMA", "UTCCP", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "LDS", "LDG", "STG", "F2FP", "FADD2", "MAPA", "UCGABAR"]
KERNELS = {          # output suffix -> (substring of the mangled name, note)
    "conv_umma": ("conv_umma_kernelILi1E", "single-CTA instantiation"),
    "conv_umma_pair": ("conv_umma_kernelILi2E", "CTA-pair instantiation (cta_group::2)"),
    "head_umma": ("head_umma_kernelI13__nv_bfloat16Lb0E", "bf16 outputs"),
    "warp_fuse_staged": ("warp_fuse_staged_kernelI13__nv_bfloat16Li512ELi7ELb1E", "bf16, flow-staged ring units"),
}


def main():
    prefix = sys.argv[1]
    names = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    funcs = re.findall(r"Function : (\S+)", names)
    for suffix, (sub, note) in KERNELS.items():
        fn = [f for f in funcs if sub in f]
        if not fn:
            print("missing", sub)
            continue
        out = subprocess.run(["cuobjdump", "-sass", "-fun", fn[0], SO], capture_output=True, text=True, check=True).stdout
        lines = [l for l in out.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l)]
        ops = collections.Counter()
        for l in lines:
            m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", l)
            if m:
                ops[m.group(1)] += 1
        counts = ", ".join(f"{k} {ops[k]}" for k in WANT if ops[k] or k == "UTMASTG")
        with open(f"{prefix}_sass_{suffix}.txt", "w") as f:
            f.write(f"# cuobjdump -sass of libclasfv_b200.so, function {fn[0]} ({note})\n")
            f.write(f"# {len(lines)} SASS instructions; mnemonic counts: {counts}\n")
            f.write("\n".join(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l) for l in lines) + "\n")
        print(suffix, len(lines), counts)


if __name__ == "__main__":
    main()
