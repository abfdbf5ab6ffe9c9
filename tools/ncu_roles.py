"""Per-line stall samples / executed instructions of one kernel from an ncu report with --import-source on.

    python tools/ncu_roles.py gpurun_out/prof.ncu-rep [top_n]

Prints the wait loops (mbarrier try_wait) with their executed counts and samples, and the top sampled lines:
the quickest way to see which role of a warp-specialised kernel everything else is waiting for.
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
isrc, iss, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = rows[2:]
tot_s = sum(int(r[iss] or 0) for r in data)
tot_e = sum(int(r[iex] or 0) for r in data)
print(f"{rows[0][1][:90]}\nsamples {tot_s}  warp-instructions {tot_e}  SASS lines {len(data)}")
print("-- sync / MMA / TMEM / bulk-copy instructions (line, executed, samples in the next 4 lines)")
for i, r in enumerate(data):
    if any(k in r[isrc] for k in ("SYNCS.PHASECHK", "UTCBAR", "UBLKCP", "BAR.SYNC", "NANOSLEEP")):
        smp = sum(int(data[j][iss] or 0) for j in range(i, min(i + 4, len(data))))
        print(f"{i:5d} {r[iex]:>9} {smp:6d} | {r[isrc][:90].strip()}")
print("-- top sampled lines")
for r in sorted(data, key=lambda r: -int(r[iss] or 0))[:top_n]:
    print(f"{r[iss]:>7} {r[iex]:>9} | {r[isrc][:100].strip()}")
if len(sys.argv) > 3:
    a, b = int(sys.argv[3]), int(sys.argv[4])
    for i in range(a, b):
        r = data[i]
        print(f"{i:5d} {r[iex]:>9} {r[iss]:>5} | {r[isrc][:100].strip()}")
