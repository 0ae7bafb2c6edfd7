"""Source-level stall-sample summary of one kernel of an ncu report captured with --import-source on:
`python tools/ncu_hotspots.py gpurun_out/prof_mega.ncu-rep [top_n]` -> per source line: samples, share, executed warp
instructions; and the kernel-wide stall-reason histogram. Reads the report with `ncu --page source --csv` (no GPU needed)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sel = ["--launch-skip", sys.argv[3], "--launch-count", "1"] if len(sys.argv) > 3 else []  # one launch of the report (0-based)
raw = subprocess.run(["ncu", "-i", rep, *sel, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur, hdr, res, nk, kname = None, None, {}, 0, None
stalls = {}
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        nk += 1
        if nk > 1:
            break
        kname = r[1]
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        kname = kname or r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        ci = {n: i for i, n in enumerate(hdr)}
        scols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
        continue
    if hdr and r[0].isdigit():
        try:
            key = (cur, int(r[0]))
            e = res.setdefault(key, [0, r[1].strip()[:120], 0])
            e[0] += int(r[ci["# Samples"]])
            e[2] += int(r[ci["Instructions Executed"]])
            for s in scols:
                stalls[s] = stalls.get(s, 0) + int(r[ci[s]])
        except (ValueError, IndexError):
            pass
tot = sum(v[0] for v in res.values())
print(f"kernel: {kname} ({"launch " + sys.argv[3] if sel else "all captured launches"} of the report)\ntotal warp-stall samples: {tot}")
print("stall reasons:", ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in sorted(stalls.items(), key=lambda x: -x[1]) if v))
print(f"{'samples':>8} {'share':>6}  {'warp-instrs':>11}  location")
for (f, l), (s, src, n) in sorted(res.items(), key=lambda x: -x[1][0])[:top_n]:
    print(f"{s:8d} {100 * s / max(tot, 1):5.1f}%  {n:11d}  {f}:{l}  {src}")
