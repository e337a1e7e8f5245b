"""Hot source lines of one kernel in an .ncu-rep (read here, no GPU): warp instructions executed and stall samples per CUDA
source line.  Usage: python scripts/ncu_hot.py REP KERNEL_REGEX [top]"""
import csv, subprocess, sys
from collections import defaultdict

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = defaultdict(lambda: [0, 0, ""])
cur_file, hdr, cur_line, total_i, total_s = "", None, None, 0, 0
launch = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) == 2:
        if r[0] == "Function Name":
            launch += 1
        continue
    if r and r[0] == "Line No":
        hdr = r
        i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    if r[0] != "":
        cur_line = (cur_file, r[0])
        agg[cur_line][2] = r[1].strip()[:110]
    else:
        try:
            ins, smp = int(r[i_inst]), int(r[i_samp])
        except ValueError:
            continue
        agg[cur_line][0] += ins
        agg[cur_line][1] += smp
        total_i += ins
        total_s += smp
print(f"total warp instructions {total_i:,}  stall samples {total_s:,}")
for (f, ln), (ins, smp, src) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100.0 * smp / max(total_s, 1):5.1f}% smp {100.0 * ins / max(total_i, 1):5.1f}% ins  {f}:{ln:>4}  {src}")

# instruction / sample share per source-line bucket (optional 4th argument: "file:lo-hi=name,...")
if len(sys.argv) > 4:
    buckets = []
    for spec in sys.argv[4].split(","):
        rng, name = spec.split("=")
        f, lh = rng.split(":")
        lo, hi = lh.split("-")
        buckets.append((f, int(lo), int(hi), name))
    tot = defaultdict(lambda: [0, 0])
    for (f, ln), (ins, smp, _) in agg.items():
        try:
            l = int(ln)
        except ValueError:
            continue
        name = "other"
        for bf, lo, hi, bn in buckets:
            if f == bf and lo <= l <= hi:
                name = bn
                break
        tot[name][0] += ins
        tot[name][1] += smp
    print("--- buckets")
    for name, (ins, smp) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        print(f"{100.0 * ins / max(total_i, 1):5.1f}% ins {100.0 * smp / max(total_s, 1):5.1f}% smp  {name}")
