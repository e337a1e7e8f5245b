"""Turn the raw outputs of scripts/gpu_r2_evidence.sh (gpurun_out/) into the tracked evidence under profiles/ (read here, no GPU):
bench lines, the launch list, the two ncu --set full summaries with hot source lines, and profiles/traffic.json."""
import collections, csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def last_json(path):
    return json.loads([l for l in open(path).read().strip().splitlines() if l.startswith("{")][-1])


d = last_json(os.path.join(G, "r2_bench_n1.log"))
json.dump(d, open(os.path.join(P, "r02_bench_n1.json"), "w"), indent=1)
json.dump(last_json(os.path.join(G, "r2_bench_ref.log")), open(os.path.join(P, "r02_bench_reference_arm.json"), "w"), indent=1)

# ---- launch list
rows = [r for r in csv.reader(open(os.path.join(G, "r2_launches_bench.csv"))) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
ix = {n: i for i, n in enumerate(rows[hdr])}
agg, seq = collections.OrderedDict(), []
for r in rows[hdr + 1:]:
    try:
        v = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    name, unit = r[ix["Kernel Name"]], r[ix["Metric Unit"]]
    ms = v * 1e-6 if unit in ("ns", "nsecond") else v * 1e-3 if unit in ("us", "usecond") else v
    if "ngf" not in name and "ntx" not in name:
        continue
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
    seq.append((name, ms))
tot = sum(a[1] for a in agg.values())
out = ["# ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv  python bench.py --steps 3 --warmup 3 --no-cpu-baseline",
       "# (scripts/gpu_r2_final.sh, end of round 2).  First 900 launches of the whole bench run: field packing, warm-up + serial",
       "# timing pass + timed 3-stream device-resident frames, e2e host-path chunks, camera e2e, dense-regime side run, InfoInv /",
       "# NeuTex side runs.  Kernels of this repo only (torch's own launches are left out).  Per-launch times are cold-cache and",
       "# serialised (ncu serialises the streams too): compare shares, not absolutes."]
for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{n[:100]:100s} n={c:4d} total_ms={ms:10.3f} share={ms / tot:.4f}")
m = [ms for n, ms in seq if "ngf_march_kernel<0, 0>" in n][:9]
c = [ms for n, ms in seq if "ngf_colour_kernel<0, 0>" in n][:9]
f = [ms for n, ms in seq if "ngf_finalize_kernel" in n][:9]
mm, cm, fm = sum(m) / len(m), sum(c) / len(c), sum(f) / len(f)
ro, ot = d["roofline"], d["roofline"]["other_kernel"]
km = {ro["kernel"][:16]: ro, ot["kernel"][:16]: ot}
out += ["# The first nine whole frames of the run (warm-up frames and the serial timing pass; one march + colour + finalize each):",
        "#   march    " + " ".join(f"{x:.4f}" for x in m) + f"  ms  mean {mm:.4f}",
        "#   colour   " + " ".join(f"{x:.4f}" for x in c) + f"  ms  mean {cm:.4f}",
        "#   finalize " + " ".join(f"{x:.4f}" for x in f) + f"  ms  mean {fm:.4f}",
        f"#   shares of the frame under ncu: march {mm / (mm + cm + fm):.3f}, colour {cm / (mm + cm + fm):.3f}, finalize {fm / (mm + cm + fm):.3f}"
        f"   (bench.py's CUDA-event shares of the serial single-stream step: march {km['ngf_march_kernel']['kernel_share_of_step']:.3f}, "
        f"colour {km['ngf_colour_kerne']['kernel_share_of_step']:.3f}; the rest is launch gaps + finalize)"]
open(os.path.join(P, "r02_launches_bench.txt"), "w").write("\n".join(out) + "\n")


# ---- ncu --set full summaries
def run(*a):
    return subprocess.run([sys.executable] + list(a), capture_output=True, text=True, cwd=ROOT).stdout


traffic = json.load(open(os.path.join(P, "traffic.json")))
for regime, target in (("hull", "hull 20"), ("dense", "dense 3"), ("infoinv", "infoinv 6")):
    if not os.path.exists(os.path.join(G, f"r2_prof_{regime}.ncu-rep")):
        continue
    rep = os.path.join(G, f"r2_prof_{regime}.ncu-rep")
    summ = run("scripts/ncu_summary.py", rep)
    txt = [f"# ncu --set full --clock-control none --import-source on, march + colour kernels of the {regime} regime, one frame of",
           f"# scripts/profile_target.py {target} (the bench's 16-pose rotation: rays come from HBM).  scripts/gpu_r2_prof.sh at the end",
           "# of round 2 (the binary the bench line was measured with); summary by scripts/ncu_summary.py, hot lines by",
           "# scripts/ncu_hot.py.  Times under ncu are cold-cache and serialised.", summ.rstrip(), ""]
    for k in ("ngf_march_kernel", "ngf_colour_kernel"):
        txt += [f"## {k}: hot source lines (share of stall samples / of warp instructions)", run("scripts/ncu_hot.py", rep, k, "14").rstrip(), ""]
    open(os.path.join(P, f"r02_ncu_{regime}_march_colour.txt"), "w").write("\n".join(txt) + "\n")
    cur, vals = None, {}
    for line in summ.splitlines():
        if line.startswith("=="):
            cur = "march" if "march" in line else "colour"
            vals[cur] = 0.0
        elif "dram__bytes_read.sum" in line or "dram__bytes_write.sum" in line:
            p = line.split()
            v, u = float(p[1]), p[2]
            vals[cur] += v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
    for k, v in vals.items():
        traffic[f"{k}_kernel_{regime}_dram_bytes_per_launch"] = int(v)
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(json.dumps({k: v for k, v in traffic.items() if k != "source"}, indent=1))
