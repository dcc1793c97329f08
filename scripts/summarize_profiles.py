"""Turn the ncu outputs a gpurun call left in gpurun_out/ into the tracked summaries under profiles/:
   <tag>_launches_raw.csv, <tag>_launch_list_summary.csv, <tag>_ncu_full_layers.csv (raw page),
   <tag>_ncu_full_summary.csv (selected metrics per captured layer), <tag>_dram_traffic_bytes.json.
Usage: python scripts/summarize_profiles.py r01e "layer names in capture order ..." """
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size"]
POSTPROC = ["peak_candidates_fused", "peak_nms", "keypoint_ids", "paf_pack", "paf_score", "limb_match", "pose_assemble"]


def main():
    tag = sys.argv[1]
    layers = sys.argv[2].split()
    # ---- launch list -------------------------------------------------------------------------
    raw = os.path.join(SRC, "launches.csv")
    rows = [r for r in csv.reader(open(raw)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = {}
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        name = r[ik].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", "")) / 1000.0  # ns -> us
    total = sum(v[1] for v in agg.values())
    with open(os.path.join(OUT, tag + "_launches_raw.csv"), "w") as f:
        f.write(open(raw).read())
    with open(os.path.join(OUT, tag + "_launch_list_summary.csv"), "w") as f:
        f.write("# ncu launch list (bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-roofline --no-u8; first 700 launches)\n")
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 700  (cold-cache, serialised: compare SHARES)\n")
        f.write("kernel,launches,total_us,share\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%s,%d,%.1f,%.3f\n" % (k.replace(",", ";"), v[0], v[1], v[1] / total))
    # ---- full capture ------------------------------------------------------------------------
    rep = os.path.join(SRC, "prof_layers.ncu-rep")
    rawcsv = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    with open(os.path.join(OUT, tag + "_ncu_full_layers.csv"), "w") as f:
        f.write(rawcsv)
    rr = list(csv.reader(rawcsv.splitlines()))
    h = rr[0]
    idx = {n: i for i, n in enumerate(h)}
    names = []
    for nm in layers:
        names += POSTPROC if nm == "postproc" else [nm]
    data = rr[2:]
    traffic = {}
    with open(os.path.join(OUT, tag + "_ncu_full_summary.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["layer", "Kernel Name"] + KEEP)
        for nm, r in zip(names, data):
            kn = r[idx["Kernel Name"]]
            w.writerow([nm, kn] + [r[idx[k]] if k in idx else "" for k in KEEP])
            rd, wr = float(r[idx["dram__bytes_read.sum"]]), float(r[idx["dram__bytes_write.sum"]])
            unit_r, unit_w = rr[1][idx["dram__bytes_read.sum"]], rr[1][idx["dram__bytes_write.sum"]]
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            short = kn.split("(")[0].replace("void ", "").replace("lwp::", "")
            traffic.setdefault(short, {})[nm] = rd * mult[unit_r] + wr * mult[unit_w]
    with open(os.path.join(OUT, tag + "_dram_traffic_bytes.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    print("wrote", tag, "launch kinds:", len(agg), "captured kernels:", len(data), "named:", len(names))


if __name__ == "__main__":
    main()
