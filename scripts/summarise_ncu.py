"""Turns ncu output (run on the GPU box, read here without a GPU) into the small summaries committed under profiles/.

    python scripts/summarise_ncu.py full gpurun_out/x.ncu-rep profiles/out.json     # one record per profiled launch
    python scripts/summarise_ncu.py launches gpurun_out/launches.csv profiles/out.json
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__waves_per_multiprocessor",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return re.sub(r".*::", "", name).replace("void ", "")


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    recs = []
    for r in rows[2:]:
        rec = {"kernel": short(r[hdr.index("Kernel Name")])}
        for k in KEEP:
            if k in hdr:
                v = r[hdr.index(k)].replace(",", "")
                try:
                    rec[f"{k} [{units[hdr.index(k)]}]"] = float(v)
                except ValueError:
                    rec[k] = v
        rd, wr = rec.get("dram__bytes_read.sum [Gbyte]"), rec.get("dram__bytes_write.sum [Gbyte]")
        t = rec.get("gpu__time_duration.sum [us]") or (rec.get("gpu__time_duration.sum [ms]", 0) * 1e3)
        if rd is not None and wr is not None and t:
            rec["dram_GBps_under_ncu"] = round((rd + wr) * 1e3 / (t * 1e-6) / 1e3, 1)
        recs.append(rec)
    json.dump({"source": rep, "note": "ncu --set full --clock-control none; per-launch values (cold caches, serialised)",
               "launches": recs}, open(out, "w"), indent=1)
    print(f"{len(recs)} launches -> {out}")


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in agg.values())
    table = [{"kernel": k, "launches": c, "total_us": round(v / 1e3, 1), "share_pct": round(100 * v / tot, 2)}
             for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    json.dump({"source": path, "note": "ncu --metrics gpu__time_duration.sum --clock-control none: cold-cache, serialised "
               "per-launch times; compare SHARES, not absolutes", "total_us": round(tot / 1e3, 1), "kernels": table},
              open(out, "w"), indent=1)
    print(f"{sum(c for c, _ in agg.values())} launches -> {out}")


if __name__ == "__main__":
    {"full": full, "launches": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
