"""Turns ncu outputs brought back from the GPU box into the committed summaries under profiles/ (development aid).

  python tools/ncu_summary.py launches <launch-list.csv> <tag>      -> profiles/ncu_<tag>_launches.csv (copy) and
                                                                       profiles/ncu_<tag>_launches_per_cycle.md
  python tools/ncu_summary.py raw <tag> <capture.ncu-rep> [...]     -> profiles/ncu_<tag>_raw_key_metrics.csv
"""
import collections
import csv
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]


def launches(path, tag):
    dst = os.path.join(ROOT, "profiles", f"ncu_{tag}_launches.csv")
    shutil.copyfile(path, dst)
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = None
    seq = []
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            rec = dict(zip(hdr, r))
            if rec.get("Metric Name") == "gpu__time_duration.sum":
                v = float(rec["Metric Value"].replace(",", ""))
                u = rec["Metric Unit"]
                v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u.split("e")[0] if u in ("ns", "us", "ms", "s") else u, 1e-6)
                seq.append((rec["Kernel Name"], v))
    # one SCF cycle = from one symm_panel launch to the next; take the LAST complete cycle (steady state)
    idx = [i for i, (k, _) in enumerate(seq) if "symm_panel_kernel" in k]
    out = [f"# ncu launch list `{os.path.basename(dst)}`: {len(seq)} launches, {len(idx)} pass-1 launches\n"]
    if len(idx) >= 2:
        a, b = idx[-2], idx[-1]
        agg = collections.OrderedDict()
        for k, v in seq[a:b]:
            name = k.split("(")[0].replace("void ", "").replace("nbd::", "")
            n, t = agg.get(name, (0, 0.0))
            agg[name] = (n + 1, t + v)
        tot = sum(t for _, t in agg.values())
        out.append("One SCF cycle (between the last two pass-1 launches; `gpu__time_duration`, serialised, cold cache - the "
                   "SHARE is what compares with the event-timed stages of `bench.py`, not the absolute time):\n")
        out.append("| kernel | launches | ms | share |\n|---|---|---|---|")
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.append(f"| `{name}` | {n} | {t:.3f} | {100 * t / tot:.1f}% |")
        out.append(f"| total | {sum(n for n, _ in agg.values())} | {tot:.3f} | |")
    md = os.path.join(ROOT, "profiles", f"ncu_{tag}_launches_per_cycle.md")
    open(md, "w").write("\n".join(out) + "\n")
    print(open(md).read())


def raw(tag, reps):
    dst = os.path.join(ROOT, "profiles", f"ncu_{tag}_raw_key_metrics.csv")
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["capture", "kernel"] + KEYS)
        for rep in reps:
            txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
            rows = list(csv.reader(txt.splitlines()))
            hdr, units = rows[0], rows[1]
            for vals in rows[2:]:
                d = dict(zip(hdr, vals))
                u = dict(zip(hdr, units))
                w.writerow([os.path.basename(rep), d.get("Kernel Name", "")] + [f"{d.get(k, '')} {u.get(k, '')}".strip() for k in KEYS])
    print(open(dst).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        raw(sys.argv[2], sys.argv[3:])
