#!/usr/bin/env python
"""Turn the scratch ncu outputs under gpurun_out/ into the committed summaries under profiles/.

    python profiles/summarize.py r1 gpurun_out/launches_r1.csv gpurun_out/prof_r1.ncu-rep

Writes profiles/<tag>_launches.md (per-kernel launch list: count, total, share, average),
profiles/<tag>_ncu_<kernel>.md (raw metrics + per-function stall samples of the --set full capture) and
updates profiles/traffic.json (dram bytes per launch of the persistent kernels, read by bench.py)."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

RAW_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def launches(tag, path):
    txt = open(path).read().splitlines()
    i = [k for k, l in enumerate(txt) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(io.StringIO("\n".join(txt[i:]))))
    agg = collections.OrderedDict()
    for r in rows:
        k = r["Kernel Name"]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    out = ["# %s: kernel launch list (ncu --metrics gpu__time_duration.sum --clock-control none)" % tag, "",
           "Command: `%s` (first launches up to the -c limit). Times are" % CMD_LAUNCH,
           "cold-cache and serialised under the profiler: compare SHARES, not absolutes.", "",
           "| kernel | launches | total us | share | avg us |", "|---|---:|---:|---:|---:|"]
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append("| `%s` | %d | %.1f | %.1f%% | %.1f |" % (k[:90], n, t / 1e3, 100 * t / tot, t / 1e3 / n))
    open(os.path.join(HERE, "%s_launches.md" % tag), "w").write("\n".join(out) + "\n")


def ncu_csv(rep, page, extra=()):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"] + list(extra), capture_output=True, text=True).stdout


def stall_groups(rep, kernel_regex):
    txt = ncu_csv(rep, "source", ["--print-source", "cuda,sass", "--kernel-name", "regex:" + kernel_regex])
    rows = list(csv.reader(io.StringIO(txt)))
    fpath, hdr = None, None
    per_line, tot = [], collections.Counter()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            col = {}
            for i, h in enumerate(hdr):
                col.setdefault(h, i)
            stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or r[0] == "":
            continue
        try:
            n = int(r[col["# Samples"]])
        except ValueError:
            continue
        c = {s: int(r[col[s]] or 0) for s in stalls}
        per_line.append((n, int(r[col["Instructions Executed"]] or 0), fpath, r[0], r[1].strip()[:90], c))
        tot["samples"] += n
        for s in stalls:
            tot[s] += c[s]
    return per_line, tot


def full(tag, rep):
    txt = ncu_csv(rep, "raw")
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    traffic_path = os.path.join(HERE, "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        short = "ppo_update" if "ppo" in name else "disc_update" if "disc" in name else name.split("(")[0]
        out = ["# %s: ncu --set full capture of `%s`" % (tag, name), "",
               "Command: `%s` (%s workload)." % (CMD_FULL, CFG), "", "| metric | value | unit |", "|---|---:|---|"]
        vals = {}
        for m in RAW_METRICS:
            if m in hdr:
                i = hdr.index(m)
                vals[m] = r[i]
                out.append("| %s | %s | %s |" % (m, r[i], units[i]))

        def num(m):
            return float(vals[m].replace(",", ""))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = num("dram__bytes_read.sum") * scale[units[hdr.index("dram__bytes_read.sum")]]
        wr = num("dram__bytes_write.sum") * scale[units[hdr.index("dram__bytes_write.sum")]]
        traffic.setdefault(CFG, {})[short] = rd + wr
        out += ["", "DRAM traffic per launch: %.3f MB read + %.3f MB written = %.3f MB." % (rd / 1e6, wr / 1e6, (rd + wr) / 1e6)]
        per_line, tot = stall_groups(rep, name.split("<")[0].split("(")[0].replace("void ", "").strip())
        if tot["samples"]:
            mix = sorted(((s, v) for s, v in tot.items() if s.startswith("stall_")), key=lambda x: -x[1])[:8]
            out += ["", "## Warp-stall sampling (%d samples)" % tot["samples"], "",
                    "Stall mix: " + ", ".join("%s %.1f%%" % (s, 100.0 * v / tot["samples"]) for s, v in mix), "",
                    "| samples | share | instructions | source line | top stalls |", "|---:|---:|---:|---|---|"]
            for n, inst, f, ln, src, c in sorted(per_line, key=lambda x: -x[0])[:30]:
                t2 = ", ".join("%s %d" % kv for kv in sorted(c.items(), key=lambda x: -x[1])[:2])
                out.append("| %d | %.1f%% | %d | `%s:%s` %s | %s |" % (n, 100.0 * n / tot["samples"], inst, f, ln,
                                                                   src.replace("|", "\\|"), t2))
        open(os.path.join(HERE, "%s_ncu_%s.md" % (tag, short)), "w").write("\n".join(out) + "\n")
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)


CFG = os.environ.get("SG_PROFILE_CFG", "cfg2")
CMD_LAUNCH = os.environ.get("SG_PROFILE_CMD_LAUNCH", "python bench.py --steps 2 --warmup 3 --no-cpu-baseline")
CMD_FULL = os.environ.get("SG_PROFILE_CMD_FULL", "ncu --set full --clock-control none --import-source on -k regex:\"disc_reg|ppo_persistent\" -s 4 -c 2 "
                          "python bench.py --steps 1 --warmup 3 --no-cpu-baseline")

if __name__ == "__main__":
    tag, launch_csv, rep = sys.argv[1:4]
    if os.path.exists(launch_csv):
        launches(tag, launch_csv)
    if os.path.exists(rep):
        full(tag, rep)
