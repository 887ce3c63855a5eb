#!/usr/bin/env python
"""Summarise ncu outputs brought back from the GPU box (read here, no GPU needed).
  python tools/ncu_summary.py launches <launches.csv>          per-kernel totals / shares of a launch list
  python tools/ncu_summary.py raw <report.ncu-rep>             key metrics of every profiled launch
  python tools/ncu_summary.py stalls <report.ncu-rep> <regex>  top stall instructions of one kernel"""
import collections
import csv
import io
import subprocess
import sys

KEY = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
       'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
       'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
       'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
       'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
       'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'smsp__inst_executed.sum']


def launches(path):
    lines = [l for l in open(path).read().splitlines() if l.startswith('"')]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(io.StringIO("\n".join(lines))):
        if row['Metric Name'] != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}[row['Metric Unit']]
        a = agg.setdefault(row['Kernel Name'].split('(')[0], [0, 0.0]); a[0] += 1; a[1] += v; tot += v
    print("total %.3f ms over %d launches" % (tot, sum(a[0] for a in agg.values())))
    for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1])[:25]:
        print("%-62s %4d launches %9.3f ms  %5.1f%%" % (k[:62], n, ms, 100 * ms / tot))


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    seen = set()
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')].split('(')[0]
        if name in seen:
            continue
        seen.add(name)
        print('---', name)
        for w in KEY:
            if w in hdr:
                print("  %-66s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))


def stalls(path, regex, top=18, which=0):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "-k", "regex:" + regex.split("<")[0]],
                         capture_output=True, text=True).stdout
    allrows = list(csv.reader(io.StringIO(out)))
    # split into launches; keep the `which`-th whose full name contains `regex`
    blocks, cur = [], None
    for r in allrows:
        if r and r[0] == "Kernel Name":
            cur = [r]; blocks.append(cur)
        elif cur is not None:
            cur.append(r)
    blocks = [b for b in blocks if regex.replace(" ", "") in b[0][1].replace("(int)", "").replace(" ", "")]
    rows = blocks[int(which)]
    hdr = rows[1]
    body = rows[2:]
    top = int(top)
    i_src, i_s, i_ex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    cols = [(j, h) for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    body = [r for r in body if len(r) > i_s and r[i_s].isdigit()]
    tot = sum(int(r[i_s]) for r in body) or 1
    print(rows[0][1][:100], "samples", tot)
    agg = collections.Counter()
    for r in body:
        for j, h in cols:
            if r[j].isdigit():
                agg[h] += int(r[j])
    print("  by reason:", ", ".join("%s %.0f%%" % (h[6:], 100 * v / tot) for h, v in agg.most_common(6)))
    for r in sorted(body, key=lambda r: -int(r[i_s]))[:top]:
        st = sorted(((int(r[j]), h[6:]) for j, h in cols if r[j].isdigit() and int(r[j]) > 0), reverse=True)[:2]
        print("  %5.1f%% ex=%9s  %-58s %s" % (100 * int(r[i_s]) / tot, r[i_ex], r[i_src].strip()[:58], st))


NAMES = {"rowsum_kernel": "rowsum_items", "dots_units_kernel": "dots", "tile_lm_sweep_kernel<1, 5, 256>": "lm_sweep_hv",
         "tile_prepare_kernel<5, 256>": "tile_prepare"}


def traffic(full_rep, iter_csv, out_json):
    """profiles/r02_traffic.json: per-launch DRAM bytes + L2 counters of the hot kernels (from ONE `ncu --set full` report) and
    the DRAM bytes of a whole outer iteration (from a dram-bytes launch list of tools/profile_step.py)."""
    import json
    out = subprocess.run(["ncu", "-i", full_rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = lambda r, name: float(r[hdr.index(name)].replace(",", ""))
    scale = lambda name: {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[hdr.index(name)]]
    kern = {}
    for r in rows[2:]:
        full = r[hdr.index("Kernel Name")]
        # the item-major row sum is the rowsum_kernel launch that takes the L2-hint instantiation (or, without it, the first one)
        key = next((v for k, v in NAMES.items() if k in full), None)
        if key is None:
            continue
        if key == "rowsum_items" and "true" not in full and "rowsum_items" in kern:
            continue
        if key in kern and not (key == "rowsum_items" and "true" in full):
            continue
        rd = col(r, "dram__bytes_read.sum") * scale("dram__bytes_read.sum")
        wr = col(r, "dram__bytes_write.sum") * scale("dram__bytes_write.sum")
        kern[key] = {"kernel": full.split("(")[0], "dram_bytes_per_launch": rd + wr, "dram_read_bytes_per_launch": rd,
                     "lts_throughput_pct": col(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                     "l2_hit_pct": col(r, "lts__t_sector_hit_rate.pct"), "l1tex_hit_pct": col(r, "l1tex__t_sector_hit_rate.pct"),
                     "duration_ms_under_ncu": col(r, "gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[units[hdr.index("gpu__time_duration.sum")]]}
    it = {"dram_bytes": 0.0, "launches": 0, "time_ms_under_ncu": 0.0}
    lines = [l for l in open(iter_csv).read().splitlines() if l.startswith('"')]
    for row in csv.DictReader(io.StringIO("\n".join(lines))):
        v = float(row["Metric Value"].replace(",", ""))
        if row["Metric Name"].startswith("dram__bytes"):
            it["dram_bytes"] += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[row["Metric Unit"]]
        elif row["Metric Name"] == "gpu__time_duration.sum":
            it["launches"] += 1; it["time_ms_under_ncu"] += v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
    json.dump({"_note": "ncu of the round-2 tree, Netflix-shape k=100 at full size on one B200: kernels = one `ncu --set full "
                        "--clock-control none` launch each (tools/gpu/r02_capture.sh -> %s); iteration = dram__bytes_read+write "
                        "summed over every launch of ONE outer iteration (%s)" % (full_rep.split("/")[-1], iter_csv.split("/")[-1]),
               "workload": {"shape": "netflix", "scale": 1.0, "k": 100, "n_gpus": 1}, "kernels": kern, "iteration": it},
              open(out_json, "w"), indent=1)
    print(json.dumps({"kernels": kern, "iteration": it}, indent=1))


if __name__ == "__main__":
    {"launches": launches, "raw": raw, "stalls": stalls, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
