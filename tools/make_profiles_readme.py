#!/usr/bin/env python
"""Regenerates profiles/README.md from the committed bench JSON lines and ncu text summaries."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = lambda *a: os.path.join(ROOT, "profiles", *a)


def main():
    d = json.load(open(P("r01_bench_netflix_k100_1gpu.json")))
    ref = json.load(open(P("r01_bench_reference_arm.json")))
    scal = {}
    for n in (2, 4, 8):
        if os.path.exists(P("r01_bench_netflix_k100_%dgpu.json" % n)):
            scal[n] = json.load(open(P("r01_bench_netflix_k100_%dgpu.json" % n)))
    out = []
    out.append("# profiles/ — round 1 evidence\n")
    out.append("All numbers: B200 (sm_100a, 148 SMs, SM clock 1965 MHz under load, throttle reasons as recorded in each JSON),\nNetflix-shape synthetic (480,189 users x 17,770 items, 100,000,003 ratings 1-5), Primal-CR++ `-s 2 -k 100 -l 5000`,\nreference init, fp64.  Regenerate this file with `python tools/make_profiles_readme.py`.\n")
    out.append("## Headline (`r01_bench_netflix_k100_1gpu.json` = `python bench.py --steps 3 --warmup 3`)\n")
    out.append("| quantity | value |\n|---|---|")
    out.append("| seconds per outer iteration, device-resident (`value`) | **%.4f s** |" % d["value"])
    out.append("| seconds per outer iteration, end to end (`e2e`: upload CSR+U+V, build CSC / work lists, 3 iterations, download U+V, divided by 3) | %.3f s |" % d["e2e"]["value"])
    cb = d.get("cpu_baseline") or {}
    out.append("| reference `omp-pmf-train -n %s` on the same box (oracle/_ref, 2.0 M-rating sample x50), in-line `cpu_baseline` / separate `--impl reference` run | %.1f s / %.1f s |" % (cb.get("cores"), cb.get("value", float("nan")), ref["value"]))
    it = d["roofline"]["iteration"]
    out.append("| algorithmic bytes per iteration `B_alg` (SURVEY 8d; %.1f N*k passes, %.1f sort+sweep passes) | %.2f TB |" % (it["counters"]["passes"], it["counters"]["sorts"], it["b_alg_bytes"] / 1e12))
    out.append("| `B_alg / t` vs measured HBM peak %.0f GB/s | %.0f GB/s = **%.2f x peak** (target was >= 0.60) |" % (d["roofline"]["peak"], it["achieved"], it["frac"]))
    out.append("| first correct path of this round (`r01_bench_netflix_k100_1gpu_v1_first_path.json`) | 0.652 s |")
    out.append("| kernels launched per timed iteration (`gpu_launches` / steps) | %d |" % (d["gpu_launches"] // d["steps"]))
    out.append("\n`B_alg / t` exceeds the HBM peak because the logical k-vector touches counted by `B_alg` are mostly served on chip:\nV (14 MB) is L2-resident for the user-major gathers, and the item-major row-sum walks U in 24 MB user blocks, so its\ngathered rows hit L2 as well.  `roofline.traffic` (ncu DRAM bytes per launch, `r01_traffic.json`) shows the real HBM\ntraffic: 8.2 GB per item-major row-sum launch against 81.6 GB algorithmic, 1.6 GB per `dots` launch.\n")
    out.append("## Per-kernel table (CUDA events inside the timed region)\n")
    out.append("| kernel | ms/step | launches/step | algorithmic GB/s |\n|---|---|---|---|")
    for k in d["roofline"]["kernels"]:
        out.append("| %s | %.2f | %.0f | %s |" % (k["name"], k["ms_per_step"], k["launches_per_step"], ("%.0f" % k["gbs"]) if k["gbs"] else "-"))
    out.append("")
    if os.path.exists(P("r01_ncu_full_size_summary.txt")):
        out.append("## ncu `--set full` at FULL size (one outer iteration; `.ncu-rep` not committed, 60 MB)\n")
        out.append("Command: `ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:\"rowsum_kernel|dots_units_kernel|tile_lm_sweep_kernel|tile_prepare_kernel\" -s 3 -c 9 python tools/profile_step.py --scale 1.0`\n(capture taken before the 128-byte row alignment: `dots` was 4.46 ms there, 4.15 ms now).\n")
        out.append("```\n" + open(P("r01_ncu_full_size_summary.txt")).read().strip() + "\n```\n")
        out.append("Reading:\n* `dots_units_kernel<8,7>` (81.6 GB algorithmic per launch): 1.58 GB of DRAM traffic, L2 hit 97 %, **l1tex throughput 98 %, L2 throughput 73 %** -> bound by the L1/L2 data path, not by HBM.\n* `rowsum_kernel<2>` item-major (81.6 GB algorithmic): 8.2 GB DRAM (U once + the 8-byte coefficient gather + indices + partial sums), L2 hit 85 %, 12.3 TB/s of L2->SM traffic; stall reason `long_scoreboard` 73 %.  Neither more loads in flight per lane (occupancy loss) nor TMA bulk row staging (per-request cost, 11.4 ms vs 6.6 ms) helped -- see DESIGN.md section 4.\n* `tile_lm_sweep_kernel<1,..>` (Hv sweep, 41 B/rating): 3.5 GB DRAM in 1.5 ms = 2.3 TB/s; issue-bound (scalar segmented scan + 2(T-1) shared-memory look-ups per rating).\n* `tile_prepare_kernel` (stall sampling, scale 0.3): short_scoreboard 32 % (shared-memory latency in the bitonic network), wait 24 %, barrier 12 %.\n")
    out.append("## Launch lists (`--metrics gpu__time_duration.sum --clock-control none`)\n")
    out.append("* `r01_ncu_launches_bench_netflix_k100.csv` — the bench command itself (`python bench.py --steps 1 --warmup 3 --no-cpu-baseline`, full size). Shares over all `pcr::` launches: rowsum_kernel 36.8 % (CUDA events: 37.6 %), dots_units 27.7 % (27.9 %), tile_prepare 12.5 % (10.3 %), lm_sweep hv 12.3 % (13.1 %): the kernel shares agree with the CUDA-event table.\n* `r01_ncu_launches_netflix0.2_k100.csv`, `r01_ncu_top_kernels_v1.md` — the FIRST correct path (v1, 0.652 s/iter) at scale 0.2, kept to show where the optimisation started (sweep: 37 warp-instructions per rating; grids sized past occupancy).\n")
    out.append("## Multi-GPU (strong scaling, same data set, users sharded by nnz; `r01_bench_netflix_k100_{2,4,8}gpu.json`)\n")
    out.append("| GPUs | s / outer iteration | parallel efficiency t1/(n tn) | e2e s / iteration | objective after 6 iterations |\n|---|---|---|---|---|")
    out.append("| 1 | %.4f | 1.00 | %.3f | %.15g |" % (d["value"], d["e2e"]["value"], d["objective"][-1]))
    for n, s in sorted(scal.items()):
        out.append("| %d | %.4f | %.2f | %.3f | %.15g |" % (n, s["value"], d["value"] / (n * s["value"]), s["e2e"]["value"], s["objective"][-1]))
    out.append("\nThe e2e column at N > 1 in these files was taken before the per-process NCCL communicator cache (each solver call\npaid `ncclCommInitRank`, 1-3 s); the driver's own scaling run is authoritative.\n")
    for fn, title in (("r01_other_shapes_1gpu.jsonl", "Other BASELINE.json shapes, scaled, 1 GPU"),
                      ("r01_other_shapes_full_size_1gpu.jsonl", "Yahoo-shape and power-law shape at FULL size, 1 GPU")):
        if not os.path.exists(P(fn)):
            continue
        out.append("## %s (`%s`, `tools/run_shapes.py`)\n" % (title, fn))
        out.append("| shape | solver | k | ratings | device GB | s / outer iteration | checks |\n|---|---|---|---|---|---|---|")
        for l in open(P(fn)):
            o = json.loads(l)
            out.append("| %s x%g (max len %d) | %s | %d | %d | %.1f | %.4f | monotone=%s, recomputed objective rel.err %.1e |" % (
                o["shape"], o["scale"], o["max_len"], "Primal-CR++" if o["solver"] == 2 else "Primal-CR", o["k"], o["nnz"],
                o["device_gb"], o["sec_per_iter"][-1], o["monotone"], o["recomputed_rel_err"]))
        out.append("")
    open(P("README.md"), "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
