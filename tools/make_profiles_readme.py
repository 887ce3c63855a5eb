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
    out.append("All numbers: B200 (sm_100a, 148 SMs; SM clocks and throttle reasons as sampled during each timed region, `clocks` key),\nNetflix-shape synthetic (480,189 users x 17,770 items, 100,000,003 ratings 1-5), Primal-CR++ `-s 2 -k 100 -l 5000`,\nreference init, fp64.  Regenerate this file with `python tools/make_profiles_readme.py`.\n")
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
    out.append("| SM clock while timed (median / max; throttle reasons) | %s / %s MHz; %s |" % (d["clocks"]["sm_mhz"], d["clocks"]["sm_max_mhz"], ", ".join(d["clocks"]["reasons"]) or "none"))
    ph = d["e2e"].get("phases_s")
    if ph:
        out.append("| e2e phases (s): engine create / set_train (upload + CSC + work lists) / set_factors / run (initial objective + %d iterations) / get_factors | %s |" % (d["steps"], " / ".join("%.3f" % ph[k_] for k_ in ("create", "set_train", "set_factors", "run", "get_factors"))))
    out.append("| kernels launched per timed iteration (`gpu_launches` / steps) | %d |" % (d["gpu_launches"] // d["steps"]))
    out.append("\n`B_alg / t` exceeds the HBM peak because the logical k-vector touches counted by `B_alg` are mostly served on chip:\nV (14 MB) is L2-resident for the user-major gathers, and the item-major row-sum walks U in 24 MB user blocks, so its\ngathered rows hit L2 as well.  `roofline.traffic` (ncu DRAM bytes per launch, `r01_traffic.json`) shows the real HBM\ntraffic: 9.0 GB per item-major row-sum launch against 81.2 GB algorithmic, 1.6 GB per `dots` launch.\nThe timed iterations (4-6 from the N(0,1) init) run 7.1 len-weighted U-side CG rounds; `tools/ab.py` (iterations 3-4, 5.8\nrounds, no per-launch events) reads 0.229 s for the same build.\n")
    out.append("## Per-kernel table (CUDA events inside the timed region)\n")
    out.append("| kernel | ms/step | launches/step | algorithmic GB/s |\n|---|---|---|---|")
    for k in d["roofline"]["kernels"]:
        out.append("| %s | %.2f | %.0f | %s |" % (k["name"], k["ms_per_step"], k["launches_per_step"], ("%.0f" % k["gbs"]) if k["gbs"] else "-"))
    out.append("")
    if os.path.exists(P("r01_ncu_full_size_summary.txt")):
        out.append("## ncu `--set full` at FULL size (one outer iteration; `.ncu-rep` not committed, 60 MB)\n")
        out.append("Commands (each after `python tools/profile_step.py --scale 1.0` had exited 0 without ncu): `ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:\"rowsum_kernel|dots_units_kernel|tile_lm_sweep_kernel\" -s 4 -c 5 python tools/profile_step.py --scale 1.0`; the same with `-k regex:\"tile_prepare_kernel\" -c 3` and `-k regex:\"hv_chunk|hv_lookup|rowsum_finalize|u_cg_step\" -s 3 -c 6`.\n")
        out.append("```\n" + open(P("r01_ncu_full_size_summary.txt")).read().strip() + "\n```\n")
        out.append("Reading:\n* `dots_units_kernel<8,7>` (81.2 GB algorithmic per launch, 4.18 ms = 19.4 TB/s): 1.6 GB of DRAM traffic, L2 hit 97 %, l1tex throughput 85 %, L2 throughput 75 % (busiest slice 92 %) -> bound by the L2 -> SM data path and load latency at a full register file, not by HBM.  Fewer shuffles, L1 / L2 cache-policy hints, a shared-memory copy of the hottest rows and 256-bit loads were all measured and changed nothing or lost (profiles/experiments/README.md).\n* `rowsum_kernel<2>` item-major (81.2 GB algorithmic, 5.82 ms = 14 TB/s): 9.0 GB DRAM (U once + the 8-byte coefficient gather through csc2csr + indices + partial sums), L2 hit 83 %, stall reason `long_scoreboard`.\n* `tile_lm_sweep_kernel<1,5,256>` (Hv sweep, 32 B/rating with the packed records): 2.7 GB DRAM in 1.12 ms = 2.4 TB/s, 5 CTAs/SM (48 registers), issue 56 %: latency of the load phase + scalar segmented scan.  `<1,5,512>` / `<1,5,1024>` are the medium (1024 < len <= 2048, two CTAs/SM) and large (<= 4096) tile geometries: 0.21 + 0.09 ms, against 0.47 ms when both shared the 4096-rating tiles (39 % full).\n* `tile_prepare_kernel<5,256>` (4.45 ms; 4.93 before the single-compare exchange): issue slots ~74 % busy -- instruction-bound (64-bit compare-exchange network in registers / shuffles, window binary searches, 5-level count scan); DRAM 5.5 GB per launch (writes dominate: sorted outputs + level-major records).\n* heavy users (`hv_*`, 621 chunks of 2048 ratings): sums 10 us + scan 20 us + look-ups 42 us per sweep on the high-priority side stream, against 144 us for the one-CTA-per-user kernel they replace.\n")
    out.append("## Launch lists (`--metrics gpu__time_duration.sum --clock-control none`)\n")
    out.append("* `r01_ncu_launches_bench_netflix_k100.csv` — the bench command itself (`python bench.py --steps 1 --warmup 3 --no-cpu-baseline`, full size; the list also holds torch's data-generation kernels, which run before the timed region). Shares over all `pcr::` launches vs the CUDA-event table above: rowsum_kernel 41.0 % (events: 42.3 %), dots_units 33.0 % (33.9 %), lm_sweep Hv 10.5 % (11.5 %), tile_prepare 6.3 % (5.2 %): the kernel shares agree.\n* `r01_ncu_launches_netflix0.2_k100.csv`, `r01_ncu_top_kernels_v1.md` — the FIRST correct path (v1, 0.652 s/iter) at scale 0.2, kept to show where the optimisation started (sweep: 37 warp-instructions per rating; grids sized past occupancy).\n* `experiments/` — measured-and-rejected variants (patches + result tables).\n* `r01_trace_one_iteration.txt` — per-launch CUDA-event trace of one outer iteration (`PRIMALCR_TRACE`): 232 launches on the main stream, busy 99.5 % of the 226 ms span (1.2 ms of gaps in total, none above 44 us); the heavy-user kernels run beside them on the side stream.\n")
    out.append("## Multi-GPU (strong scaling, same data set, users sharded by nnz; `r01_bench_netflix_k100_{2,4,8}gpu.json`)\n")
    out.append("| GPUs | s / outer iteration | parallel efficiency t1/(n tn) | e2e s / iteration | objective after 6 iterations |\n|---|---|---|---|---|")
    out.append("| 1 | %.4f | 1.00 | %.3f | %.15g |" % (d["value"], d["e2e"]["value"], d["objective"][-1]))
    for n, s in sorted(scal.items()):
        out.append("| %d | %.4f | %.2f | %.3f | %.15g |" % (n, s["value"], d["value"] / (n * s["value"]), s["e2e"]["value"], s["objective"][-1]))
    out.append("\nEvery rank's kernel totals are in the `per_rank` key of the N > 1 files (user shards balanced by nnz: all kernels within 0.5 %\nacross ranks at N = 8).  e2e at N > 1 re-attaches to the process's cached NCCL communicator and allocates from the\nstream-ordered pool (plain cudaMalloc is ~10x slower once NCCL has enabled peer access: 0.50 -> 0.13 s/iter at N = 2).\nThe N = 2/4/8 files predate the pitched factor download: the N = 4 e2e still holds 0.30 s of `get_factors` (its staging\nbuffer made the pool grow; at N = 1 the same fix took get_factors from 0.097 to 0.008 s).\n")
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
    if os.path.exists(P("r01_bench_yahoo_k100_4gpu.json")):
        y = json.load(open(P("r01_bench_yahoo_k100_4gpu.json")))
        nc = [k_ for k_ in y["roofline"]["kernels"] if k_["name"] == "nccl_allreduce"]
        out.append("## Yahoo shape at full size on 4 GPUs (`r01_bench_yahoo_k100_4gpu.json`, `bench.py --gpus 4 --workload yahoo --steps 2 --warmup 1`)\n")
        out.append("%.4f s per outer iteration (1 GPU: 0.929 s -> %.0f %% parallel efficiency), e2e %.3f s; the %d all-reduces of the 500 MB V-side\nvectors take %.1f ms per iteration (NCCL over NVLink 5).\n" % (
            y["value"], 100 * 0.9287 / (4 * y["value"]), y["e2e"]["value"], int(nc[0]["launches_per_step"]) if nc else 0, nc[0]["ms_per_step"] if nc else float("nan")))
    open(P("README.md"), "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
