// driver.hpp -- host-side driver shared by the drop-in shim (shim/pcr_shim.cpp) and our own CLI (cli/pmf_train.cpp):
// runs pcr()/pcrpp() on flat CSR + flat factors through the C ABI, sharding users over PRIMALCR_GPUS GPUs of the box
// (one host thread + one engine per GPU, NCCL allreduce of the V-side sums).
#pragma once
#include "primalcr.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <thread>
#include <vector>

namespace pcrhost {

struct FlatCsr { int64_t d1, d2, nnz; const int64_t *row_ptr; const int32_t *item; const double *rating; };

inline void die(const char *what, int rc) {
    fprintf(stderr, "primalcr_b200: %s failed (%d): %s\n", what, rc, primalcr_last_error());
    exit(1);
}
#define PCRHOST_CK(call) do { int _rc = (call); if (_rc != 0) ::pcrhost::die(#call, _rc); } while (0)

inline void log_line(const char *line, void *) { std::cout << line << std::endl; }

// wall-clock laps of the host side, printed to stderr when PRIMALCR_VERBOSE_SETUP is set
struct Lap {
    bool on; std::chrono::steady_clock::time_point t;
    Lap() : on(getenv("PRIMALCR_VERBOSE_SETUP") != nullptr), t(std::chrono::steady_clock::now()) {}
    void operator()(const char *what) {
        if (!on) return;
        const auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[primalcr host] %-32s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

// find_levels pcrpp.cpp:38-49 as ONE global ascending table of distinct lround(rating).  One pass over the ratings on all
// host threads (each keeps its own small sorted table, merged at the end): 100 M ratings take ~0.1 s instead of ~1 s.
inline std::vector<int64_t> global_levels(const FlatCsr &X) {
    std::vector<int64_t> lv;
#pragma omp parallel
    {
        std::vector<int64_t> mine;
        int64_t last = 0; bool have_last = false;
#pragma omp for schedule(static) nowait
        for (int64_t e = 0; e < X.nnz; ++e) {
            const int64_t l = llround(X.rating[e]);
            if (have_last && l == last) continue;
            last = l; have_last = true;
            auto it = std::lower_bound(mine.begin(), mine.end(), l);
            if (it == mine.end() || *it != l) mine.insert(it, l);
        }
#pragma omp critical
        for (int64_t l : mine) {
            auto it = std::lower_bound(lv.begin(), lv.end(), l);
            if (it == lv.end() || *it != l) lv.insert(it, l);
        }
    }
    if (lv.empty()) lv.push_back(0);
    return lv;
}

inline int gpus_from_env() {
    int gpus = 1;
    if (const char *g = getenv("PRIMALCR_GPUS")) gpus = std::max(1, atoi(g));
    return gpus;
}

// U (d1 x k) and V (d2 x k) row-major, updated in place -- like the reference's mat_t& U, mat_t& V
inline void solve(const primalcr_config &base, const FlatCsr &X, const FlatCsr &XT, double *U, double *V, int gpus) {
    const int k = base.k;
    Lap lap;
    // Primal-CR (-s 1) compares the exact ratings (pcr.cpp:23): no level table, any rating scale
    const std::vector<int64_t> levels = base.solver == PRIMALCR_SOLVER_PCRPP ? global_levels(X) : std::vector<int64_t>();
    std::vector<int64_t> bounds(gpus + 1, 0);            // contiguous user shards balanced by nnz
    for (int r = 1; r < gpus; ++r) {
        const int64_t target = (int64_t)((double)X.nnz * r / gpus);
        bounds[r] = std::lower_bound(X.row_ptr, X.row_ptr + X.d1 + 1, target) - X.row_ptr;
        bounds[r] = std::min<int64_t>(std::max(bounds[r], bounds[r - 1]), X.d1);
    }
    bounds[gpus] = X.d1;
    lap("level table + shard bounds");
    char uid[128] = {0};
    if (gpus > 1) PCRHOST_CK(primalcr_nccl_unique_id(uid));
    auto worker = [&](int rank) {
        primalcr_config cfg = base;
        cfg.device = rank;
        primalcr_engine *e = nullptr;
        Lap wl;
        PCRHOST_CK(primalcr_create(&e, &cfg));
        if (rank == 0) wl("create engine (CUDA context)");
        if (!levels.empty()) PCRHOST_CK(primalcr_set_levels(e, levels.data(), (int)levels.size()));
        PCRHOST_CK(primalcr_comm_init(e, rank, gpus, gpus > 1 ? uid : nullptr));
        const int64_t u0 = bounds[rank], u1 = bounds[rank + 1];
        auto shard = [&](const FlatCsr &F, std::vector<int64_t> &rp) {
            rp.assign(F.row_ptr + u0, F.row_ptr + u1 + 1);
            const int64_t b = rp[0];
            for (auto &x : rp) x -= b;
            return b;
        };
        std::vector<int64_t> rp, rpt;
        const int64_t b0 = shard(X, rp);
        PCRHOST_CK(primalcr_set_train_csr(e, u1 - u0, X.d2, rp.back(), rp.data(), X.item + b0, X.rating + b0));
        if (XT.nnz != 0) {
            const int64_t bt = shard(XT, rpt);
            PCRHOST_CK(primalcr_set_test_csr(e, rpt.back(), rpt.data(), XT.item + bt, XT.rating + bt));
        }
        if (rank == 0) wl("comm init + set_train/test");
        PCRHOST_CK(primalcr_set_factors(e, U + (size_t)u0 * k, V));
        if (rank == 0) wl("set_factors");
        PCRHOST_CK(primalcr_run(e, log_line, nullptr));
        if (rank == 0) wl("run (all iterations)");
        PCRHOST_CK(primalcr_get_factors(e, U + (size_t)u0 * k, rank == 0 ? V : nullptr));
        if (rank == 0) wl("get_factors");
        primalcr_destroy(e);
        if (rank == 0) wl("destroy engine");
    };
    if (gpus == 1) worker(0);
    else {
        std::vector<std::thread> th;
        for (int r = 0; r < gpus; ++r) th.emplace_back(worker, r);
        for (auto &t : th) t.join();
    }
}

}  // namespace pcrhost
