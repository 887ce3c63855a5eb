// pcr_shim.cpp -- drop-in replacements for the reference's solver entry points
//
//     extern "C" void pcr  (smat_t &X, mat_t &U, mat_t &V, testset_t &T, parameter &param);   pmf.h:54, pcr.cpp:616
//     extern "C" void pcrpp(smat_t &X, mat_t &U, mat_t &V, testset_t &T, parameter &param);   pmf.h:55, pcrpp.cpp:841
//
// Compile this ONE file inside the reference tree (it includes the reference's own util.h / pmf.h) and link it,
// together with libprimalcr_b200.so, in place of pcr.o and pcrpp.o (reference Makefile:9-13).  pmf-train.cpp,
// util.cpp, the data formats and the model file stay untouched.  See INTEGRATION.md.
//
// What it does: flattens smat_t's row-major arrays (row_ptr / col_idx / val_t, util.h:157-271) and the testset_t
// (util.h:354-386, grouped like convert(testset_t&) util.cpp:250-274) into plain CSR buffers, flattens mat_t U, V,
// calls the C ABI (include/primalcr.h), prints the reference's log lines and copies U, V back.
// PRIMALCR_GPUS=N (default 1) shards the users over N GPUs of the box: one host thread + one engine per GPU.
#include "util.h"
#include "pmf.h"
#include "primalcr.h"

#include <thread>

namespace {

struct Flat {
    long d1, d2, nnz;
    std::vector<int64_t> row_ptr;
    std::vector<int32_t> item;
    std::vector<double> rating;
};

void die(const char *what, int rc) {
    fprintf(stderr, "primalcr_b200: %s failed (%d): %s\n", what, rc, primalcr_last_error());
    exit(1);
}
#define CK(call) do { int _rc = (call); if (_rc != 0) die(#call, _rc); } while (0)

void log_line(const char *line, void *) { std::cout << line << std::endl; }

Flat flatten_train(smat_t &R) {
    Flat f; f.d1 = R.rows; f.d2 = R.cols; f.nnz = R.nnz;
    f.row_ptr.assign(R.row_ptr, R.row_ptr + R.rows + 1);
    f.item.resize(R.nnz); f.rating.assign(R.val_t, R.val_t + R.nnz);
    for (long e = 0; e < R.nnz; ++e) f.item[e] = (int32_t)R.col_idx[e];
    return f;
}

// convert(testset_t&, d1, d2) util.cpp:250-274, literally (entries are taken in file order)
Flat flatten_test(testset_t &T, long d1, long d2) {
    Flat f; f.d1 = d1; f.d2 = d2; f.nnz = T.nnz;
    f.row_ptr.assign(d1 + 1, 0); f.item.resize(T.nnz); f.rating.resize(T.nnz);
    long cc = 0;
    for (long j = 0; j < d1; ++j) {
        f.row_ptr[j] = cc;
        for (; cc < T.nnz; ++cc) {
            if (T.T[cc].i > j) break;
            f.item[cc] = T.T[cc].j; f.rating[cc] = T.T[cc].v;
        }
    }
    f.row_ptr[d1] = cc;
    f.nnz = cc;
    return f;
}

std::vector<double> flatten(const mat_t &M, int k) {
    std::vector<double> out(M.size() * (size_t)k);
    for (size_t i = 0; i < M.size(); ++i) for (int j = 0; j < k; ++j) out[i * k + j] = M[i][j];
    return out;
}

std::vector<int64_t> global_levels(const Flat &X) {
    std::vector<int64_t> lv;
    for (double v : X.rating) {
        const int64_t l = llround(v);
        auto it = std::lower_bound(lv.begin(), lv.end(), l);
        if (it == lv.end() || *it != l) lv.insert(it, l);
    }
    if (lv.empty()) lv.push_back(0);
    return lv;
}

void solve(int solver, smat_t &R, mat_t &U, mat_t &V, testset_t &T, parameter &param) {
    const int k = param.k;
    Flat X = flatten_train(R);
    Flat XT = flatten_test(T, X.d1, X.d2);
    R.clear_space();                                   // the reference's convert() frees X too (util.cpp:244)
    std::vector<double> Uf = flatten(U, k), Vf = flatten(V, k);
    std::vector<int64_t> levels = global_levels(X);
    int gpus = 1;
    if (const char *g = getenv("PRIMALCR_GPUS")) gpus = std::max(1, atoi(g));
    // contiguous user shards balanced by nnz
    std::vector<long> bounds(gpus + 1, 0);
    for (int r = 1; r < gpus; ++r) {
        const double target = (double)X.nnz * r / gpus;
        bounds[r] = std::lower_bound(X.row_ptr.begin(), X.row_ptr.end(), (int64_t)target) - X.row_ptr.begin();
        if (bounds[r] > X.d1) bounds[r] = X.d1;
        if (bounds[r] < bounds[r - 1]) bounds[r] = bounds[r - 1];
    }
    bounds[gpus] = X.d1;
    char uid[128] = {0};
    if (gpus > 1) CK(primalcr_nccl_unique_id(uid));
    auto worker = [&](int rank) {
        primalcr_config cfg; primalcr_default_config(&cfg);
        cfg.solver = solver; cfg.k = k; cfg.lambda = param.lambda; cfg.stepsize = param.stepsize;
        cfg.maxiter = param.maxiter; cfg.ndcg_k = param.ndcg_k; cfg.do_predict = param.do_predict; cfg.device = rank;
        primalcr_engine *e = nullptr;
        CK(primalcr_create(&e, &cfg));
        CK(primalcr_set_levels(e, levels.data(), (int)levels.size()));
        CK(primalcr_comm_init(e, rank, gpus, gpus > 1 ? uid : nullptr));
        const long u0 = bounds[rank], u1 = bounds[rank + 1];
        auto shard = [&](const Flat &F, std::vector<int64_t> &rp) {
            rp.assign(F.row_ptr.begin() + u0, F.row_ptr.begin() + u1 + 1);
            const int64_t base = rp[0];
            for (auto &x : rp) x -= base;
            return base;
        };
        std::vector<int64_t> rp, rpt;
        const int64_t b0 = shard(X, rp), bt = shard(XT, rpt);
        CK(primalcr_set_train_csr(e, u1 - u0, X.d2, rp.back(), rp.data(), X.item.data() + b0, X.rating.data() + b0));
        if (XT.nnz != 0) CK(primalcr_set_test_csr(e, rpt.back(), rpt.data(), XT.item.data() + bt, XT.rating.data() + bt));
        CK(primalcr_set_factors(e, Uf.data() + (size_t)u0 * k, Vf.data()));
        CK(primalcr_run(e, log_line, nullptr));
        CK(primalcr_get_factors(e, Uf.data() + (size_t)u0 * k, rank == 0 ? Vf.data() : nullptr));
        primalcr_destroy(e);
    };
    if (gpus == 1) worker(0);
    else {
        std::vector<std::thread> th;
        for (int r = 0; r < gpus; ++r) th.emplace_back(worker, r);
        for (auto &t : th) t.join();
    }
    for (size_t i = 0; i < U.size(); ++i) for (int j = 0; j < k; ++j) U[i][j] = Uf[i * k + j];
    for (size_t i = 0; i < V.size(); ++i) for (int j = 0; j < k; ++j) V[i][j] = Vf[i * k + j];
}

}  // namespace

extern "C" {
void pcrpp(smat_t &R, mat_t &U, mat_t &V, testset_t &T, parameter &param) { solve(PRIMALCR_SOLVER_PCRPP, R, U, V, T, param); }
void pcr(smat_t &R, mat_t &U, mat_t &V, testset_t &T, parameter &param) { solve(PRIMALCR_SOLVER_PCR, R, U, V, T, param); }
}
