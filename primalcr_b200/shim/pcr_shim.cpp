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
#include "../host/driver.hpp"

namespace {

std::vector<double> flatten(const mat_t &M, int k) {
    std::vector<double> out(M.size() * (size_t)k);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)M.size(); ++i) for (int j = 0; j < k; ++j) out[(size_t)i * k + j] = M[i][j];
    return out;
}

void solve(int solver, smat_t &R, mat_t &U, mat_t &V, testset_t &T, parameter &param) {
    pcrhost::Lap lap;
    const int k = param.k;
    // training set: smat_t's row-major arrays (util.h:157-271): row_ptr (long) / col_idx (unsigned) / val_t (double), items
    // ascending per row.  They already ARE a CSR with 64-bit offsets, 32-bit ids and fp64 values: handed to the C ABI in
    // place, no copies (ids are < 2^31: the loader reads them with %d, util.h:126).
    const long d1 = R.rows, d2 = R.cols, nnz = R.nnz;
    static_assert(sizeof(long) == sizeof(int64_t) && sizeof(unsigned) == sizeof(int32_t), "LP64 expected");
    const int64_t *row_ptr = reinterpret_cast<const int64_t *>(R.row_ptr);
    const int32_t *item = reinterpret_cast<const int32_t *>(R.col_idx);
    const double *rating = R.val_t;
    // test set: convert(testset_t&, d1, d2) util.cpp:250-274, literally (entries are taken in file order)
    std::vector<int64_t> rpt((size_t)d1 + 1, 0);
    std::vector<int32_t> itt((size_t)T.nnz);
    std::vector<double> rat((size_t)T.nnz);
    long cc = 0;
    for (long j = 0; j < d1; ++j) {
        rpt[j] = cc;
        for (; cc < T.nnz; ++cc) {
            if (T.T[cc].i > j) break;
            itt[cc] = T.T[cc].j; rat[cc] = T.T[cc].v;
        }
    }
    rpt[d1] = cc;
    std::vector<double> Uf = flatten(U, k), Vf = flatten(V, k);
    lap("shim: flatten test set, U, V");
    primalcr_config cfg; primalcr_default_config(&cfg);
    cfg.solver = solver; cfg.k = k; cfg.lambda = param.lambda; cfg.stepsize = param.stepsize;
    cfg.maxiter = param.maxiter; cfg.ndcg_k = param.ndcg_k; cfg.do_predict = param.do_predict;
    cfg.threads = param.threads;
    pcrhost::FlatCsr fx{d1, d2, nnz, row_ptr, item, rating};
    pcrhost::FlatCsr ft{d1, d2, cc, rpt.data(), itt.data(), rat.data()};
    pcrhost::solve(cfg, fx, ft, Uf.data(), Vf.data(), pcrhost::gpus_from_env());
    lap("shim: solve (see host laps)");
    R.clear_space();                                   // the reference's convert() frees X too (util.cpp:244)
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)U.size(); ++i) for (int j = 0; j < k; ++j) U[i][j] = Uf[(size_t)i * k + j];
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)V.size(); ++i) for (int j = 0; j < k; ++j) V[i][j] = Vf[(size_t)i * k + j];
    lap("shim: copy U, V back into mat_t");
}

}  // namespace

extern "C" {
void pcrpp(smat_t &R, mat_t &U, mat_t &V, testset_t &T, parameter &param) { solve(PRIMALCR_SOLVER_PCRPP, R, U, V, T, param); }
void pcr(smat_t &R, mat_t &U, mat_t &V, testset_t &T, parameter &param) { solve(PRIMALCR_SOLVER_PCR, R, U, V, T, param); }
}
