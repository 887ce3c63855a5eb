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
    for (size_t i = 0; i < M.size(); ++i) for (int j = 0; j < k; ++j) out[i * k + j] = M[i][j];
    return out;
}

void solve(int solver, smat_t &R, mat_t &U, mat_t &V, testset_t &T, parameter &param) {
    const int k = param.k;
    // training set: smat_t's row-major arrays (util.h:157-271): row_ptr / col_idx / val_t, items ascending per row
    const long d1 = R.rows, d2 = R.cols, nnz = R.nnz;
    std::vector<int64_t> row_ptr(R.row_ptr, R.row_ptr + d1 + 1);
    std::vector<int32_t> item((size_t)nnz);
    std::vector<double> rating(R.val_t, R.val_t + nnz);
    for (long e = 0; e < nnz; ++e) item[e] = (int32_t)R.col_idx[e];
    R.clear_space();                                   // the reference's convert() frees X too (util.cpp:244)
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
    primalcr_config cfg; primalcr_default_config(&cfg);
    cfg.solver = solver; cfg.k = k; cfg.lambda = param.lambda; cfg.stepsize = param.stepsize;
    cfg.maxiter = param.maxiter; cfg.ndcg_k = param.ndcg_k; cfg.do_predict = param.do_predict;
    cfg.threads = param.threads;
    pcrhost::FlatCsr fx{d1, d2, nnz, row_ptr.data(), item.data(), rating.data()};
    pcrhost::FlatCsr ft{d1, d2, cc, rpt.data(), itt.data(), rat.data()};
    pcrhost::solve(cfg, fx, ft, Uf.data(), Vf.data(), pcrhost::gpus_from_env());
    for (size_t i = 0; i < U.size(); ++i) for (int j = 0; j < k; ++j) U[i][j] = Uf[i * k + j];
    for (size_t i = 0; i < V.size(); ++i) for (int j = 0; j < k; ++j) V[i][j] = Vf[i * k + j];
}

}  // namespace

extern "C" {
void pcrpp(smat_t &R, mat_t &U, mat_t &V, testset_t &T, parameter &param) { solve(PRIMALCR_SOLVER_PCRPP, R, U, V, T, param); }
void pcr(smat_t &R, mat_t &U, mat_t &V, testset_t &T, parameter &param) { solve(PRIMALCR_SOLVER_PCR, R, U, V, T, param); }
}
