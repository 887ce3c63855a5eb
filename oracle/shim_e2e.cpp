// shim_e2e.cpp -- MEASUREMENT TOOL (test infrastructure, built into oracle/_ref like gpu-omp-pmf-train).
//
// Times the real drop-in call at full size: the reference's containers (smat_t, testset_t, mat_t =
// vector<vector<double>> in pageable memory, parameter) are filled exactly as run_pcrpp() fills them
// (pmf-train.cpp:247-275: load -> initial(U), initial(V) -> pcrpp()), then `pcrpp(R, U, V, T, param)` -- our
// shim (primalcr_b200/shim/pcr_shim.cpp) + libprimalcr_b200.so -- is called once and wall-clocked.  Only the
// text parsing of load() is skipped: the CSR comes from a binary dump written by bench.py, because the reference's
// fgets/sscanf loader needs ~10 minutes for 100 M ratings and is not part of the solver call being measured.
//
//   shim-e2e <csr.bin> <k> <lambda> <maxiter> [solver=2]
//   csr.bin = int64 d1, d2, nnz; int64 row_ptr[d1+1]; int32 item[nnz]; float64 rating[nnz]
// stdout: the solver's own log lines, then one line  "SHIM_E2E seconds=<wall of the pcrpp() call> iters=<maxiter>".
#include "util.h"
#include "pmf.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>

int main(int argc, char **argv) {
    if (argc < 5) { fprintf(stderr, "usage: shim-e2e csr.bin k lambda maxiter [solver]\n"); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    long hdr[3];
    if (fread(hdr, sizeof(long), 3, f) != 3) return 1;
    const long d1 = hdr[0], d2 = hdr[1], nnz = hdr[2];
    smat_t R;
    R.rows = d1; R.cols = d2; R.nnz = nnz; R.mem_alloc_by_me = true; R.with_weights = false;
    R.row_ptr = (long *)malloc(sizeof(long) * (d1 + 1));
    R.col_idx = (unsigned *)malloc(sizeof(unsigned) * (nnz > 0 ? nnz : 1));
    R.val_t = (double *)malloc(sizeof(double) * (nnz > 0 ? nnz : 1));
    // the column-major half of smat_t is never read by pcr()/pcrpp() (convert() util.cpp:219-247 walks the rows)
    R.col_ptr = (long *)calloc(d2 + 1, sizeof(long)); R.row_idx = (unsigned *)malloc(sizeof(unsigned)); R.val = (double *)malloc(sizeof(double));
    if (fread(R.row_ptr, sizeof(long), d1 + 1, f) != (size_t)(d1 + 1)) return 1;
    if (fread(R.col_idx, sizeof(unsigned), nnz, f) != (size_t)nnz) return 1;
    if (fread(R.val_t, sizeof(double), nnz, f) != (size_t)nnz) return 1;
    fclose(f);
    testset_t T;                      // no test set: T.nnz == 0 (bench.py runs with -p 0)
    T.rows = d1; T.cols = d2; T.nnz = 0;
    parameter param;
    param.k = atoi(argv[2]); param.lambda = atof(argv[3]); param.maxiter = atoi(argv[4]);
    param.solver_type = argc > 5 ? atoi(argv[5]) : PCRPP;
    param.do_predict = 0; param.threads = 1;
    mat_t U, V;
    initial(U, d1, param.k); initial(V, d2, param.k);     // util.cpp:80-93, as run_pcrpp() does
    const auto t0 = std::chrono::steady_clock::now();
    if (param.solver_type == PCR) pcr(R, U, V, T, param); else pcrpp(R, U, V, T, param);
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    double chk = 0.0;
    for (size_t i = 0; i < V.size(); ++i) chk += V[i][0];
    printf("SHIM_E2E seconds=%.6f iters=%d checksum=%.17g\n", sec, param.maxiter, chk);
    return 0;
}
