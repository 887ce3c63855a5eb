// ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Thin C-ABI driver around the UNMODIFIED reference objects (pcrpp.o, pcr.o, util.o compiled by
// oracle/Makefile straight from /root/reference into oracle/_ref/).  It lets the tests call the
// reference's own stage functions (all have external linkage) on flat arrays and read back results
// at full fp64 precision (the reference CLI prints 6 digits only).  The driver loop in ref_train()
// repeats the body of pcrpp() (pcrpp.cpp:857-878) / pcr() (pcr.cpp:637-679) around the reference's
// update_V*/update_U* so that per-iteration objectives can be captured.
//
// Only built where /root/reference exists; outputs go to oracle/_ref/ (git-ignored, travels with gpurun).
#include "util.h"
#include "pmf.h"

// re-declaration of the reference's private struct (pcrpp.cpp:5-14)
struct infor_ui {
    long num_levels;
    vec_t mm_sorted;
    vector<long> vals_sorted;
    vector<long> d2bar_sorted;
    vector<long> count_right;
    long len;
};

// reference stage functions (pcrpp.cpp / pcr.cpp; external linkage, C++ mangled)
double* comp_m_new(const mat_t& U, const mat_t& V, SparseMat* X, int r);
vec_t get_sorted_mm(double* m, long start, long end, long len, vector<long>& perm_ind);
mat_t obtain_g_new(const mat_t& U, const mat_t& V, SparseMat* X, double* m, double lambda);
vec_t compute_Ha_new(const vec_t& a, double* m, const mat_t& U, SparseMat* X, int r, double lambda);
double objective_new(double* m, const mat_t& U, const mat_t& V, SparseMat* X, double lambda);
double* update_V_new(SparseMat* X, double lambda, double stepsize, int r, const mat_t& U, mat_t& V, double& now_obj);
mat_t update_U_new(SparseMat* X, double* m, double lambda, double stepsize, int r, const mat_t& V, const mat_t& U, double& now_obj);
infor_ui* precompute_ui(long i, const mat_t& V, SparseMat* X, double* m);
vec_t obtain_g_u_new(long i, const mat_t& V, double lambda, const vec_t& ui, infor_ui* p);
double objective_u_new(long i, infor_ui* p, const vec_t& ui, double lambda);
vec_t obtain_Hs_new(long i, const vec_t& s, const mat_t& V, infor_ui* p, double lambda);

double objective(double* m, const mat_t& U, const mat_t& V, SparseMat* X, double lambda);
double* comp_m(const mat_t& U, const mat_t& V, SparseMat* X, int r);
mat_t obtain_g(const mat_t& U, const mat_t& V, SparseMat* X, double* m, double lambda);
vec_t compute_Ha(const vec_t& a, double* m, const mat_t& U, SparseMat* X, int r, double lambda);
double* update_V(SparseMat* X, double lambda, double stepsize, int r, const mat_t& U, mat_t& V, double& now_obj);
mat_t update_U(SparseMat* X, double* m, double lambda, double stepsize, int r, const mat_t& V, const mat_t& U, double& now_obj);

namespace {

SparseMat* make_sparse(long d1, long d2, long nnz, const long* index, const long* rows, const double* vals) {
    SparseMat* X = new SparseMat(d1, d2, nnz > 0 ? nnz : 1);
    X->nnz = nnz;
    for (long i = 0; i <= d1; ++i) X->index[i] = index[i];
    for (long i = 0; i < d1; ++i)
        for (long e = index[i]; e < index[i + 1]; ++e) { X->rows[e] = rows[e]; X->cols[e] = i; X->vals[e] = vals[e]; }
    return X;
}

mat_t to_mat(const double* A, long n, int r) {
    mat_t M(n, vec_t(r));
    for (long i = 0; i < n; ++i) for (int j = 0; j < r; ++j) M[i][j] = A[i * r + j];
    return M;
}

void from_mat(const mat_t& M, double* A, int r) {
    for (size_t i = 0; i < M.size(); ++i) for (int j = 0; j < r; ++j) A[i * r + j] = M[i][j];
}

}  // namespace

extern "C" {

// util.cpp:80-93 -- the default-seeded N(0,1) stream the CLI uses for U and V
void ref_initial(double* out, long n, long k) {
    mat_t X;
    initial(X, n, k);
    from_mat(X, out, (int)k);
}

void ref_comp_m(long d1, long d2, long nnz, const long* index, const long* rows, const double* vals,
                const double* U, const double* V, int r, double* m_out) {
    omp_set_num_threads(1);
    SparseMat* X = make_sparse(d1, d2, nnz, index, rows, vals);
    mat_t Um = to_mat(U, d1, r), Vm = to_mat(V, d2, r);
    double* m = comp_m_new(Um, Vm, X, r);
    for (long i = 0; i < nnz; ++i) m_out[i] = m[i];
    delete[] m; delete X;
}

void ref_sorted_mm(const double* mm, long len, double* mm_sorted, long* perm) {
    vector<long> perm_ind(len, 0);
    vector<double> tmp(mm, mm + len);
    vec_t s = get_sorted_mm(tmp.data(), 0, len - 1, len, perm_ind);
    for (long i = 0; i < len; ++i) { mm_sorted[i] = s[i]; perm[i] = perm_ind[i]; }
}

void ref_obtain_g_new(long d1, long d2, long nnz, const long* index, const long* rows, const double* vals,
                      const double* U, const double* V, int r, const double* m, double lambda, double* g) {
    omp_set_num_threads(1);
    SparseMat* X = make_sparse(d1, d2, nnz, index, rows, vals);
    mat_t Um = to_mat(U, d1, r), Vm = to_mat(V, d2, r);
    mat_t G = obtain_g_new(Um, Vm, X, const_cast<double*>(m), lambda);
    from_mat(G, g, r);
    delete X;
}

void ref_compute_Ha_new(long d1, long d2, long nnz, const long* index, const long* rows, const double* vals,
                        const double* a, const double* m, const double* U, int r, double lambda, double* Ha) {
    omp_set_num_threads(1);
    SparseMat* X = make_sparse(d1, d2, nnz, index, rows, vals);
    mat_t Um = to_mat(U, d1, r);
    vec_t av(a, a + d2 * r);
    vec_t H = compute_Ha_new(av, const_cast<double*>(m), Um, X, r, lambda);
    for (long i = 0; i < d2 * r; ++i) Ha[i] = H[i];
    delete X;
}

double ref_objective_new(long d1, long d2, long nnz, const long* index, const long* rows, const double* vals,
                         const double* m, const double* U, const double* V, int r, double lambda) {
    omp_set_num_threads(1);
    SparseMat* X = make_sparse(d1, d2, nnz, index, rows, vals);
    mat_t Um = to_mat(U, d1, r), Vm = to_mat(V, d2, r);
    double o = objective_new(const_cast<double*>(m), Um, Vm, X, lambda);
    delete X;
    return o;
}

void ref_user_stage(long len, const long* rows, const double* vals, const double* m, const double* V, long d2, int r,
                    double lambda, const double* ui, const double* s, double* g, double* obj, double* Hs) {
    omp_set_num_threads(1);
    long index[2] = {0, len};
    SparseMat* X = make_sparse(1, d2, len, index, rows, vals);
    mat_t Vm = to_mat(V, d2, r);
    infor_ui* p = precompute_ui(0, Vm, X, const_cast<double*>(m));
    vec_t u(ui, ui + r), sv(s, s + r);
    vec_t gv = obtain_g_u_new(0, Vm, lambda, u, p);
    *obj = objective_u_new(0, p, u, lambda);
    vec_t hv = obtain_Hs_new(0, sv, Vm, p, lambda);
    for (int k = 0; k < r; ++k) { g[k] = gv[k]; Hs[k] = hv[k]; }
    delete p; delete X;
}

void ref_eval(long d1, long d2, long nnz, const long* index, const long* rows, const double* vals,
              const double* U, const double* V, int r, int ndcg_k, double* out) {
    omp_set_num_threads(1);
    SparseMat* X = make_sparse(d1, d2, nnz, index, rows, vals);
    mat_t Um = to_mat(U, d1, r), Vm = to_mat(V, d2, r);
    pair<double, double> res = compute_pairwise_error_ndcg(Um, Vm, X, ndcg_k);
    out[0] = res.first; out[1] = res.second;
    delete X;
}

double ref_pcr_objective(long d1, long d2, long nnz, const long* index, const long* rows, const double* vals,
                         const double* m, const double* U, const double* V, int r, double lambda) {
    omp_set_num_threads(1);
    SparseMat* X = make_sparse(d1, d2, nnz, index, rows, vals);
    mat_t Um = to_mat(U, d1, r), Vm = to_mat(V, d2, r);
    double o = objective(const_cast<double*>(m), Um, Vm, X, lambda);
    delete X;
    return o;
}

void ref_pcr_obtain_g(long d1, long d2, long nnz, const long* index, const long* rows, const double* vals,
                      const double* U, const double* V, int r, const double* m, double lambda, double* g) {
    omp_set_num_threads(1);
    SparseMat* X = make_sparse(d1, d2, nnz, index, rows, vals);
    mat_t Um = to_mat(U, d1, r), Vm = to_mat(V, d2, r);
    mat_t G = obtain_g(Um, Vm, X, const_cast<double*>(m), lambda);
    from_mat(G, g, r);
    delete X;
}

void ref_pcr_compute_Ha(long d1, long d2, long nnz, const long* index, const long* rows, const double* vals,
                        const double* a, const double* m, const double* U, int r, double lambda, double* Ha) {
    omp_set_num_threads(1);
    SparseMat* X = make_sparse(d1, d2, nnz, index, rows, vals);
    mat_t Um = to_mat(U, d1, r);
    vec_t av(a, a + d2 * r);
    vec_t H = compute_Ha(av, const_cast<double*>(m), Um, X, r, lambda);
    for (long i = 0; i < d2 * r; ++i) Ha[i] = H[i];
    delete X;
}

// One call = `maxiter` outer iterations of the reference solver, single-threaded (race-free).
// obj[0..maxiter]; evals[(maxiter+1)*4] = train err, train ndcg, test err, test ndcg (0 when do_predict==0).
void ref_train(int solver, long d1, long d2, long nnz, const long* index, const long* rows, const double* vals,
               long nnz_t, const long* index_t, const long* rows_t, const double* vals_t,
               double* U, double* V, int r, double lambda, double stepsize, int maxiter, int do_predict,
               int ndcg_k, int threads, double* obj, double* evals) {
    omp_set_num_threads(threads > 0 ? threads : 1);
    SparseMat* X = make_sparse(d1, d2, nnz, index, rows, vals);
    SparseMat* XT = make_sparse(d1, d2, nnz_t, index_t, rows_t, vals_t);
    mat_t Um = to_mat(U, d1, r), Vm = to_mat(V, d2, r);
    double now_obj = 0.0;
    double* m = solver == 2 ? comp_m_new(Um, Vm, X, r) : comp_m(Um, Vm, X, r);
    now_obj = solver == 2 ? objective_new(m, Um, Vm, X, lambda) : objective(m, Um, Vm, X, lambda);
    obj[0] = now_obj;
    for (int iter = 0; iter <= maxiter; ++iter) {
        if (iter > 0) {
            delete[] m;
            if (solver == 2) {
                m = update_V_new(X, lambda, stepsize, r, Um, Vm, now_obj);
                Um = update_U_new(X, m, lambda, stepsize, r, Vm, Um, now_obj);
            } else {
                m = update_V(X, lambda, stepsize, r, Um, Vm, now_obj);
                Um = update_U(X, m, lambda, stepsize, r, Vm, Um, now_obj);
            }
            obj[iter] = now_obj;
        }
        if (evals) {
            double* e = evals + iter * 4;
            e[0] = e[1] = e[2] = e[3] = 0.0;
            if (do_predict) {
                pair<double, double> a = compute_pairwise_error_ndcg(Um, Vm, X, ndcg_k);
                e[0] = a.first; e[1] = a.second;
                if (nnz_t != 0) {
                    pair<double, double> b = compute_pairwise_error_ndcg(Um, Vm, XT, ndcg_k);
                    e[2] = b.first; e[3] = b.second;
                }
            }
        }
    }
    delete[] m;
    from_mat(Um, U, r); from_mat(Vm, V, r);
    delete X; delete XT;
}

}  // extern "C"
