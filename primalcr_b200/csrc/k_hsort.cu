// k_hsort.cu -- segmented sort of the heavy users' scores (users with more ratings than the largest tile), hand-written:
// the heavy end of get_sorted_mm pcrpp.cpp:52-83 (the reference calls std::sort on every user; tiles of short users are
// sorted by the bitonic networks of k_tiles.cu).
//
// Every heavy user is already cut into chunks of HEAVY_CHUNK = 2048 ratings (k_heavy.cu).  The sort is a grid over those
// chunks at every stage, so its critical path does not depend on the longest user:
//   1. chunk sort : one CTA sorts one chunk in shared memory (bitonic network on (score, position) pairs);
//   2. merge pass p = 0, 1, ... : runs of 2048 << p ratings are merged pairwise; every CTA produces ONE output tile of 2048
//      ratings, finds its two input ranges with a merge-path binary search, merges them in shared memory and writes the
//      tile with coalesced stores.  A user of len ratings needs ceil(log2(ceil(len / 2048))) passes; users that are done
//      simply have no work in later passes.
// Order: ascending score, ties (equal as doubles, so -0.0 == +0.0) by original position -- the same total order as the
// tile kernels (the reference's std::sort is unstable, no output of it depends on the order of ties), hence identical
// sorted arrays on every path.  The data ping-pongs between the final arrays (SortedMeta::s / pos) and a scratch pair; the starting side is chosen per user so that the last pass lands in the
// final arrays.  12 bytes per rating and pass; no library call on the path (round 1 used cub::DeviceSegmentedSort here).
#include "kernels.h"
#include <math_constants.h>

namespace pcr {

#define LAUNCH(ctx, name, bytes, kernel, grid, block, smem, ...)                         \
    do {                                                                                 \
        (ctx).prof->begin(name, (ctx).stream, (double)(bytes));                          \
        kernel<<<(grid), (block), (smem), (ctx).stream>>>(__VA_ARGS__);                  \
        (ctx).prof->end((ctx).stream);                                                   \
        PCR_CUDA(cudaGetLastError());                                                    \
    } while (0)

static const int HS_N = HEAVY_CHUNK;        // ratings per chunk / output tile
static const int HS_T = 256;                // threads per CTA
static const int HS_E = HS_N / HS_T;        // outputs per thread in the merge

__device__ __forceinline__ bool hs_less(double ka, int pa, double kb, int pb) { return ka < kb || (ka == kb && pa < pb); }

__device__ __forceinline__ int hs_passes(int len) {       // merge passes a user of len ratings needs
    const int chunks = (len + HS_N - 1) / HS_N;
    int p = 0;
    while ((1 << p) < chunks) ++p;
    return p;
}

// ---------------------------------------------------------------- 1. chunk sort
__global__ void __launch_bounds__(HS_T) hs_chunk_sort_kernel(HeavySortPlan h, const double *__restrict__ m,
                                                             double *__restrict__ s_final, int32_t *__restrict__ pos_final) {
    __shared__ double keys[HS_N];
    __shared__ int32_t idx[HS_N];
    const int q = h.chunk_user[blockIdx.x], lo = h.chunk_lo[blockIdx.x];
    const i64 begin = h.begin[q];
    const int len = (int)(h.end[q] - begin);
    const int cnt = len - lo < HS_N ? len - lo : HS_N;
    const int tid = threadIdx.x;
    for (int j = tid; j < HS_N; j += HS_T) {
        keys[j] = j < cnt ? m[begin + lo + j] : CUDART_INF;       // padding sorts last (ties with a real +inf: larger position)
        idx[j] = j < cnt ? (int32_t)(begin + lo + j) : 0x7fffffff;
    }
    __syncthreads();
    for (int k = 2; k <= HS_N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (HS_N >> 1); t += HS_T) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const bool asc = (i & k) == 0;
                const double ki = keys[i], kp = keys[p];
                const int32_t ii = idx[i], ip = idx[p];
                const bool gt = hs_less(kp, ip, ki, ii);
                if (gt == asc) { keys[i] = kp; keys[p] = ki; idx[i] = ip; idx[p] = ii; }
            }
            __syncthreads();
        }
    }
    // the side this user starts on: after hs_passes(len) ping-pongs the data must sit in the final arrays
    const bool to_tmp = (hs_passes(len) & 1) != 0;
    double *so = to_tmp ? h.tmp_s + h.off[q] + lo : s_final + begin + lo;
    int32_t *po = to_tmp ? h.tmp_pos + h.off[q] + lo : pos_final + begin + lo;
    for (int j = tid; j < cnt; j += HS_T) { so[j] = keys[j]; po[j] = idx[j]; }
}

// ---------------------------------------------------------------- 2. one merge pass
// number of elements taken from run A among the first d outputs of merge(A, B) (stable: A before B on ties)
__device__ __forceinline__ int hs_merge_path(const double *ka, const int32_t *pa, int na, const double *kb, const int32_t *pb, int nb, int d) {
    int lo = d - nb > 0 ? d - nb : 0, hi = d < na ? d : na;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int j = d - 1 - mid;
        if (!hs_less(kb[j], pb[j], ka[mid], pa[mid])) lo = mid + 1; else hi = mid;      // A[mid] <= B[j]: it precedes B[j]
    }
    return lo;
}

__global__ void __launch_bounds__(HS_T) hs_merge_kernel(HeavySortPlan h, int pass, double *__restrict__ s_final,
                                                        int32_t *__restrict__ pos_final) {
    __shared__ double sk[HS_N];
    __shared__ int32_t sp[HS_N];
    __shared__ int s_a[2];
    const int q = h.chunk_user[blockIdx.x], lo = h.chunk_lo[blockIdx.x];
    const i64 begin = h.begin[q];
    const int len = (int)(h.end[q] - begin);
    const int P = hs_passes(len);
    if (pass >= P) return;                                   // this user is already sorted
    const int src_side = (P & 1) ^ (pass & 1);               // 1: scratch pair, 0: final arrays
    const double *ks = src_side ? h.tmp_s + h.off[q] : s_final + begin;
    const int32_t *ps = src_side ? h.tmp_pos + h.off[q] : pos_final + begin;
    double *kd = src_side ? s_final + begin : h.tmp_s + h.off[q];
    int32_t *pd = src_side ? pos_final + begin : h.tmp_pos + h.off[q];
    const i64 rl = (i64)HS_N << pass;                        // run length of this pass
    const i64 pair0 = ((i64)lo / (2 * rl)) * (2 * rl);
    const int na = (int)((pair0 + rl < len ? pair0 + rl : len) - pair0);
    const i64 bstart = pair0 + na;
    const int nb = (int)((pair0 + 2 * rl < len ? pair0 + 2 * rl : len) - bstart);
    const int d0 = (int)(lo - pair0);
    const int d1 = d0 + HS_N < na + nb ? d0 + HS_N : na + nb;
    const int tid = threadIdx.x;
    if (tid == 0) s_a[0] = hs_merge_path(ks + pair0, ps + pair0, na, ks + bstart, ps + bstart, nb, d0);
    if (tid == 32) s_a[1] = hs_merge_path(ks + pair0, ps + pair0, na, ks + bstart, ps + bstart, nb, d1);
    __syncthreads();
    const int a0 = s_a[0], a1 = s_a[1];
    const int b0 = d0 - a0, b1 = d1 - a1;
    const int ca = a1 - a0, cb = b1 - b0, total = ca + cb;   // total == d1 - d0 <= HS_N
    for (int j = tid; j < total; j += HS_T) {
        const i64 src = j < ca ? pair0 + a0 + j : bstart + b0 + (j - ca);
        sk[j] = ks[src]; sp[j] = ps[src];
    }
    __syncthreads();
    // every thread merges HS_E consecutive outputs
    double ok[HS_E]; int32_t op[HS_E];
    const int o0 = tid * HS_E;
    if (o0 < total) {
        int i = hs_merge_path(sk, sp, ca, sk + ca, sp + ca, cb, o0);
        int j = o0 - i;
#pragma unroll
        for (int e = 0; e < HS_E; ++e) {
            const bool has_a = i < ca, has_b = j < cb;
            bool take_a = has_a;
            if (has_a && has_b) take_a = !hs_less(sk[ca + j], sp[ca + j], sk[i], sp[i]);
            if (has_a || has_b) {
                const int src = take_a ? i : ca + j;
                ok[e] = sk[src]; op[e] = sp[src];
                if (take_a) ++i; else ++j;
            }
        }
    }
    __syncthreads();
    if (o0 < total) {
#pragma unroll
        for (int e = 0; e < HS_E; ++e) if (o0 + e < total) { sk[o0 + e] = ok[e]; sp[o0 + e] = op[e]; }
    }
    __syncthreads();
    for (int j = tid; j < total; j += HS_T) { kd[lo + j] = sk[j]; pd[lo + j] = sp[j]; }
}

void k_heavy_sort(Ctx &c, const HeavySortPlan &h, const double *m, double *s_sorted, int32_t *pos_sorted) {
    if (h.n_chunks <= 0) return;
    LAUNCH(c, "heavy_chunk_sort", 0.0, hs_chunk_sort_kernel, (unsigned)h.n_chunks, HS_T, 0, h, m, s_sorted, pos_sorted);
    for (int p = 0; p < h.max_passes; ++p)
        LAUNCH(c, "heavy_merge", 0.0, hs_merge_kernel, (unsigned)h.n_chunks, HS_T, 0, h, p, s_sorted, pos_sorted);
}

}  // namespace pcr
