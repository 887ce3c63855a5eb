// block_prims.cuh -- deterministic block-wide scan / sum used by the per-user and heavy-user kernels.
#pragma once
#include "common.cuh"

namespace pcr {

#ifndef FULL
#define FULL 0xffffffffu
#endif


// a[0..n) -> exclusive prefix sums in place, a[n] = total.  Fixed summation tree => deterministic.
// All threads call; caller synchronises before; ends with __syncthreads().
template <typename T, int THREADS>
__device__ __forceinline__ void block_excl_scan(T *a, int n, T *wsum /* [THREADS/32 + 1] shared */) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = (n + THREADS - 1) / THREADS;
    int lo = tid * chunk; if (lo > n) lo = n;
    int hi = lo + chunk;  if (hi > n) hi = n;
    T local = 0;
    for (int q = lo; q < hi; ++q) local += a[q];
    T incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { T t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        T w = lane < THREADS / 32 ? wsum[lane] : (T)0;
        T wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { T t = __shfl_up_sync(FULL, wi, o); if (lane >= o) wi += t; }
        if (lane < THREADS / 32) wsum[lane] = wi - w;
        if (lane == 31) wsum[THREADS / 32] = wi;
    }
    __syncthreads();
    T run = wsum[warp] + (incl - local);
    for (int q = lo; q < hi; ++q) { T v = a[q]; a[q] = run; run += v; }
    if (tid == 0) a[n] = wsum[THREADS / 32];
    __syncthreads();
}

template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double *wsum /* [THREADS/32] shared */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    __syncthreads();
    if (lane == 0) wsum[warp] = v;
    __syncthreads();
    double r = 0;
    if (warp == 0) {
        r = lane < THREADS / 32 ? wsum[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(FULL, r, o);
    }
    return r;   // valid in warp 0
}


}  // namespace pcr
