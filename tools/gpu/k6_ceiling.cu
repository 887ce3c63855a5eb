// k6_ceiling.cu -- MICRO-BENCHMARK (not product code): what would a shared-memory-resident per-user CG round cost?
//
// SURVEY 2.1 K6 asks for a per-user truncated Newton-CG that stages the user's V rows once and runs all CG rounds from
// shared memory.  Before building it, measure its ceiling: one CTA owns one user of L ratings (k = 100: an L x 800-byte
// tile), stages the rows from a 14 MB item table through L2 (random item ids), then runs R rounds of
//     b_j = V_j . p  (8-lane groups, shuffle reduce, as dots_units_kernel)   -> "sweep" stand-in c_j = f(b_j)
//     Hp  = sum_j c_j V_j  (lanes own 16-byte column chunks, warps split the rows, cross-warp sum in shared memory)
//     p  <- Hp-dependent update (serial dependency between rounds, as in CG)
// and reports SM-clocks per rating per round (both passes) against the same two passes of the product kernels through
// L2 (dots 4.2-4.6 ms + user-major row sum 4.65 ms per 1e8 ratings on one B200 = ~26 clk per rating and SM).
// The shared-memory floor is 2 x 800 B / (128 B/clk) = 12.5 clk.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o k6_ceiling k6_ceiling.cu && ./k6_ceiling
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define FULL 0xffffffffu
static const int K = 100, LD = 112, NCH = 50;       // 50 double2 chunks per row

template <int L>
__global__ void __launch_bounds__(256) k6_kernel(const double *__restrict__ V, const int *__restrict__ items, int n_users, int rounds,
                                                 double *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smraw[];
    double2 *tile = reinterpret_cast<double2 *>(smraw);                 // [L][NCH]
    double *b = reinterpret_cast<double *>(tile + (size_t)L * NCH);     // [L]
    double2 *hp_part = reinterpret_cast<double2 *>(b + L);              // [8][64]
    double2 *pvec = hp_part + 8 * 64;                                   // [64]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lg = lane & 7, grp = lane >> 3;
    for (int u = blockIdx.x; u < n_users; u += gridDim.x) {
        // ---- stage: gather L rows through L2 (one 8-lane group per row, 7 x 16-byte loads per lane)
        for (int j = warp * 4 + grp; j < L; j += 32) {
            const double2 *row = reinterpret_cast<const double2 *>(V + (size_t)items[(size_t)u * L + j] * LD);
#pragma unroll
            for (int i = 0; i < 7; ++i) { const int c = lg + 8 * i; if (c < NCH) tile[(size_t)j * NCH + c] = __ldg(row + c); }
        }
        if (tid < 64) pvec[tid] = make_double2(1e-3 * (tid + 1), -1e-3 * tid);
        __syncthreads();
        for (int r = 0; r < rounds; ++r) {
            // ---- b_j = V_j . p
            double2 pr[7];
#pragma unroll
            for (int i = 0; i < 7; ++i) { const int c = lg + 8 * i; pr[i] = c < NCH ? pvec[c] : make_double2(0.0, 0.0); }
            for (int j = warp * 4 + grp; j < L; j += 32) {
                double ax = 0.0, ay = 0.0;
#pragma unroll
                for (int i = 0; i < 7; ++i) {
                    const int c = lg + 8 * i;
                    if (c < NCH) { const double2 x = tile[(size_t)j * NCH + c]; ax = fma(pr[i].x, x.x, ax); ay = fma(pr[i].y, x.y, ay); }
                }
                double s = ax + ay;
                s += __shfl_xor_sync(FULL, s, 4); s += __shfl_xor_sync(FULL, s, 2); s += __shfl_xor_sync(FULL, s, 1);
                if (lg == 0) b[j] = s * 0.5;          // sweep stand-in
            }
            __syncthreads();
            // ---- Hp = sum_j c_j V_j : warp w takes rows w, w+8, ...; lane owns chunks lane and lane+32
            double2 a0 = make_double2(0.0, 0.0), a1 = make_double2(0.0, 0.0);
#pragma unroll 4
            for (int j = warp; j < L; j += 8) {
                const double c = b[j];
                const double2 x0 = tile[(size_t)j * NCH + lane];
                a0.x = fma(c, x0.x, a0.x); a0.y = fma(c, x0.y, a0.y);
                if (lane + 32 < NCH) { const double2 x1 = tile[(size_t)j * NCH + lane + 32]; a1.x = fma(c, x1.x, a1.x); a1.y = fma(c, x1.y, a1.y); }
            }
            hp_part[warp * 64 + lane] = a0; hp_part[warp * 64 + lane + 32] = a1;
            __syncthreads();
            if (tid < NCH) {
                double2 h = make_double2(0.0, 0.0);
#pragma unroll
                for (int w = 0; w < 8; ++w) { const double2 t = hp_part[w * 64 + tid]; h.x += t.x; h.y += t.y; }
                const double2 p0 = pvec[tid];
                pvec[tid] = make_double2(p0.x * 0.999 + 1e-6 * h.x, p0.y * 0.999 + 1e-6 * h.y);
            }
            __syncthreads();
        }
        if (tid < NCH) { out[(size_t)u * 2 * NCH + 2 * tid] = pvec[tid].x; out[(size_t)u * 2 * NCH + 2 * tid + 1] = pvec[tid].y; }
        __syncthreads();
    }
}

template <int L>
static void run(const double *V, int d2, int rounds, int sms) {
    const int n_users = 200000 * 64 / L;                                   // 12.8 M ratings
    std::vector<int> h((size_t)n_users * L);
    unsigned long long s = 12345;
    for (auto &x : h) { s = s * 6364136223846793005ull + 1442695040888963407ull; x = (int)((s >> 33) % (unsigned)d2); }
    int *items; double *out;
    cudaMalloc(&items, h.size() * 4); cudaMalloc(&out, (size_t)n_users * 2 * NCH * 8);
    cudaMemcpy(items, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = (size_t)L * NCH * 16 + (size_t)L * 8 + 8 * 64 * 16 + 64 * 16;
    cudaFuncSetAttribute(k6_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k6_kernel<L>, 256, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms0 = 0, ms1 = 0;
    for (int pass = 0; pass < 2; ++pass) {
        const int R = pass == 0 ? 0 : rounds;                              // pass 0: staging only
        k6_kernel<L><<<per_sm * sms, 256, smem>>>(V, items, n_users, R, out);       // warm-up
        cudaEventRecord(e0);
        k6_kernel<L><<<per_sm * sms, 256, smem>>>(V, items, n_users, R, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(pass == 0 ? &ms0 : &ms1, e0, e1);
    }
    const double ratings = (double)n_users * L;
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double clk_per = (ms1 - ms0) * 1e-3 * clk_khz * 1e3 * sms / (ratings * rounds);
    printf("{\"L\": %d, \"ctas_per_sm\": %d, \"smem_bytes\": %zu, \"stage_ms\": %.3f, \"stage_GBps\": %.0f, \"rounds\": %d, \"total_ms\": %.3f, "
           "\"ns_per_rating_round\": %.5f, \"sm_clk_per_rating_round_at_max_clock\": %.2f, \"err\": \"%s\"}\n",
           L, per_sm, smem, ms0, ratings * 800 / (ms0 * 1e-3) / 1e9, rounds, ms1, (ms1 - ms0) * 1e6 / (ratings * rounds), clk_per,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(items); cudaFree(out);
}

int main() {
    const int d2 = 17770;
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    std::vector<double> hv((size_t)d2 * LD, 0.0);
    for (size_t i = 0; i < hv.size(); ++i) hv[i] = (double)((i * 2654435761u) % 1000) * 1e-3 - 0.5;
    double *V; cudaMalloc(&V, hv.size() * 8); cudaMemcpy(V, hv.data(), hv.size() * 8, cudaMemcpyHostToDevice);
    run<32>(V, d2, 8, prop.multiProcessorCount);
    run<64>(V, d2, 8, prop.multiProcessorCount);
    run<128>(V, d2, 8, prop.multiProcessorCount);
    run<256>(V, d2, 8, prop.multiProcessorCount);
    return 0;
}
