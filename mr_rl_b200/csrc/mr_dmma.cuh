// FP64 tensor-pipe building block shared by the GP kernels: one CTA (256 threads, 2 x 4 warps, each
// 64 x 32) accumulates a 128 x 128 tile  C += A . B^T  with A [128 x K] and B [128 x K] row-major
// (K contiguous, "TN" form), K = k_tiles * 16, through a 3-stage cp.async shared-memory pipeline and
// mma.sync.m8n8k4.f64 (DMMA).  tcgen05 has no fp64 kind, so this is the tensor path for fp64 on sm_100a.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace mr {

constexpr int GP_BM = 128, GP_BN = 128, GP_BK = 16, GP_LD = GP_BK + 4;   // +4 doubles: conflict-free fragments
constexpr int GP_STAGE = (GP_BM + GP_BN) * GP_LD;                         // doubles per pipeline stage
constexpr int GP_STAGES = 3;
constexpr size_t kDmmaSmemBytes = (size_t)GP_STAGES * GP_STAGE * sizeof(double);

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Fragment owner: acc[i][j][e] is C[warp_m*64 + i*8 + (lane>>2)][warp_n*32 + j*8 + (lane&3)*2 + e].
__device__ __forceinline__ void dmma_tile_tn(const double* __restrict__ a_src, int64_t lda,
                                             const double* __restrict__ b_src, int64_t ldb, int k_tiles,
                                             double (&acc)[8][4][2], double* smem) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warp_m = warp >> 2, warp_n = warp & 3;
    const int g = lane >> 2, t4 = lane & 3;
    auto load_stage = [&](int stage, int kt) {
        double* As = smem + stage * GP_STAGE;
        double* Bs = As + GP_BM * GP_LD;
        const int c0 = kt * GP_BK;
        // 128 rows x 16 doubles = 1024 16-byte chunks per operand; 256 threads x 4
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int chunk = tid + it * 256;
            const int r = chunk >> 3, cc = (chunk & 7) * 2;
            cp_async16(As + r * GP_LD + cc, a_src + (int64_t)r * lda + c0 + cc);
            cp_async16(Bs + r * GP_LD + cc, b_src + (int64_t)r * ldb + c0 + cc);
        }
    };
#pragma unroll
    for (int s = 0; s < GP_STAGES - 1; ++s) {
        if (s < k_tiles) load_stage(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < k_tiles; ++kt) {
        cp_async_wait<GP_STAGES - 2>();
        __syncthreads();
        const int nxt = kt + GP_STAGES - 1;
        if (nxt < k_tiles) load_stage(nxt % GP_STAGES, nxt);
        cp_async_commit();
        const double* As = smem + (kt % GP_STAGES) * GP_STAGE + (warp_m * 64) * GP_LD;
        const double* Bs = smem + (kt % GP_STAGES) * GP_STAGE + GP_BM * GP_LD + (warp_n * 32) * GP_LD;
#pragma unroll
        for (int k4 = 0; k4 < GP_BK / 4; ++k4) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = As[(i * 8 + g) * GP_LD + k4 * 4 + t4];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[(j * 8 + g) * GP_LD + k4 * 4 + t4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();                                            // smem may be reused by the caller
}

}  // namespace mr
