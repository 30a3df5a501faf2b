// Single-step kernel for the generated-noise mode (matched model), warp-specialised:
//
//   warps 0 .. kTile/32-1   integrators: one env per thread, fp64 RK45 step (FP64 pipe, long dependent chains)
//   service warp(s)         (a) the CTA's TMA issue: bulk loads of tile k+3, bulk stores of tile k
//                           (b) the 8 standard normals per env of tile k+1 (Philox4x32 + Box-Muller: integer / fp32 /
//                               MUFU pipes, four independent envs per lane) into a double-buffered shared stage
//
// Why: with every thread doing both, the kernel is latency-bound (r01 ncu: 1.0 eligible warps per cycle, issue slots
// 58 % busy, 16 warps/SM at the 128-register cap) and the RNG's ~190 instructions per env-step cost their full issue
// time.  The draws depend only on (seed, global env index, env-step index) — not on loaded data — so another warp can
// produce them ahead of time; its instruction stream is independent of the fp64 chains and fills the idle issue slots,
// and the integrators shrink to the noise-free kernel's register footprint.
// MEASURED AND REJECTED as the default (B200, 2^20 envs, fp64, profiles/r02_ncu_step_ws.txt): 45.0 us per launch against
// 32.0 us for env_step_tma_kernel (fp32 storage 35.2 against 29.3).  The draws of a tile are ~600 warp-instructions and
// the service warp also issues the tile's 20 bulk copies, while each integrator warp has ~420: the ONE service warp per
// CTA is the critical path (the integrators' top stall is the CTA barrier that waits for it), and at 96 registers the
// integrators spill (LDL/STL 0.85 M warp-instructions per launch).  Two service warps per CTA would need 80 registers
// per thread at 4 CTAs/SM, or cost a quarter of the integrator warps at 3 CTAs/SM; the register file, not the issue
// slots, is what the step is short of.  Kept, selectable with mr_set_step_path(4) / MR_STEP_PATH=ws, so the measurement
// can be repeated; results are bit-identical to the plain kernel (tests/test_gpu_env.py).
// Noise semantics are identical to PhiloxNoise::draw8 (same counters, same blocks, same Box-Muller), so this kernel and
// env_step_tma_kernel<T, MR_NOISE_PHILOX, false> produce bit-identical results (tested).
#pragma once

#include "mr_step_tma.cuh"

namespace mr {

template <class T> struct WsCfg {
    static constexpr int kTile = TileOf<T>::value;
    static constexpr int kSvcWarps = kTile / 128;                 // one service warp per 128 envs (4 envs per lane)
    static constexpr int kThreads = kTile + 32 * kSvcWarps;
    // resident CTAs per SM: fp64 4 x 160 threads (96 registers, 16 integrator warps as in the plain TMA kernel);
    // fp32 2 x 320 threads (96 registers; 3 CTAs would leave 64 registers per thread and spill the fp64 controller)
    static constexpr int kCtas = sizeof(T) == 8 ? TmaWarps<T>::value * 32 / kTile : 2;
};

template <class T, int kTile = TileOf<T>::value>
struct StepSmemWs {
    TileIn<T, kTile> in[kStagesIn];
    TileOut<T, kTile> out[kStagesOut];
    alignas(128) T zero[kTile];
    alignas(16) float4 z[2][2][kTile];                            // [buffer][normals 0-3 | 4-7][env of the tile]
    alignas(8) uint64_t full[kStagesIn];
};

template <class T>
__global__ void __launch_bounds__(WsCfg<T>::kThreads, WsCfg<T>::kCtas)
env_step_tma_ws_kernel(StateView<T> st, const T* __restrict__ actions, OutView<T> out, NoiseView nv, TimeView tv,
                       Params p, int64_t n_tiles, int64_t n_total) {
    constexpr int kTile = WsCfg<T>::kTile;
    constexpr int kSvc = 32 * WsCfg<T>::kSvcWarps;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    StepSmemWs<T>& sm = *reinterpret_cast<StepSmemWs<T>*>(smem_raw);
    const int tid = threadIdx.x;
    const bool svc = tid >= kTile;                                // warp-uniform
    const int lane_s = tid - kTile;                               // index inside the service group
    const bool elect = lane_s == 0;
    const int64_t first = blockIdx.x, stride = gridDim.x;

    if (elect) {
        for (int s = 0; s < kStagesIn; ++s) mbar_init(&sm.full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (!svc) sm.zero[tid] = (T)0;
    fence_async_smem();
    __syncthreads();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    constexpr uint32_t kRow = kTile * sizeof(T);
    constexpr uint32_t kInBytes = 5 * kRow + 2 * kRow + kTile * 4;
    const uint64_t off = step_offset(nv);

    auto issue_loads = [&](int s, int64_t tile) {
        const int64_t i0 = tile * kTile;
        uint64_t* bar = &sm.full[s];
        TileIn<T, kTile>& b = sm.in[s];
        mbar_expect_tx(bar, kInBytes);
        bulk_load(b.x, st.x + i0, kRow, bar);
        bulk_load(b.y, st.y + i0, kRow, bar);
        bulk_load(b.fx, st.fx + i0, kRow, bar);
        bulk_load(b.fy, st.fy + i0, kRow, bar);
        bulk_load(b.h, st.h + i0, kRow, bar);
        bulk_load(b.act, actions + 2 * i0, 2 * kRow, bar);
        bulk_load(b.counter, st.counter + i0, kTile * 4, bar);
    };
    // the service group's draws for one tile: kTile / kSvc independent envs per lane
    auto draw_tile = [&](int buf, int64_t tile) {
        const int64_t i0 = tile * kTile;
#pragma unroll
        for (int q = 0; q < kTile / kSvc; ++q) {
            const int j = lane_s + q * kSvc;
            PhiloxNoise nz;
            nz.seek(nv.env_base + (uint64_t)(i0 + j), off);
            float z8[8];
            nz.draw8(p, z8);
            sm.z[buf][0][j] = make_float4(z8[0], z8[1], z8[2], z8[3]);
            sm.z[buf][1][j] = make_float4(z8[4], z8[5], z8[6], z8[7]);
        }
    };

    if (svc) {
        if (elect) {
            for (int s = 0; s < kStagesIn; ++s) {
                const int64_t tile = first + (int64_t)s * stride;
                if (tile < n_tiles) issue_loads(s, tile);
            }
        }
        if (first < n_tiles) draw_tile(0, first);
    }
    __syncthreads();                                              // z[0] is ready

    int it = 0;
    for (int64_t tile = first; tile < n_tiles; tile += stride, ++it) {
        const int s = it % kStagesIn;
        const uint32_t parity = (uint32_t)(it / kStagesIn) & 1u;
        const int so = it % kStagesOut;
        const int64_t i0 = tile * kTile;
        TileOut<T, kTile>& bo = sm.out[so];

        if (svc) {
            if (elect) bulk_wait_read<kStagesOut - 1>();          // out[so] (used kStagesOut tiles ago) has been read out
            __syncthreads();                                      // [A] in[s] consumed by the integrators, out[so] free
            if (elect) {
                const int64_t nxt = tile + (int64_t)kStagesIn * stride;
                if (nxt < n_tiles) issue_loads(s, nxt);
            }
            if (tile + stride < n_tiles) draw_tile((it + 1) & 1, tile + stride);   // read by the integrators after [B]
        } else {
            mbar_wait(&sm.full[s], parity);
            Env e;
            const TileIn<T, kTile>& bi = sm.in[s];
            e.x = (double)bi.x[tid]; e.y = (double)bi.y[tid]; e.fx = (double)bi.fx[tid]; e.fy = (double)bi.fy[tid];
            const T h_raw = bi.h[tid];
            e.counter = bi.counter[tid]; e.status = 0; e.spx = e.spy = 0.0;
            double f_t, al;
            if constexpr (sizeof(T) == 8) { const double2 a2 = reinterpret_cast<const double2*>(bi.act)[tid]; f_t = a2.x; al = a2.y; }
            else { const float2 a2 = reinterpret_cast<const float2*>(bi.act)[tid]; f_t = a2.x; al = a2.y; }
            __syncthreads();                                      // [A]

            const double t = time_at(tv, e.counter, p.dt);
            const double tb = t + p.dt, tb2 = tb + p.dt;
            e.h = decode_h<T>(h_raw, tb - t);
            const float4 za = sm.z[it & 1][0][tid], zb = sm.z[it & 1][1][tid];
            const float z8[8] = {za.x, za.y, za.z, za.w, zb.x, zb.y, zb.z, zb.w};
            PhiloxNoise nz;                                       // only the rare multi-attempt path draws from it
            nz.seek(nv.env_base + (uint64_t)(i0 + tid), off);
            nz.blk += 2;
            e.counter += 1;                                       // MR_env.py:80
            sim_step_drawn(e, t, tb, tb2, f_t, al, p, nz, z8);
            const Observation o = observe(e, p);
            double d_out = o.d, il_next = tb2 - tb;
            if (p.auto_reset && o.done) {                         // reported obs = first obs of the new episode
                int32_t cur = 0; int ov = 0;
                auto_reset_env<MR_NOISE_PHILOX, false>(e, nv, n_total, i0 + tid, cur, off, p, ov);
                d_out = sqrt(e.x * e.x + e.y * e.y);
                il_next = p.dt;
            }
            bo.x[tid] = (T)e.x; bo.y[tid] = (T)e.y; bo.fx[tid] = (T)e.fx; bo.fy[tid] = (T)e.fy;
            bo.h[tid] = encode_h<T>(e.h, il_next);
            bo.counter[tid] = e.counter;
            bo.d[tid] = (T)d_out; bo.rew[tid] = (T)o.rew; bo.done[tid] = o.done ? 1 : 0;
            if (out.sp) { bo.spx[tid] = (T)e.spx; bo.spy[tid] = (T)e.spy; }
            if (e.status) st.status[i0 + tid] |= (uint8_t)e.status;   // rare: sticky flags, plain store
            fence_async_smem();
        }
        __syncthreads();                                          // [B] tile results complete in out[so], next z ready
        if (svc && elect) {
            bulk_store(st.x + i0, bo.x, kRow);
            bulk_store(st.y + i0, bo.y, kRow);
            bulk_store(st.fx + i0, bo.fx, kRow);
            bulk_store(st.fy + i0, bo.fy, kRow);
            bulk_store(st.h + i0, bo.h, kRow);
            bulk_store(st.counter + i0, bo.counter, kTile * 4);
            if (out.obs) {
                bulk_store(out.obs + i0, bo.x, kRow);
                bulk_store(out.obs + out.stride + i0, bo.y, kRow);
                if (out.goal) {                                   // goal = (0,0), MR_env.py:57
                    bulk_store(out.obs + 2 * out.stride + i0, sm.zero, kRow);
                    bulk_store(out.obs + 3 * out.stride + i0, sm.zero, kRow);
                }
                bulk_store(out.obs + 4 * out.stride + i0, bo.d, kRow);
            }
            if (out.rew) bulk_store(out.rew + i0, bo.rew, kRow);
            if (out.done) bulk_store(out.done + i0, bo.done, kTile);
            if (out.sp) {
                bulk_store(out.sp + i0, bo.spx, kRow);
                bulk_store(out.sp + out.stride + i0, bo.spy, kRow);
            }
            bulk_commit();
        }
    }
    if (svc && elect) bulk_wait_read<0>();                        // smem must outlive the last bulk stores
}

}  // namespace mr
