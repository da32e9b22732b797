// Bulk-copy (TMA) staged versions of the HBM-bound CNN-stack kernels (bf16 activations, C % 8 == 0, ld == C).
//
// In the channels-last layout one sample's [L, C] block is CONTIGUOUS in memory, so a persistent CTA moves whole
// samples with one cp.async.bulk per tensor (global -> shared, completion on an mbarrier) two samples ahead of the
// arithmetic, and writes its result back with one cp.async.bulk (shared -> global).  Every byte crosses HBM exactly
// once, in 16..32 KB requests that keep ~100 KB in flight per SM (the register-streaming kernels in
// kernels_fast.cuh kept 16 KB in flight and ran at 1.7 TB/s).  The arithmetic reads shared memory: thread =
// (channel pair, position segment), a warp reads 128 contiguous bytes per row.
//
//   reference ops (CNN_pre.py:39-51): BatchNorm1d (batch statistics) -> ReLU -> MaxPool1d(10, 2) -> Dropout
#pragma once
#include "common.cuh"
#include "kernels.cuh"
#include "kernels_fast.cuh"
#include "gemm_tc.cuh"

namespace emb {

__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float2 lds2(const bf16* p) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p)); }
__device__ __forceinline__ void sts2(bf16* p, float2 v) { *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v.x, v.y); }
__device__ __forceinline__ float bf16_rt(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

constexpr int KT_THREADS = 256;

struct PoolBwdArgs {
    const bf16 *y, *ga;                     // pre-BN conv output [B, Lc, C]; gradient of the pooled output [B, Lp, C]
    const uint8_t* amax;                    // [B, Lp, C] from the forward: offset 0..9 of the window's first maximum, 255 = no gradient
    const float *scale, *shift, *mean, *rstd, *gamma;
    const double* bstats_in;                // MODE 1: [sum dz | sum dz * xhat] (already global under data parallelism)
    double* bstats_out;                     // MODE 0
    bf16* dy;                               // MODE 1: gradient w.r.t. the conv output
    float* dbias;                           // MODE 1: conv bias gradient (sum of dy; analytically zero behind BatchNorm)
    double n;                               // MODE 1: elements per channel of the (global) batch
    int B, Lc, Lp, C, nseg, P;              // nseg position segments of P (even) positions each
    float drop_p;
};

// Backward of Dropout -> MaxPool1d(10, 2) -> ReLU -> BatchNorm for one conv layer.
//   MODE 0: only the two per-channel BatchNorm reductions  sum(dz), sum(dz * xhat)             (reads y, d(a), amax)
//   MODE 1: recomputes dz and applies  dy = gamma * rstd * (dz - mean(dz) - xhat * mean(dz * xhat))   (reads the same; writes dy)
// dz is never written to memory (the two-kernel version wrote and re-read it), and the arg-max is not recomputed: the
// forward kernel stored, per pooled element, WHERE its gradient goes (first maximum of the window, strict '>'; nowhere
// if the maximum was not positive or the element was dropped).
//   MODE 0: every routed gradient lands on exactly one position, so sum(dz) = sum_j g_j and
//           sum(dz * xhat) = sum_j g_j * xhat[2j + code_j]: one dynamic-row shared-memory read per window.
//   MODE 1: windows are 5 position pairs wide, so pair p receives from windows p-4 .. p; a 5-deep register ring of
//           (target pair, g) is rotated by unrolling five pairs; inside a pair only its first maximum can receive
//           gradient, so one accumulator per pair suffices.  No warm-up windows are processed, only loaded.
// One CTA = one sample at a time (three bulk copies into one shared-memory stage); 3 CTAs per SM overlap each other's loads.
template <int MODE>
__global__ void __launch_bounds__(KT_THREADS, 3)
pool_bn_bwd_tma_kernel(const PoolBwdArgs g) {
    extern __shared__ uint8_t kt_smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)kt_smem_raw + 127) & ~(uintptr_t)127);
    const int C = g.C, Lc = g.Lc, Lp = g.Lp;
    const uint32_t y_bytes = (uint32_t)Lc * C * 2, a_bytes = (uint32_t)Lp * C * 2, i_bytes = ((uint32_t)Lp * C + 15) & ~15u;
    const uint32_t y_sz = (y_bytes + 127) & ~127u, a_sz = (a_bytes + 127) & ~127u, i_sz = (i_bytes + 127) & ~127u;
    uint64_t* full = (uint64_t*)(smem + y_sz + a_sz + i_sz);
    __shared__ float red[2][KT_THREADS];

    const int t = threadIdx.x;
    const int pairs = C >> 1;
    const int cp = t % pairs, seg = t / pairs;
    const bool active = seg < g.nseg;
    const int c = 2 * cp;

    if (t == 0) {
        mbar_init(full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    float scx = 0, scy = 0, shx = 0, shy = 0, mux = 0, muy = 0, rsx = 0, rsy = 0;
    float gsx = 0, gsy = 0, dbnx = 0, dbny = 0, dgnx = 0, dgny = 0;
    if (active) {
        scx = g.scale[c]; scy = g.scale[c + 1]; shx = g.shift[c]; shy = g.shift[c + 1];
        mux = g.mean[c]; muy = g.mean[c + 1]; rsx = g.rstd[c]; rsy = g.rstd[c + 1];
        if (MODE) {
            gsx = g.gamma[c] * rsx; gsy = g.gamma[c + 1] * rsy;
            dbnx = (float)(g.bstats_in[c] / g.n); dbny = (float)(g.bstats_in[c + 1] / g.n);
            dgnx = (float)(g.bstats_in[C + c] / g.n); dgny = (float)(g.bstats_in[C + c + 1] / g.n);
        }
    }
    const float k1x = gsx * rsx * dgnx, k1y = gsy * rsy * dgny;
    const float k0x = gsx * dbnx - k1x * mux, k0y = gsy * dbny - k1y * muy;
    const float inv_keep = g.drop_p > 0.f ? 1.f / (1.f - g.drop_p) : 1.f;
    const int i0 = seg * g.P, i1 = min(Lc, i0 + g.P);      // positions [i0, i1) are owned by this thread (i0 even)

    float acc0x = 0, acc0y = 0, acc1x = 0, acc1y = 0;      // MODE 0: sum dz, sum dz*xhat;  MODE 1: sum dy (acc0)

    int it = 0;
    for (int b = blockIdx.x; b < g.B; b += gridDim.x, ++it) {
        __syncthreads();                       // everyone is done with the previous sample's tile (and the barrier init is visible)
        if (t == 0) {
            mbar_expect_tx(full, y_bytes + a_bytes + i_bytes);
            bulk_load(smem, g.y + (size_t)b * Lc * C, y_bytes, full);
            bulk_load(smem + y_sz, g.ga + (size_t)b * Lp * C, a_bytes, full);
            bulk_load(smem + y_sz + a_sz, g.amax + (size_t)b * Lp * C, i_bytes, full);
        }
        mbar_wait(full, (uint32_t)it & 1u);
        const bf16* ys = (const bf16*)smem + c;
        const bf16* gs = (const bf16*)(smem + y_sz) + c;
        const uint8_t* is = smem + y_sz + a_sz + c;
        bf16* od = MODE ? g.dy + (size_t)b * Lc * C + c : nullptr;

        if (MODE == 0) {
            const int w0 = i0 >> 1, w1 = min(Lp, (i0 + g.P) >> 1);      // windows [w0, w1)
            if (active) {
                for (int j = w0; j < w1; ++j) {
                    const uint32_t code = *reinterpret_cast<const uint16_t*>(is + (size_t)j * C);
                    const float2 gv = lds2(gs + (size_t)j * C);
                    const uint32_t cx = code & 0xFFu, cy = code >> 8;
                    if (cx != 255u) {
                        const float yv = __bfloat162float(ys[(size_t)(2 * j + (int)cx) * C]);
                        const float gg = gv.x * inv_keep;
                        acc0x += gg;
                        acc1x = fmaf(gg, (yv - mux) * rsx, acc1x);
                    }
                    if (cy != 255u) {
                        const float yv = __bfloat162float(ys[(size_t)(2 * j + (int)cy) * C + 1]);
                        const float gg = gv.y * inv_keep;
                        acc0y += gg;
                        acc1y = fmaf(gg, (yv - muy) * rsy, acc1y);
                    }
                }
            }
        } else if (active && i0 < Lc) {
            const int p0 = i0 >> 1, p1 = i1 >> 1;                        // whole pairs [p0, p1); an odd last position is handled below
            // ring slot (j - p0 + 5) % 5 holds window j: the pair it sends its gradient to (or -1) and the gradient
            int tx[5], ty[5];
            float2 gr[5];
            auto load_win = [&](int j, int slot) {
                if (j >= 0 && j < Lp) {
                    const uint32_t code = *reinterpret_cast<const uint16_t*>(is + (size_t)j * C);
                    const uint32_t cx = code & 0xFFu, cy = code >> 8;
                    gr[slot] = lds2(gs + (size_t)j * C);
                    tx[slot] = cx != 255u ? j + (int)(cx >> 1) : -1;
                    ty[slot] = cy != 255u ? j + (int)(cy >> 1) : -1;
                } else { tx[slot] = ty[slot] = -1; gr[slot] = make_float2(0.f, 0.f); }
            };
#pragma unroll
            for (int m = 1; m <= 4; ++m) load_win(p0 - m, 5 - m);         // windows p0-4 .. p0-1 -> slots 1..4
            for (int pb = p0; pb < p1; pb += 5) {
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const int p = pb + k;
                    if (p < p1) {
                        load_win(p, k);                                  // window p overwrites the slot of window p-5
                        float dx = 0.f, dyv = 0.f;
#pragma unroll
                        for (int m = 0; m < 5; ++m) {
                            dx += tx[m] == p ? gr[m].x : 0.f;
                            dyv += ty[m] == p ? gr[m].y : 0.f;
                        }
                        const float2 u0 = lds2(ys + (size_t)(2 * p) * C), u1 = lds2(ys + (size_t)(2 * p + 1) * C);
                        // the pair's first maximum (the only position of the pair that can receive gradient)
                        const bool sx = fmaf(u1.x, scx, shx) > fmaf(u0.x, scx, shx), sy = fmaf(u1.y, scy, shy) > fmaf(u0.y, scy, shy);
                        // dy = gs * (dz - dbn - (y - mu) * rs * dgn) = gs * dz - k1 * y - k0;  the two-kernel version stored dz in
                        // bf16 between its passes: keep that rounding point
                        const float dxr = gsx * bf16_rt(dx * inv_keep), dyr = gsy * bf16_rt(dyv * inv_keep);
                        float2 o0, o1;
                        o0.x = fmaf(-k1x, u0.x, (sx ? 0.f : dxr) - k0x);
                        o1.x = fmaf(-k1x, u1.x, (sx ? dxr : 0.f) - k0x);
                        o0.y = fmaf(-k1y, u0.y, (sy ? 0.f : dyr) - k0y);
                        o1.y = fmaf(-k1y, u1.y, (sy ? dyr : 0.f) - k0y);
                        sts2(od + (size_t)(2 * p) * C, o0);              // plain global stores: 128 contiguous bytes per warp and row
                        sts2(od + (size_t)(2 * p + 1) * C, o1);
                        acc0x += o0.x + o1.x; acc0y += o0.y + o1.y;
                    }
                }
            }
            if ((i1 & 1) && i1 == Lc) {       // odd Lc: the last position is never inside a window, dz = 0
                const int pos = Lc - 1;
                const float2 u = lds2(ys + (size_t)pos * C);
                float2 o;
                o.x = fmaf(-k1x, u.x, -k0x);
                o.y = fmaf(-k1y, u.y, -k0y);
                sts2(od + (size_t)pos * C, o);
                acc0x += o.x; acc0y += o.y;
            }
        }
    }
    // per-channel totals: fold the segments of a channel pair, one atomic per channel and CTA
    __syncthreads();
    red[0][t] = acc0x; red[1][t] = acc0y;
    __syncthreads();
    if (t < pairs) {
        float sx = 0, sy = 0;
        for (int s = 0; s < g.nseg; ++s) { sx += red[0][t + s * pairs]; sy += red[1][t + s * pairs]; }
        if (MODE == 0) { atomicAdd(&g.bstats_out[2 * t], (double)sx); atomicAdd(&g.bstats_out[2 * t + 1], (double)sy); }
        else if (g.dbias) { atomicAdd(&g.dbias[2 * t], sx); atomicAdd(&g.dbias[2 * t + 1], sy); }
    }
    if (MODE == 0) {
        __syncthreads();
        red[0][t] = acc1x; red[1][t] = acc1y;
        __syncthreads();
        if (t < pairs) {
            float sx = 0, sy = 0;
            for (int s = 0; s < g.nseg; ++s) { sx += red[0][t + s * pairs]; sy += red[1][t + s * pairs]; }
            atomicAdd(&g.bstats_out[C + 2 * t], (double)sx);
            atomicAdd(&g.bstats_out[C + 2 * t + 1], (double)sy);
        }
    }
}

// host-side geometry: threads = pairs * nseg <= 256
inline void kt_segments(int C, int L, int unit, int* nseg, int* P) {
    const int pairs = C / 2;
    int ns = std::max(1, KT_THREADS / pairs);
    int p = round_up(cdiv(L, ns), unit);
    ns = cdiv(L, p);
    *nseg = ns;
    *P = p;
}
inline int kt_max_smem() { return tc_max_smem() - 4096; }      // the kernels also hold 2 KB of static shared memory
inline bool kt_ok(int C, int ld) { return (C % 8) == 0 && ld == C && C / 2 <= KT_THREADS; }
inline size_t kt_bwd_smem(int Lc, int Lp, int C) {
    const size_t y_sz = ((size_t)Lc * C * 2 + 127) & ~(size_t)127, a_sz = ((size_t)Lp * C * 2 + 127) & ~(size_t)127;
    const size_t i_sz = ((((size_t)Lp * C + 15) & ~(size_t)15) + 127) & ~(size_t)127;
    return y_sz + a_sz + i_sz + 64 + 128;
}
// CTAs of `smem` dynamic bytes that fit one SM (228 KB, 1 KB reserved per CTA), capped
inline int kt_ctas_per_sm(size_t smem, int cap) { return std::max(1, std::min(cap, (int)((228 * 1024) / (smem + 1024 + 2048)))); }

}  // namespace emb
