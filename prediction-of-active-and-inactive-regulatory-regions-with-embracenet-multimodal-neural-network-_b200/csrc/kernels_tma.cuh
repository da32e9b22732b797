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
    const bf16 *y, *a, *ga;                 // pre-BN conv output [B, Lc, C]; pooled output and its gradient [B, Lp, C]
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
//   MODE 0: only the two per-channel BatchNorm reductions  sum(dz), sum(dz * xhat)             (reads y, a, d(a))
//   MODE 1: recomputes dz and applies  dy = gamma * rstd * (dz - mean(dz) - xhat * mean(dz * xhat))   (reads y, a, d(a); writes dy)
// dz is never written to memory (the two-kernel version wrote and re-read it).
//
// Windows are 5 position PAIRS wide and advance by one pair, so everything is done per pair: pm[p] = max(z[2p], z[2p+1]),
// sel[p] = (z[2p+1] > z[2p]).  The max-pool gradient of window j goes to the FIRST maximum (strict '>'), i.e. to position
// 2p + sel[p] of the first pair p in j..j+4 with the largest pm -- and within a pair only that one position can ever
// receive gradient, so a 5-deep ring of ONE accumulator per pair replaces the 10-position ring.  Gradient flows only
// where the pooled output survived Dropout and ReLU (a > 0).  The ring is rotated by unrolling five windows.
// One CTA = one sample at a time (single shared-memory stage); 3-4 CTAs per SM overlap each other's loads.
template <int MODE>
__global__ void __launch_bounds__(KT_THREADS, 3)
pool_bn_bwd_tma_kernel(const PoolBwdArgs g) {
    extern __shared__ uint8_t kt_smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)kt_smem_raw + 127) & ~(uintptr_t)127);
    const int C = g.C, Lc = g.Lc, Lp = g.Lp;
    const uint32_t y_bytes = (uint32_t)Lc * C * 2, a_bytes = (uint32_t)Lp * C * 2;
    const uint32_t y_sz = (y_bytes + 127) & ~127u, a_sz = (a_bytes + 127) & ~127u;
    uint64_t* full = (uint64_t*)(smem + y_sz + 2 * a_sz);
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
    // positions [i0, i1) are owned by this thread; the windows that touch them are j_first .. j_last
    const int i0 = seg * g.P, i1 = min(Lc, i0 + g.P);
    const int j_first = max(0, (i0 >> 1) - 4);
    const int j_last = i1 >= Lc ? Lp - 1 : min(Lp - 1, (i1 >> 1) - 1);
    // pairs behind the last window (2*Lp .. 2*Lp+7) are finished once window Lp-1 is done: keep emitting through them
    const int j_emit_last = (j_last == Lp - 1) ? min(Lp + 3, (i1 >> 1) - 1) : j_last;

    float acc0x = 0, acc0y = 0, acc1x = 0, acc1y = 0;      // MODE 0: sum dz, sum dz*xhat;  MODE 1: sum dy (acc0)

    int it = 0;
    for (int b = blockIdx.x; b < g.B; b += gridDim.x, ++it) {
        __syncthreads();                       // everyone is done with the previous sample's tile (and the barrier init is visible)
        if (t == 0) {
            mbar_expect_tx(full, y_bytes + 2 * a_bytes);
            bulk_load(smem, g.y + (size_t)b * Lc * C, y_bytes, full);
            bulk_load(smem + y_sz, g.a + (size_t)b * Lp * C, a_bytes, full);
            bulk_load(smem + y_sz + a_sz, g.ga + (size_t)b * Lp * C, a_bytes, full);
        }
        mbar_wait(full, (uint32_t)it & 1u);
        const bf16* ys = (const bf16*)smem + c;
        const bf16* as = (const bf16*)(smem + y_sz) + c;
        const bf16* gs = (const bf16*)(smem + y_sz + a_sz) + c;
        bf16* od = MODE ? g.dy + (size_t)b * Lc * C + c : nullptr;

        if (MODE == 0) {
            // Reductions only.  Every routed gradient lands on exactly one position, so
            //   sum(dz) = sum_j gg_j      and      sum(dz * xhat) = sum_j gg_j * xhat[argmax_j]:
            // no per-position accumulators and no warm-up windows; the thread owns WINDOWS [w0, w1).
            const int w0 = i0 >> 1, w1 = min(Lp, (i0 + g.P) >> 1);
            if (active && w0 < w1) {
                float2 pmr[5], ysr[5];            // pair maximum of z and the y value at the pair's first maximum
                auto load_pair = [&](int p, int slot) {
                    const float2 u0 = lds2(ys + (size_t)(2 * p) * C), u1 = lds2(ys + (size_t)(2 * p + 1) * C);
                    const float z0x = fmaf(u0.x, scx, shx), z1x = fmaf(u1.x, scx, shx);
                    const float z0y = fmaf(u0.y, scy, shy), z1y = fmaf(u1.y, scy, shy);
                    const bool sx = z1x > z0x, sy = z1y > z0y;
                    pmr[slot] = make_float2(sx ? z1x : z0x, sy ? z1y : z0y);
                    ysr[slot] = make_float2(sx ? u1.x : u0.x, sy ? u1.y : u0.y);
                };
#pragma unroll
                for (int m = 0; m < 4; ++m) load_pair(w0 + m, m);
                pmr[4] = ysr[4] = make_float2(0.f, 0.f);
                for (int jb = w0; jb < w1; jb += 5) {
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        const int j = jb + k;
                        if (j < w1) {
                            load_pair(j + 4, (k + 4) % 5);
                            const float2 av = lds2(as + (size_t)j * C), gv = lds2(gs + (size_t)j * C);
                            float bmx = pmr[k % 5].x, byx = ysr[k % 5].x, bmy = pmr[k % 5].y, byy = ysr[k % 5].y;
#pragma unroll
                            for (int m = 1; m < 5; ++m) {
                                const float2 pm = pmr[(k + m) % 5], yy = ysr[(k + m) % 5];
                                if (pm.x > bmx) { bmx = pm.x; byx = yy.x; }
                                if (pm.y > bmy) { bmy = pm.y; byy = yy.y; }
                            }
                            const float ggx = av.x > 0.f ? gv.x * inv_keep : 0.f, ggy = av.y > 0.f ? gv.y * inv_keep : 0.f;
                            acc0x += ggx; acc0y += ggy;
                            acc1x = fmaf(ggx, (byx - mux) * rsx, acc1x);
                            acc1y = fmaf(ggy, (byy - muy) * rsy, acc1y);
                        }
                    }
                }
            }
        } else if (active && i0 < Lc) {
            // ring slot of pair p: (p - j_first) % 5
            float2 y0r[5], y1r[5], pmr[5], dpr[5];
            bool s0r[5], s1r[5];
            auto load_pair = [&](int p, int slot) {
                const float2 u0 = lds2(ys + (size_t)(2 * p) * C), u1 = lds2(ys + (size_t)(2 * p + 1) * C);
                const float z0x = fmaf(u0.x, scx, shx), z1x = fmaf(u1.x, scx, shx);
                const float z0y = fmaf(u0.y, scy, shy), z1y = fmaf(u1.y, scy, shy);
                y0r[slot] = u0; y1r[slot] = u1;
                s0r[slot] = z1x > z0x; s1r[slot] = z1y > z0y;
                pmr[slot] = make_float2(fmaxf(z0x, z1x), fmaxf(z0y, z1y));
                dpr[slot] = make_float2(0.f, 0.f);
            };
            auto emit_pair = [&](int p, int slot) {
                const float2 u0 = y0r[slot], u1 = y1r[slot], d = dpr[slot];
                const bool sx = s0r[slot], sy = s1r[slot];
                // dy = gs * (dz - dbn - (y - mu) * rs * dgn) = gs * dz - k1 * y - k0;  the two-kernel version stored dz in bf16
                // between its passes: keep that rounding point
                const float dxr = gsx * bf16_rt(d.x), dyr = gsy * bf16_rt(d.y);
                float2 o0, o1;
                o0.x = fmaf(-k1x, u0.x, (sx ? 0.f : dxr) - k0x);
                o1.x = fmaf(-k1x, u1.x, (sx ? dxr : 0.f) - k0x);
                o0.y = fmaf(-k1y, u0.y, (sy ? 0.f : dyr) - k0y);
                o1.y = fmaf(-k1y, u1.y, (sy ? dyr : 0.f) - k0y);
                sts2(od + (size_t)(2 * p) * C, o0);          // plain global stores: 128 contiguous bytes per warp and row
                sts2(od + (size_t)(2 * p + 1) * C, o1);
                acc0x += o0.x + o1.x; acc0y += o0.y + o1.y;
            };
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                if (j_first + m <= Lp + 3) load_pair(j_first + m, m);
                else { y0r[m] = y1r[m] = dpr[m] = make_float2(0.f, 0.f); pmr[m] = make_float2(-INFINITY, -INFINITY); s0r[m] = s1r[m] = false; }
            }
            y0r[4] = y1r[4] = dpr[4] = make_float2(0.f, 0.f); pmr[4] = make_float2(-INFINITY, -INFINITY); s0r[4] = s1r[4] = false;
            for (int jb = j_first; jb <= j_emit_last; jb += 5) {
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const int j = jb + k;
                    if (j <= j_emit_last) {
                        if (j <= j_last) {
                            load_pair(j + 4, (k + 4) % 5);           // window j = pairs j .. j+4 = slots k .. k+4 (mod 5)
                            const float2 av = lds2(as + (size_t)j * C), gv = lds2(gs + (size_t)j * C);
                            int bx = 0, by = 0;
                            float bmx = pmr[k % 5].x, bmy = pmr[k % 5].y;
#pragma unroll
                            for (int m = 1; m < 5; ++m) {
                                const float2 pm = pmr[(k + m) % 5];
                                if (pm.x > bmx) { bmx = pm.x; bx = m; }
                                if (pm.y > bmy) { bmy = pm.y; by = m; }
                            }
                            const float ggx = av.x > 0.f ? gv.x * inv_keep : 0.f, ggy = av.y > 0.f ? gv.y * inv_keep : 0.f;
#pragma unroll
                            for (int m = 0; m < 5; ++m) {
                                dpr[(k + m) % 5].x += (m == bx) ? ggx : 0.f;
                                dpr[(k + m) % 5].y += (m == by) ? ggy : 0.f;
                            }
                        }
                        if (2 * j >= i0) emit_pair(j, k % 5);        // pair j can no longer change
                    }
                }
            }
            // positions that were never inside a window (at most one, for odd Lc): dz = 0
            for (int pos = max(i0, 2 * Lp + 8); pos < i1; ++pos) {
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
        else { atomicAdd(&g.dbias[2 * t], sx); atomicAdd(&g.dbias[2 * t + 1], sy); }
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

struct PoolFwdArgs {
    const bf16* y;                          // [B, Lc, C]
    const float *scale, *shift;
    bf16* a;                                // [B, Lp, C]
    const RngState* rng;
    int64_t row_offset;
    uint32_t rng_stream;
    int B, Lc, Lp, C, nseg, P;              // nseg segments of P pooled positions
    float drop_p;
};

// Forward: a = Dropout(MaxPool1d(10, 2)(ReLU(scale * y + shift))).  DROP: 0 none, 2 Philox (replayed uniforms stay on the
// register-streaming kernel).  thread = (channel pair, segment of pooled positions); window max = max of 5 pair maxima.
template <int DROP>
__global__ void __launch_bounds__(KT_THREADS, 4)
bn_relu_pool_drop_fwd_tma_kernel(const PoolFwdArgs g) {
    extern __shared__ uint8_t kt_smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)kt_smem_raw + 127) & ~(uintptr_t)127);
    const int C = g.C, Lc = g.Lc, Lp = g.Lp;
    const uint32_t y_bytes = (uint32_t)Lc * C * 2;
    const uint32_t y_sz = (y_bytes + 127) & ~127u;
    uint64_t* full = (uint64_t*)(smem + y_sz);

    const int t = threadIdx.x;
    const int pairs = C >> 1;
    const int cp = t % pairs, seg = t / pairs;
    const bool active = seg < g.nseg;
    const int c = 2 * cp;
    if (t == 0) {
        mbar_init(full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    float scx = 0, scy = 0, shx = 0, shy = 0;
    if (active) { scx = g.scale[c]; scy = g.scale[c + 1]; shx = g.shift[c]; shy = g.shift[c + 1]; }
    const float inv_keep = DROP ? 1.f / (1.f - g.drop_p) : 1.f;
    RngState rs = {0, 0};
    if (DROP == 2) rs = *g.rng;
    const int j0 = seg * g.P, j1 = min(Lp, j0 + g.P);

    int it = 0;
    for (int b = blockIdx.x; b < g.B; b += gridDim.x, ++it) {
        __syncthreads();
        if (t == 0) {
            mbar_expect_tx(full, y_bytes);
            bulk_load(smem, g.y + (size_t)b * Lc * C, y_bytes, full);
        }
        mbar_wait(full, (uint32_t)it & 1u);
        const bf16* ys = (const bf16*)smem + c;
        bf16* od = g.a + (size_t)b * Lp * C + c;
        if (active && j0 < Lp) {
            // pair maxima m[i] = max(relu(z[2i]), relu(z[2i+1])); pooled[j] = max(m[j .. j+4])
            float2 w0, w1, w2, w3;
            auto pm = [&](int i) {
                const float2 u0 = lds2(ys + (size_t)(2 * i) * C), u1 = lds2(ys + (size_t)(2 * i + 1) * C);
                float2 m;
                m.x = fmaxf(fmaxf(fmaf(u0.x, scx, shx), fmaf(u1.x, scx, shx)), 0.f);
                m.y = fmaxf(fmaxf(fmaf(u0.y, scy, shy), fmaf(u1.y, scy, shy)), 0.f);
                return m;
            };
            w0 = pm(j0); w1 = pm(j0 + 1); w2 = pm(j0 + 2); w3 = pm(j0 + 3);
            uint4 blk0 = make_uint4(0, 0, 0, 0);
            for (int j = j0; j < j1; ++j) {
                const float2 m = pm(j + 4);
                float2 r;
                r.x = fmaxf(fmaxf(fmaxf(w0.x, w1.x), fmaxf(w2.x, w3.x)), m.x);
                r.y = fmaxf(fmaxf(fmaxf(w0.y, w1.y), fmaxf(w2.y, w3.y)), m.y);
                if (DROP == 2) {
                    if ((j & 3) == 0 || j == j0) {
                        blk0 = rng_cnn_block(rs, g.rng_stream, (uint64_t)(g.row_offset + b), C, cp, Lp, j >> 2);
                    }
                    const float ux = rng_cnn_u16(blk0, (uint32_t)j & 3u, 0u);
                    const float uy = rng_cnn_u16(blk0, (uint32_t)j & 3u, 1u);
                    r.x = (ux >= g.drop_p) ? r.x * inv_keep : 0.f;
                    r.y = (uy >= g.drop_p) ? r.y * inv_keep : 0.f;
                }
                sts2(od + (size_t)j * C, r);
                w0 = w1; w1 = w2; w2 = w3; w3 = m;
            }
        }
    }
}

// host-side geometry shared by both kernels: threads = pairs * nseg <= 256
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
    return y_sz + 2 * a_sz + 64 + 128;
}
inline size_t kt_fwd_smem(int Lc, int C) {
    const size_t y_sz = ((size_t)Lc * C * 2 + 127) & ~(size_t)127;
    return y_sz + 64 + 128;
}
// CTAs of `smem` dynamic bytes that fit one SM (228 KB, 1 KB reserved per CTA), capped
inline int kt_ctas_per_sm(size_t smem, int cap) { return std::max(1, std::min(cap, (int)((228 * 1024) / (smem + 1024 + 2048)))); }

}  // namespace emb
