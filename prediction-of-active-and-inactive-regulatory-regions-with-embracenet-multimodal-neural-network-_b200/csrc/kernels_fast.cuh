// Bandwidth-oriented versions of the CNN-stack kernels for even / multiple-of-8 channel counts (every point of
// the reference's search space: 16..512 channels).  kernels.cuh keeps the shape-generic versions used for odd
// channel counts.  All activations are channels-last [B, L, C] (leading dimension ld, even).
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace emb {

// ---- two-channel load/store helpers -----------------------------------------------------------------------
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 ld2(const bf16* p) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p)); }
__device__ __forceinline__ void st2(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
__device__ __forceinline__ void st2(bf16* p, float2 v) { *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v.x, v.y); }
__device__ __forceinline__ void st8(float* p, const float* v) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8(bf16* p, const float* v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = make_uint4(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b),
                                              *reinterpret_cast<uint32_t*>(&c), *reinterpret_cast<uint32_t*>(&d));
}
template <typename T> __device__ __forceinline__ float round_like(float v);
template <> __device__ __forceinline__ float round_like<float>(float v) { return v; }
template <> __device__ __forceinline__ float round_like<bf16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// ---------------------------------------------------------------------------------------------
// K1 forward, vectorised: each thread produces 8 consecutive output channels of one position from the
// shared-memory weight-row table (two 16-byte table reads per tap), stores them as one 16-byte (bf16) or two
// 16-byte (fp32) words, and keeps the BatchNorm sum / sum-of-squares of its 8 channels in registers, so layer 0
// needs no separate statistics pass.  Requires C1 % 8 == 0 and 256 % (C1/8) == 0.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
onehot_conv_fwd_v8_kernel(const uint8_t* __restrict__ bases, const float* __restrict__ w, const float* __restrict__ bias,
                          T* __restrict__ y, double* __restrict__ stats, int B, int C1, int k, int ld) {
    extern __shared__ float smem[];
    float* tab = smem;                         // [k][4][C1]
    float* sb = tab + k * 4 * C1;              // [C1]
    float* red = sb + C1;                      // [256][16] reduction scratch
    uint8_t* sbase = (uint8_t*)(red + 256 * 16);
    const int p = (k - 1) / 2;
    const int groups = C1 / 8;
    for (int i = threadIdx.x; i < k * 4 * C1; i += blockDim.x) {
        int tap = i / (4 * C1), r = i - tap * 4 * C1, c = r / C1, o = r - c * C1;
        tab[i] = w[((size_t)o * 4 + c) * k + tap];
    }
    for (int i = threadIdx.x; i < C1; i += blockDim.x) sb[i] = bias[i];
    const int og = threadIdx.x % groups;       // constant per thread: 256 % groups == 0
    float s1[8], s2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < SEQ_LEN + 2 * p; i += blockDim.x) {
            int l = i - p;
            sbase[i] = (l >= 0 && l < SEQ_LEN) ? bases[(size_t)b * SEQ_LEN + l] : 4;
        }
        __syncthreads();
        for (int item = threadIdx.x; item < SEQ_LEN * groups; item += 256) {
            const int l = item / groups;
            float acc[8];
            const float4* b4 = reinterpret_cast<const float4*>(sb + og * 8);
            float4 t0 = b4[0], t1 = b4[1];
            acc[0] = t0.x; acc[1] = t0.y; acc[2] = t0.z; acc[3] = t0.w; acc[4] = t1.x; acc[5] = t1.y; acc[6] = t1.z; acc[7] = t1.w;
            for (int tap = 0; tap < k; ++tap) {
                int c = sbase[l + tap];
                if (c < 4) {
                    const float4* r4 = reinterpret_cast<const float4*>(tab + (tap * 4 + c) * C1 + og * 8);
                    float4 a0 = r4[0], a1 = r4[1];
                    acc[0] += a0.x; acc[1] += a0.y; acc[2] += a0.z; acc[3] += a0.w;
                    acc[4] += a1.x; acc[5] += a1.y; acc[6] += a1.z; acc[7] += a1.w;
                }
            }
            st8(y + ((size_t)b * SEQ_LEN + l) * ld + og * 8, acc);
            if (stats) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { float v = round_like<T>(acc[i]); s1[i] += v; s2[i] += v * v; }
            }
        }
    }
    if (stats) {   // threads with the same og hold partials of the same 8 channels
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; ++i) { red[threadIdx.x * 16 + i] = s1[i]; red[threadIdx.x * 16 + 8 + i] = s2[i]; }
        __syncthreads();
        for (int i = threadIdx.x; i < groups * 16; i += blockDim.x) {
            int g = i / 16, j = i - g * 16;
            double tot = 0;
            for (int t = g; t < 256; t += groups) tot += red[t * 16 + j];
            int c = g * 8 + (j & 7);
            atomicAdd(&stats[(j < 8 ? 0 : C1) + c], tot);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1 backward, histogram form: thread = (output channel pair, tap); the loop runs over SOURCE positions l', whose
// base is the same for the whole CTA, so the per-base accumulator is chosen by a uniform branch instead of four
// predicated adds.  The sample's dy tile is staged in shared memory once (coalesced) and re-read by the 15 taps
// from there.  Per-sample fp32 partials are folded into fp64 so that the heavy cancellation BatchNorm induces
// (sum_l dy = 0 per channel) does not cost accuracy.   Requires C1 even and (C1/2)*k <= 1024.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024)
onehot_conv_bwd_hist_kernel(const uint8_t* __restrict__ bases, const T* __restrict__ dy, float* __restrict__ dw, int B, int C1, int k, int ld) {
    extern __shared__ float smem[];
    float* dys = smem;                                   // [256][C1] fp32
    uint8_t* sbase = (uint8_t*)(dys + SEQ_LEN * C1);     // [256]
    const int p = (k - 1) / 2;
    const int pairs = C1 / 2;
    const int tap = threadIdx.x / pairs, op = threadIdx.x - tap * pairs;
    const bool active = tap < k;
    double d[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) d[i][c] = 0.0;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < SEQ_LEN * pairs; i += blockDim.x) {
            int l = i / pairs, o2 = i - l * pairs;
            float2 v = ld2(dy + ((size_t)b * SEQ_LEN + l) * ld + 2 * o2);
            dys[l * C1 + 2 * o2] = v.x;
            dys[l * C1 + 2 * o2 + 1] = v.y;
        }
        for (int i = threadIdx.x; i < SEQ_LEN; i += blockDim.x) sbase[i] = bases[(size_t)b * SEQ_LEN + i];
        __syncthreads();
        if (active) {
            float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
            // output position l reads source l' = l + tap - p  <=>  source l' feeds output l = l' - tap + p
            for (int ls = 0; ls < SEQ_LEN; ++ls) {
                const int c = sbase[ls];                                  // uniform across the CTA
                const int l = ls - tap + p;
                if (l < 0 || l >= SEQ_LEN) continue;
                const float2 v = *reinterpret_cast<const float2*>(dys + l * C1 + 2 * op);
                switch (c) {
                    case 0: a0[0] += v.x; a1[0] += v.y; break;
                    case 1: a0[1] += v.x; a1[1] += v.y; break;
                    case 2: a0[2] += v.x; a1[2] += v.y; break;
                    default: a0[3] += v.x; a1[3] += v.y; break;
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) { d[0][c] += a0[c]; d[1][c] += a1[c]; }
        }
    }
    if (active) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) atomicAdd(&dw[((size_t)(2 * op + i) * 4 + c) * k + tap], (float)d[i][c]);
    }
}

// ---------------------------------------------------------------------------------------------
// K2 forward, streaming: thread = (sample, channel pair) walks the positions once.  MaxPool1d(10, 2) over
// ReLU(BN(y)) is the max of 5 consecutive pair-maxima, kept in a 4-register window: every y element is read
// exactly once (the generic kernel reads it 5 times through L1).  Four position pairs (8 rows) are loaded per trip
// before any is used, so each thread keeps 8 independent 4-byte loads in flight.  DROP: 0 none, 1 replayed
// uniforms, 2 Philox.  Requires C even.
// ---------------------------------------------------------------------------------------------
template <typename T, int DROP>
__global__ void __launch_bounds__(256)
bn_relu_pool_drop_fwd_stream_kernel(const T* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                                    T* __restrict__ a, int B, int Lc, int Lp, int C, int ld, float drop_p,
                                    const float* __restrict__ drop_u, RngState const* rng, uint32_t rng_stream, int64_t row_offset,
                                    uint8_t* __restrict__ amax) {
    const int pairs = C / 2;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)B * pairs) return;
    const int cp = t % pairs, b = t / pairs;
    const int c = 2 * cp;
    const float2 sc = make_float2(scale[c], scale[c + 1]), sh = make_float2(shift[c], shift[c + 1]);
    const T* src = y + (size_t)b * Lc * ld + c;
    T* dst = a + (size_t)b * Lp * ld + c;
    const float inv_keep = DROP ? 1.f / (1.f - drop_p) : 1.f;
    RngState rs = {0, 0};
    if (DROP == 2) rs = *rng;
    uint4 blk0 = make_uint4(0, 0, 0, 0);                     // current Philox block of this channel pair
    float2 w0 = make_float2(0.f, 0.f), w1 = w0, w2 = w0, w3 = w0;   // relu folded in: window values start at 0
    uint32_t sx = 0, sy = 0;                 // bit q: the pair held in window slot q has its (first) maximum at its SECOND position
    const int n_pairs = Lp + 4;
    uint8_t* adst = amax ? amax + (size_t)b * Lp * C + c : nullptr;

    auto step = [&](int i, float2 u0, float2 u1) {
        const float z0x = fmaxf(fmaf(u0.x, sc.x, sh.x), 0.f), z1x = fmaxf(fmaf(u1.x, sc.x, sh.x), 0.f);
        const float z0y = fmaxf(fmaf(u0.y, sc.y, sh.y), 0.f), z1y = fmaxf(fmaf(u1.y, sc.y, sh.y), 0.f);
        float2 m;
        m.x = fmaxf(z0x, z1x);
        m.y = fmaxf(z0y, z1y);
        if (amax) { sx |= (z1x > z0x ? 1u : 0u) << 4; sy |= (z1y > z0y ? 1u : 0u) << 4; }
        if (i >= 4) {
            const int j = i - 4;
            float2 r;
            r.x = fmaxf(fmaxf(fmaxf(w0.x, w1.x), fmaxf(w2.x, w3.x)), m.x);
            r.y = fmaxf(fmaxf(fmaxf(w0.y, w1.y), fmaxf(w2.y, w3.y)), m.y);
            // where the max-pool gradient will go: offset (0..9) of the FIRST maximum inside the window, 255 = nowhere
            // (window maximum not positive, or dropped below).  Consumed by pool_bn_bwd_tma_kernel.
            uint32_t cx = 255u, cy = 255u;
            if (amax) {
                // first slot that attains the maximum r: independent equality tests instead of a running-maximum chain
                int bx = 4, by = 4;
                bx = w3.x == r.x ? 3 : bx; by = w3.y == r.y ? 3 : by;
                bx = w2.x == r.x ? 2 : bx; by = w2.y == r.y ? 2 : by;
                bx = w1.x == r.x ? 1 : bx; by = w1.y == r.y ? 1 : by;
                bx = w0.x == r.x ? 0 : bx; by = w0.y == r.y ? 0 : by;
                if (r.x > 0.f) cx = 2u * bx + ((sx >> bx) & 1u);
                if (r.y > 0.f) cy = 2u * by + ((sy >> by) & 1u);
            }
            if (DROP) {
                float ux, uy;
                if (DROP == 1) {
                    ux = drop_u[((size_t)b * C + c) * Lp + j];
                    uy = drop_u[((size_t)b * C + c + 1) * Lp + j];
                } else {
                    if ((j & 3) == 0) {                       // uniform across the warp: every thread refreshes at the same j
                        blk0 = rng_cnn_block(rs, rng_stream, (uint64_t)(row_offset + b), C, cp, Lp, j >> 2);
                    }
                    ux = rng_cnn_u16(blk0, (uint32_t)j & 3u, 0u);
                    uy = rng_cnn_u16(blk0, (uint32_t)j & 3u, 1u);
                }
                if (ux >= drop_p) r.x *= inv_keep; else { r.x = 0.f; cx = 255u; }
                if (uy >= drop_p) r.y *= inv_keep; else { r.y = 0.f; cy = 255u; }
            }
            st2(dst + (size_t)j * ld, r);
            if (amax) *reinterpret_cast<uint16_t*>(adst + (size_t)j * C) = (uint16_t)(cx | (cy << 8));
        }
        w0 = w1; w1 = w2; w2 = w3; w3 = m;
        sx >>= 1; sy >>= 1;
    };

    int i = 0;
    for (; i + 4 <= n_pairs; i += 4) {
        float2 u[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) u[q] = ld2(src + (size_t)(2 * i + q) * ld);
#pragma unroll
        for (int q = 0; q < 4; ++q) step(i + q, u[2 * q], u[2 * q + 1]);
    }
    for (; i < n_pairs; ++i) step(i, ld2(src + (size_t)(2 * i) * ld), ld2(src + (size_t)(2 * i + 1) * ld));
}

// ---------------------------------------------------------------------------------------------
// K2 forward, packed (bf16, C % 8 == 0, ld == C): thread = (sample, 8 channels), 16-byte loads, and the pooling done on the
// STORED bf16 values instead of on fp32 BatchNorm outputs.
//
// BatchNorm's affine map z = y * sc + sh is monotone in y (increasing for sc > 0, decreasing for sc < 0), so
//     max_i relu(z_i) = relu(fma(max_i y'_i, |sc|, sh)),    y' = y with its sign flipped where sc < 0
// (fma(-y, -sc, sh) == fma(y, sc, sh) exactly, and rounding is monotone, so the result is bit-identical to applying BatchNorm
// first).  The sign flip is one XOR per channel PAIR, every max / equality test is one packed bf16x2 instruction per channel pair,
// and only the pooled value -- one element in two -- goes through the fp32 fma, ReLU, Dropout and the bf16 rounding.  The
// arg-max code (offset 0..9 of the window's FIRST maximum, 255 = no gradient) comes from packed equality masks: the window is a
// ring of 4 previous pair maxima + the current one, each with the mask "second position of the pair is larger".
// Ties are decided on y' (the reference's fp64 BatchNorm output ties exactly when y ties), where the streaming kernel above
// compared fp32 z.  About half the instructions per element of that kernel, which was bound by instruction issue at 2.3 TB/s.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hmax2_u(uint32_t a, uint32_t b) {
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t heq2_m(uint32_t a, uint32_t b) {
    return __heq2_mask(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
}
__device__ __forceinline__ uint32_t hgt2_m(uint32_t a, uint32_t b) {
    return __hgt2_mask(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
}

template <int DROP>
__global__ void __launch_bounds__(256, 2)
k2_fwd_packed_kernel(const bf16* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift, bf16* __restrict__ a,
                     int B, int Lc, int Lp, int C, float drop_p, const float* __restrict__ drop_u, RngState const* rng, uint32_t rng_stream,
                     int64_t row_offset, uint8_t* __restrict__ amax) {
    const int groups = C >> 3;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)B * groups) return;
    const int g = (int)(t % groups), b = (int)(t / groups);
    const int c0 = 8 * g;
    float asc[8], sh[8];
    uint32_t fm[4];
    {
        const float4 s0 = *reinterpret_cast<const float4*>(scale + c0), s1 = *reinterpret_cast<const float4*>(scale + c0 + 4);
        const float4 h0 = *reinterpret_cast<const float4*>(shift + c0), h1 = *reinterpret_cast<const float4*>(shift + c0 + 4);
        const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        const float hh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) { asc[i] = fabsf(sc[i]); sh[i] = hh[i]; }
#pragma unroll
        for (int w = 0; w < 4; ++w) fm[w] = (sc[2 * w] < 0.f ? 0x00008000u : 0u) | (sc[2 * w + 1] < 0.f ? 0x80000000u : 0u);
    }
    const uint4* src = reinterpret_cast<const uint4*>(y + (size_t)b * Lc * C + c0);
    const int rs4 = C >> 3;                               // uint4 per row
    bf16* dst = a + (size_t)b * Lp * C + c0;
    uint8_t* adst = amax ? amax + (size_t)b * Lp * C + c0 : nullptr;
    const float inv_keep = DROP ? 1.f / (1.f - drop_p) : 1.f;
    RngState rs = {0, 0};
    if (DROP == 2) rs = *rng;
    uint32_t rv[4][4], rm[4][4];                          // ring [slot][word]: pair maximum (packed y'), "second is larger" mask
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int w = 0; w < 4; ++w) { rv[q][w] = 0u; rm[q][w] = 0u; }
    const int n_pairs = Lp + 4;
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    for (int i0 = 0; i0 < n_pairs; i0 += 4) {
        uint4 u[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int row = 2 * i0 + q;
            u[q] = row < Lc ? __ldg(src + (size_t)row * rs4) : zero4;
        }
        uint4 blk[4];                                     // Philox blocks of the four channel pairs for outputs j0 .. j0 + 3
        const int j0 = i0 - 4;
        if (DROP == 2 && j0 >= 0) {
#pragma unroll
            for (int w = 0; w < 4; ++w) blk[w] = rng_cnn_block(rs, rng_stream, (uint64_t)(row_offset + b), C, (c0 >> 1) + w, Lp, j0 >> 2);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {                     // pair i = i0 + q lives in ring slot q afterwards; slots q, q+1, .. hold pairs i-4 .. i-1
            const uint32_t p0[4] = {u[2 * q].x ^ fm[0], u[2 * q].y ^ fm[1], u[2 * q].z ^ fm[2], u[2 * q].w ^ fm[3]};
            const uint32_t p1[4] = {u[2 * q + 1].x ^ fm[0], u[2 * q + 1].y ^ fm[1], u[2 * q + 1].z ^ fm[2], u[2 * q + 1].w ^ fm[3]};
            uint32_t mv[4], mm[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) { mv[w] = hmax2_u(p0[w], p1[w]); mm[w] = hgt2_m(p1[w], p0[w]); }
            const int j = i0 + q - 4;
            if (j >= 0 && j < Lp) {
                uint32_t aw[4], cw[2] = {0u, 0u};
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const uint32_t v0 = rv[q][w], v1 = rv[(q + 1) & 3][w], v2 = rv[(q + 2) & 3][w], v3 = rv[(q + 3) & 3][w];
                    const uint32_t r = hmax2_u(hmax2_u(hmax2_u(v0, v1), hmax2_u(v2, v3)), mv[w]);
                    // packed codes: 2 * slot + (second position larger), first slot (oldest pair) that attains the maximum wins
                    uint32_t code = 0x00080008u | (mm[w] & 0x00010001u);
                    uint32_t e = heq2_m(v3, r);
                    code = (e & (0x00060006u | (rm[(q + 3) & 3][w] & 0x00010001u))) | (~e & code);
                    e = heq2_m(v2, r);
                    code = (e & (0x00040004u | (rm[(q + 2) & 3][w] & 0x00010001u))) | (~e & code);
                    e = heq2_m(v1, r);
                    code = (e & (0x00020002u | (rm[(q + 1) & 3][w] & 0x00010001u))) | (~e & code);
                    e = heq2_m(v0, r);
                    code = (e & (rm[q][w] & 0x00010001u)) | (~e & code);
                    // the pooled value: BatchNorm on the maximum, ReLU, Dropout
                    float zx = fmaf(__uint_as_float(r << 16), asc[2 * w], sh[2 * w]);
                    float zy = fmaf(__uint_as_float(r & 0xFFFF0000u), asc[2 * w + 1], sh[2 * w + 1]);
                    bool kx = zx > 0.f, ky = zy > 0.f;
                    if (DROP) {
                        float ux, uy;
                        if (DROP == 1) {
                            ux = drop_u[((size_t)b * C + c0 + 2 * w) * Lp + j];
                            uy = drop_u[((size_t)b * C + c0 + 2 * w + 1) * Lp + j];
                        } else {
                            ux = rng_cnn_u16(blk[w], (uint32_t)q, 0u);        // j & 3 == q: j0 is a multiple of 4
                            uy = rng_cnn_u16(blk[w], (uint32_t)q, 1u);
                        }
                        kx = kx && ux >= drop_p;
                        ky = ky && uy >= drop_p;
                    }
                    zx = kx ? zx * inv_keep : 0.f;
                    zy = ky ? zy * inv_keep : 0.f;
                    __nv_bfloat162 o2 = __floats2bfloat162_rn(zx, zy);
                    aw[w] = *reinterpret_cast<uint32_t*>(&o2);
                    const uint32_t cx = kx ? (code & 0xFFu) : 255u, cy = ky ? ((code >> 16) & 0xFFu) : 255u;
                    cw[w >> 1] |= (cx | (cy << 8)) << (16 * (w & 1));
                }
                *reinterpret_cast<uint4*>(dst + (size_t)j * C) = make_uint4(aw[0], aw[1], aw[2], aw[3]);
                if (adst) *reinterpret_cast<uint2*>(adst + (size_t)j * C) = make_uint2(cw[0], cw[1]);
            }
#pragma unroll
            for (int w = 0; w < 4; ++w) { rv[q][w] = mv[w]; rm[q][w] = mm[w]; }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K2 backward stage 1, streaming: thread = (sample, channel pair).  A 10-position ring of y and of the
// accumulating dz lives in registers; when pair i arrives window j = i-4 is complete, its gradient goes to the
// FIRST maximum (strict '>'), and positions 2j, 2j+1 can no longer change, so they are emitted.  y, a and d(a)
// are each read once, dz written once.  blockDim = (32 channel pairs, 8 samples); per-channel BatchNorm
// reductions are folded over the 8 samples in shared memory, then one fp64 atomic per channel and CTA.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
pool_bn_bwd_stream_kernel(const T* __restrict__ y, const T* __restrict__ a, const T* __restrict__ ga,
                          const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                          const float* __restrict__ rstd, T* __restrict__ dz_out, double* __restrict__ bstats, int B, int Lc, int Lp,
                          int C, int ld, float drop_p) {
    __shared__ float red[4][8][33];
    const int pairs = C / 2;
    const int cp = blockIdx.x * 32 + threadIdx.x;
    const int b = blockIdx.y * 8 + threadIdx.y;
    const bool ok = cp < pairs && b < B;
    float sdz[2] = {0.f, 0.f}, sdzx[2] = {0.f, 0.f};
    if (ok) {
        const int c = 2 * cp;
        const float scx = scale[c], scy = scale[c + 1], shx = shift[c], shy = shift[c + 1];
        const float mux = mean[c], muy = mean[c + 1], rsx = rstd[c], rsy = rstd[c + 1];
        const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
        const T* ysrc = y + (size_t)b * Lc * ld + c;
        const T* asrc = a + (size_t)b * Lp * ld + c;
        const T* gsrc = ga + (size_t)b * Lp * ld + c;
        T* dst = dz_out + (size_t)b * Lc * ld + c;
        float2 yr[10], dr[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) { yr[i] = make_float2(0.f, 0.f); dr[i] = make_float2(0.f, 0.f); }
        // ring slot s holds position 2*(i-4) + s after pair i has been loaded
        auto step = [&](int i, float2 y0, float2 y1, float2 av, float2 gv) {
            yr[8] = y0;
            yr[9] = y1;
            dr[8] = make_float2(0.f, 0.f);
            dr[9] = make_float2(0.f, 0.f);
            if (i >= 4) {
                const int j = i - 4;
                if (av.x > 0.f) {      // kept by dropout and the window maximum was positive
                    int best = 0;
                    float bm = -INFINITY;
#pragma unroll
                    for (int s = 0; s < 10; ++s) { float z = fmaf(yr[s].x, scx, shx); if (z > bm) { bm = z; best = s; } }
                    const float g = gv.x * inv_keep;
#pragma unroll
                    for (int s = 0; s < 10; ++s) dr[s].x += (s == best) ? g : 0.f;
                }
                if (av.y > 0.f) {
                    int best = 0;
                    float bm = -INFINITY;
#pragma unroll
                    for (int s = 0; s < 10; ++s) { float z = fmaf(yr[s].y, scy, shy); if (z > bm) { bm = z; best = s; } }
                    const float g = gv.y * inv_keep;
#pragma unroll
                    for (int s = 0; s < 10; ++s) dr[s].y += (s == best) ? g : 0.f;
                }
                // positions 2j and 2j+1 are final
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    st2(dst + (size_t)(2 * j + s) * ld, dr[s]);
                    sdz[0] += dr[s].x; sdz[1] += dr[s].y;
                    sdzx[0] += dr[s].x * (yr[s].x - mux) * rsx;
                    sdzx[1] += dr[s].y * (yr[s].y - muy) * rsy;
                }
            }
#pragma unroll
            for (int s = 0; s < 8; ++s) { yr[s] = yr[s + 2]; dr[s] = dr[s + 2]; }
        };
        const int n_pairs = Lp + 4;
        const float2 zero2 = make_float2(0.f, 0.f);
        int i = 0;
        for (; i < 4 && i < n_pairs; ++i) step(i, ld2(ysrc + (size_t)(2 * i) * ld), ld2(ysrc + (size_t)(2 * i + 1) * ld), zero2, zero2);
        for (; i + 2 <= n_pairs; i += 2) {
            // all eight loads of two pairs are issued before the first is consumed
            const float2 ya = ld2(ysrc + (size_t)(2 * i) * ld), yb = ld2(ysrc + (size_t)(2 * i + 1) * ld);
            const float2 yc = ld2(ysrc + (size_t)(2 * i + 2) * ld), yd = ld2(ysrc + (size_t)(2 * i + 3) * ld);
            const float2 a0 = ld2(asrc + (size_t)(i - 4) * ld), g0 = ld2(gsrc + (size_t)(i - 4) * ld);
            const float2 a1 = ld2(asrc + (size_t)(i - 3) * ld), g1 = ld2(gsrc + (size_t)(i - 3) * ld);
            step(i, ya, yb, a0, g0);
            step(i + 1, yc, yd, a1, g1);
        }
        for (; i < n_pairs; ++i)
            step(i, ld2(ysrc + (size_t)(2 * i) * ld), ld2(ysrc + (size_t)(2 * i + 1) * ld), ld2(asrc + (size_t)(i - 4) * ld), ld2(gsrc + (size_t)(i - 4) * ld));
        // after the loop slots 0..7 hold positions 2*Lp .. 2*Lp+7; everything beyond was never inside a window
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const int l = 2 * Lp + s;
            if (l < Lc) {
                st2(dst + (size_t)l * ld, dr[s]);
                sdz[0] += dr[s].x; sdz[1] += dr[s].y;
                sdzx[0] += dr[s].x * (yr[s].x - mux) * rsx;
                sdzx[1] += dr[s].y * (yr[s].y - muy) * rsy;
            }
        }
        for (int l = 2 * Lp + 8; l < Lc; ++l) st2(dst + (size_t)l * ld, make_float2(0.f, 0.f));
    }
    red[0][threadIdx.y][threadIdx.x] = sdz[0];
    red[1][threadIdx.y][threadIdx.x] = sdz[1];
    red[2][threadIdx.y][threadIdx.x] = sdzx[0];
    red[3][threadIdx.y][threadIdx.x] = sdzx[1];
    __syncthreads();
    if (threadIdx.y < 4 && cp < pairs) {
        double tot = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot += red[threadIdx.y][i][threadIdx.x];
        const int c = 2 * cp + (threadIdx.y & 1);
        atomicAdd(&bstats[(threadIdx.y < 2 ? 0 : C) + c], tot);
    }
}

// K2 backward stage 2, two channels per thread
template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_apply_v2_kernel(const T* __restrict__ y, T* __restrict__ dz, const double* __restrict__ bstats, const float* __restrict__ gamma,
                       const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ dbias, int64_t R, int C,
                       int ld, double n) {
    __shared__ float s1[2][8][33];
    const int pairs = C / 2;
    const int cp = blockIdx.x * 32 + threadIdx.x;
    float2 acc = make_float2(0.f, 0.f);
    if (cp < pairs) {
        const int c = 2 * cp;
        const float2 dbn = make_float2((float)(bstats[c] / n), (float)(bstats[c + 1] / n));
        const float2 dgn = make_float2((float)(bstats[C + c] / n), (float)(bstats[C + c + 1] / n));
        const float2 mu = make_float2(mean[c], mean[c + 1]), rs = make_float2(rstd[c], rstd[c + 1]);
        const float2 gs = make_float2(gamma[c] * rs.x, gamma[c + 1] * rs.y);
#pragma unroll 4
        for (int64_t r = (int64_t)blockIdx.y * 8 + threadIdx.y; r < R; r += (int64_t)gridDim.y * 8) {
            float2 yv = ld2(y + r * ld + c), dv = ld2(dz + r * ld + c);
            float2 v;
            v.x = gs.x * (dv.x - dbn.x - (yv.x - mu.x) * rs.x * dgn.x);
            v.y = gs.y * (dv.y - dbn.y - (yv.y - mu.y) * rs.y * dgn.y);
            st2(dz + r * ld + c, v);
            acc.x += v.x;
            acc.y += v.y;
        }
    }
    s1[0][threadIdx.y][threadIdx.x] = acc.x;
    s1[1][threadIdx.y][threadIdx.x] = acc.y;
    __syncthreads();
    if (dbias && threadIdx.y < 2 && cp < pairs) {
        float t = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += s1[threadIdx.y][i][threadIdx.x];
        atomicAdd(&dbias[2 * cp + threadIdx.y], t);
    }
}

// BatchNorm batch statistics, two channels per thread
template <typename T>
__global__ void __launch_bounds__(256)
bn_stats_v2_kernel(const T* __restrict__ y, double* __restrict__ stats, int64_t R, int C, int ld) {
    __shared__ float s[4][8][33];
    const int pairs = C / 2;
    const int cp = blockIdx.x * 32 + threadIdx.x;
    float2 a = make_float2(0.f, 0.f), q = make_float2(0.f, 0.f);
    if (cp < pairs) {
#pragma unroll 4
        for (int64_t r = (int64_t)blockIdx.y * 8 + threadIdx.y; r < R; r += (int64_t)gridDim.y * 8) {
            float2 v = ld2(y + r * ld + 2 * cp);
            a.x += v.x; a.y += v.y;
            q.x += v.x * v.x; q.y += v.y * v.y;
        }
    }
    s[0][threadIdx.y][threadIdx.x] = a.x;
    s[1][threadIdx.y][threadIdx.x] = a.y;
    s[2][threadIdx.y][threadIdx.x] = q.x;
    s[3][threadIdx.y][threadIdx.x] = q.y;
    __syncthreads();
    if (threadIdx.y < 4 && cp < pairs) {
        double tot = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot += s[threadIdx.y][i][threadIdx.x];
        atomicAdd(&stats[(threadIdx.y < 2 ? 0 : C) + 2 * cp + (threadIdx.y & 1)], tot);
    }
}

// ---- eight-channel (16-byte) helpers ------------------------------------------------------------------------
__device__ __forceinline__ void ld8(const float* p, float* v) {
    float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const bf16* p, float* v) {
    uint4 q = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
}

// K2 backward stage 2, eight channels (one 16-byte vector) per thread: block = (C/8 channel groups, rows)
template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_apply_v8_kernel(const T* __restrict__ y, T* __restrict__ dz, const double* __restrict__ bstats, const float* __restrict__ gamma,
                       const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ dbias, int64_t R, int C,
                       int ld, double n) {
    extern __shared__ float sred[];                 // [blockDim.y][C]
    const int g = threadIdx.x, c0 = 8 * g;
    float dbn[8], dgn[8], mu[8], rs[8], gs[8], acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        dbn[i] = (float)(bstats[c0 + i] / n);
        dgn[i] = (float)(bstats[C + c0 + i] / n);
        mu[i] = mean[c0 + i];
        rs[i] = rstd[c0 + i];
        gs[i] = gamma[c0 + i] * rs[i];
        acc[i] = 0.f;
    }
    for (int64_t r = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; r < R; r += (int64_t)gridDim.x * blockDim.y) {
        float yv[8], dv[8];
        ld8(y + r * ld + c0, yv);
        ld8(dz + r * ld + c0, dv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            dv[i] = gs[i] * (dv[i] - dbn[i] - (yv[i] - mu[i]) * rs[i] * dgn[i]);
            acc[i] += dv[i];
        }
        st8(dz + r * ld + c0, dv);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) sred[threadIdx.y * C + c0 + i] = acc[i];
    __syncthreads();
    for (int c = threadIdx.y * blockDim.x + threadIdx.x; dbias && c < C; c += blockDim.x * blockDim.y) {
        float t = 0.f;
        for (int ry = 0; ry < (int)blockDim.y; ++ry) t += sred[ry * C + c];
        atomicAdd(&dbias[c], t);
    }
}

// ---------------------------------------------------------------------------------------------
// K1 forward with TAP-PAIR tables: two adjacent taps are looked up at once (5x5 base combinations incl. "outside"),
// which halves the shared-memory reads and adds of the gather-sum.  tab2[(pair*25 + b0*5 + b1)*C1 + o].
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
onehot_conv_fwd_pair_kernel(const uint8_t* __restrict__ bases, const float* __restrict__ w, const float* __restrict__ bias,
                            T* __restrict__ y, double* __restrict__ stats, int B, int C1, int k, int ld) {
    extern __shared__ float smem[];
    const int n_pairs = (k + 1) / 2;               // the last pair of an odd k has an empty second tap
    float* tab = smem;                             // [n_pairs][25][C1]
    float* sb = tab + n_pairs * 25 * C1;           // [C1]
    float* red = sb + C1;                          // [256][16]
    uint8_t* sbase = (uint8_t*)(red + 256 * 16);   // [256 + 2p + 1]
    const int p = (k - 1) / 2;
    const int groups = C1 / 8;
    for (int i = threadIdx.x; i < n_pairs * 25 * C1; i += blockDim.x) {
        const int o = i % C1, combo = (i / C1) % 25, pr = i / (25 * C1);
        const int b0 = combo / 5, b1 = combo % 5, t0 = 2 * pr, t1 = 2 * pr + 1;
        float v = 0.f;
        if (b0 < 4) v += w[((size_t)o * 4 + b0) * k + t0];
        if (b1 < 4 && t1 < k) v += w[((size_t)o * 4 + b1) * k + t1];
        tab[i] = v;
    }
    for (int i = threadIdx.x; i < C1; i += blockDim.x) sb[i] = bias[i];
    const int og = threadIdx.x % groups;
    float s1[8], s2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < SEQ_LEN + 2 * p + 1; i += blockDim.x) {
            int l = i - p;
            sbase[i] = (l >= 0 && l < SEQ_LEN) ? bases[(size_t)b * SEQ_LEN + l] : 4;
        }
        __syncthreads();
        for (int item = threadIdx.x; item < SEQ_LEN * groups; item += 256) {
            const int l = item / groups;
            float acc[8];
            const float4* b4 = reinterpret_cast<const float4*>(sb + og * 8);
            float4 t0 = b4[0], t1 = b4[1];
            acc[0] = t0.x; acc[1] = t0.y; acc[2] = t0.z; acc[3] = t0.w; acc[4] = t1.x; acc[5] = t1.y; acc[6] = t1.z; acc[7] = t1.w;
            for (int pr = 0; pr < n_pairs; ++pr) {
                const int combo = sbase[l + 2 * pr] * 5 + sbase[l + 2 * pr + 1];
                const float4* r4 = reinterpret_cast<const float4*>(tab + (pr * 25 + combo) * C1 + og * 8);
                float4 a0 = r4[0], a1 = r4[1];
                acc[0] += a0.x; acc[1] += a0.y; acc[2] += a0.z; acc[3] += a0.w;
                acc[4] += a1.x; acc[5] += a1.y; acc[6] += a1.z; acc[7] += a1.w;
            }
            st8(y + ((size_t)b * SEQ_LEN + l) * ld + og * 8, acc);
            if (stats) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { float v = round_like<T>(acc[i]); s1[i] += v; s2[i] += v * v; }
            }
        }
    }
    if (stats) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; ++i) { red[threadIdx.x * 16 + i] = s1[i]; red[threadIdx.x * 16 + 8 + i] = s2[i]; }
        __syncthreads();
        for (int i = threadIdx.x; i < groups * 16; i += blockDim.x) {
            int g = i / 16, j = i - g * 16;
            double tot = 0;
            for (int t = g; t < 256; t += groups) tot += red[t * 16 + j];
            int c = g * 8 + (j & 7);
            atomicAdd(&stats[(j < 8 ? 0 : C1) + c], tot);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1 backward with per-base position lists: the 256 positions of a sample are bucketed by base once (counting
// sort in shared memory); thread (channel pair, tap) then walks the four lists with no branch and no bounds check
// (the staged dy tile carries p zero rows on both sides).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024)
onehot_conv_bwd_lists_kernel(const uint8_t* __restrict__ bases, const T* __restrict__ dy, float* __restrict__ dw, int B, int C1, int k, int ld) {
    extern __shared__ float smem[];
    const int p = (k - 1) / 2;
    float* dys = smem;                                           // [256 + 2p][C1] fp32, rows shifted by p
    uint8_t* plist = (uint8_t*)(dys + (SEQ_LEN + 2 * p) * C1);    // [256] positions grouped by base
    __shared__ int cnt[4], start[5];
    const int pairs = C1 / 2;
    const int tap = threadIdx.x / pairs, op = threadIdx.x - tap * pairs;
    const bool active = tap < k;
    double d[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) d[i][c] = 0.0;
    for (int i = threadIdx.x; i < p * C1; i += blockDim.x) { dys[i] = 0.f; dys[(SEQ_LEN + p) * C1 + i] = 0.f; }
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < 4) cnt[threadIdx.x] = 0;
        for (int i = threadIdx.x; i < SEQ_LEN * pairs; i += blockDim.x) {
            int l = i / pairs, o2 = i - l * pairs;
            float2 v = ld2(dy + ((size_t)b * SEQ_LEN + l) * ld + 2 * o2);
            *reinterpret_cast<float2*>(dys + (l + p) * C1 + 2 * o2) = v;
        }
        __syncthreads();
        // counting sort of the positions by base (warp 0: ballot-based, keeps ascending position order per base)
        if (threadIdx.x < 32) {
            int base_cnt[4] = {0, 0, 0, 0};
            uint8_t mine[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                mine[q] = bases[(size_t)b * SEQ_LEN + q * 32 + threadIdx.x];
#pragma unroll
                for (int c = 0; c < 4; ++c) base_cnt[c] += __popc(__ballot_sync(0xffffffffu, mine[q] == c));
            }
            int st[5];
            st[0] = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) st[c + 1] = st[c] + base_cnt[c];
            if (threadIdx.x < 5) start[threadIdx.x] = st[threadIdx.x];
            int run[4] = {st[0], st[1], st[2], st[3]};
#pragma unroll
            for (int q = 0; q < 8; ++q) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    unsigned m = __ballot_sync(0xffffffffu, mine[q] == c);
                    if (mine[q] == c) plist[run[c] + __popc(m & ((1u << threadIdx.x) - 1))] = (uint8_t)(q * 32 + threadIdx.x);
                    run[c] += __popc(m);
                }
            }
        }
        __syncthreads();
        if (active) {
            // source position ls feeds output l = ls - tap + p, stored at row l + p = ls - tap + 2p of dys
            const float* col = dys + (2 * p - tap) * C1 + 2 * op;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float a0 = 0.f, a1 = 0.f;
                const int e = start[c + 1];
                for (int i = start[c]; i < e; ++i) {
                    const float2 v = *reinterpret_cast<const float2*>(col + (int)plist[i] * C1);
                    a0 += v.x;
                    a1 += v.y;
                }
                d[0][c] += a0;
                d[1][c] += a1;
            }
        }
    }
    if (active) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) atomicAdd(&dw[((size_t)(2 * op + i) * 4 + c) * k + tap], (float)d[i][c]);
    }
}

// bf16 variant of the per-base-list K1 backward: the staged dy tile stays in bf16 (half the shared-memory bytes of the
// fp32 tile: the kernel is bound by shared-memory reads), 16-byte coalesced global loads.  Requires C1 % 8 == 0.
__global__ void __launch_bounds__(1024)
onehot_conv_bwd_lists16_kernel(const uint8_t* __restrict__ bases, const bf16* __restrict__ dy, float* __restrict__ dw, int B, int C1, int k, int ld) {
    extern __shared__ float smem[];
    const int p = (k - 1) / 2;
    bf16* dys = (bf16*)smem;                                      // [256 + 2p][C1] bf16, rows shifted by p
    uint8_t* plist = (uint8_t*)(dys + (SEQ_LEN + 2 * p) * C1);    // [256] positions grouped by base
    __shared__ int start[5];
    const int pairs = C1 / 2;
    const int tap = threadIdx.x / pairs, op = threadIdx.x - tap * pairs;
    const bool active = tap < k;
    double d[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) d[i][c] = 0.0;
    for (int i = threadIdx.x; i < p * C1 / 2; i += blockDim.x) {
        ((uint32_t*)dys)[i] = 0u;
        ((uint32_t*)(dys + (SEQ_LEN + p) * C1))[i] = 0u;
    }
    const int vec_per_row = C1 / 8;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < SEQ_LEN * vec_per_row; i += blockDim.x) {
            const int l = i / vec_per_row, v = i - l * vec_per_row;
            *reinterpret_cast<uint4*>(dys + (l + p) * C1 + 8 * v) = *reinterpret_cast<const uint4*>(dy + ((size_t)b * SEQ_LEN + l) * ld + 8 * v);
        }
        // counting sort of the positions by base (warp 0: ballot-based, keeps ascending position order per base)
        if (threadIdx.x < 32) {
            int base_cnt[4] = {0, 0, 0, 0};
            uint8_t mine[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                mine[q] = bases[(size_t)b * SEQ_LEN + q * 32 + threadIdx.x];
#pragma unroll
                for (int c = 0; c < 4; ++c) base_cnt[c] += __popc(__ballot_sync(0xffffffffu, mine[q] == c));
            }
            int st[5];
            st[0] = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) st[c + 1] = st[c] + base_cnt[c];
            if (threadIdx.x < 5) start[threadIdx.x] = st[threadIdx.x];
            int run[4] = {st[0], st[1], st[2], st[3]};
#pragma unroll
            for (int q = 0; q < 8; ++q) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    unsigned m = __ballot_sync(0xffffffffu, mine[q] == c);
                    if (mine[q] == c) plist[run[c] + __popc(m & ((1u << threadIdx.x) - 1))] = (uint8_t)(q * 32 + threadIdx.x);
                    run[c] += __popc(m);
                }
            }
        }
        __syncthreads();
        if (active) {
            // source position ls feeds output l = ls - tap + p, stored at row l + p = ls - tap + 2p of dys
            const bf16* col = dys + (2 * p - tap) * C1 + 2 * op;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
                const int e = start[c + 1];
                int i = start[c];
                for (; i + 1 < e; i += 2) {                     // two independent chains
                    const uint32_t w0 = *reinterpret_cast<const uint32_t*>(col + (int)plist[i] * C1);
                    const uint32_t w1 = *reinterpret_cast<const uint32_t*>(col + (int)plist[i + 1] * C1);
                    a0 += __uint_as_float(w0 << 16); a1 += __uint_as_float(w0 & 0xFFFF0000u);
                    b0 += __uint_as_float(w1 << 16); b1 += __uint_as_float(w1 & 0xFFFF0000u);
                }
                if (i < e) {
                    const uint32_t w0 = *reinterpret_cast<const uint32_t*>(col + (int)plist[i] * C1);
                    a0 += __uint_as_float(w0 << 16); a1 += __uint_as_float(w0 & 0xFFFF0000u);
                }
                d[0][c] += a0 + b0;
                d[1][c] += a1 + b1;
            }
        }
    }
    if (active) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) atomicAdd(&dw[((size_t)(2 * op + i) * 4 + c) * k + tap], (float)d[i][c]);
    }
}

// ---------------------------------------------------------------------------------------------
// K1 forward with TAP-TRIPLE tables: three adjacent taps are looked up at once (5^3 = 125 base combinations incl.
// "outside"), so a k = 15 convolution is 5 table rows per output instead of 15 (or 8 with tap pairs): the kernel is bound by
// shared-memory reads.  tab3[(triple * 125 + b0 * 25 + b1 * 5 + b2) * C1 + o]: 160 KB for C1 = 64, k = 15 -> one persistent
// 1024-thread CTA per SM.  Requires C1 % 8 == 0, 1024 % (C1 / 8) == 0.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024, 1)
onehot_conv_fwd_triple_kernel(const uint8_t* __restrict__ bases, const float* __restrict__ w, const float* __restrict__ bias,
                              T* __restrict__ y, double* __restrict__ stats, int B, int C1, int k, int ld) {
    extern __shared__ float smem[];
    const int n_tr = (k + 2) / 3;                  // the last triple of k = 5 / 11 has empty taps
    float* tab = smem;                             // [n_tr][125][C1]
    float* sb = tab + n_tr * 125 * C1;             // [C1]
    float* red = sb + C1;                          // [32 warps][groups][16]
    const int p = (k - 1) / 2;
    const int groups = C1 / 8;
    uint8_t* sbase = (uint8_t*)(red + 32 * groups * 16);   // [256 + 2p + 2]
    for (int i = threadIdx.x; i < n_tr * 125 * C1; i += blockDim.x) {
        const int o = i % C1, combo = (i / C1) % 125, tr = i / (125 * C1);
        const int b0 = combo / 25, b1 = (combo / 5) % 5, b2 = combo % 5, t0 = 3 * tr;
        float v = 0.f;
        if (b0 < 4) v += w[((size_t)o * 4 + b0) * k + t0];
        if (b1 < 4 && t0 + 1 < k) v += w[((size_t)o * 4 + b1) * k + t0 + 1];
        if (b2 < 4 && t0 + 2 < k) v += w[((size_t)o * 4 + b2) * k + t0 + 2];
        tab[i] = v;
    }
    for (int i = threadIdx.x; i < C1; i += blockDim.x) sb[i] = bias[i];
    const int og = threadIdx.x % groups;           // constant per thread: 1024 % groups == 0
    float s1[8], s2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < SEQ_LEN + 2 * p + 2; i += blockDim.x) {
            const int l = i - p;
            sbase[i] = (l >= 0 && l < SEQ_LEN) ? bases[(size_t)b * SEQ_LEN + l] : 4;
        }
        __syncthreads();
        for (int item = threadIdx.x; item < SEQ_LEN * groups; item += blockDim.x) {
            const int l = item / groups;
            float acc[8];
            const float4* b4 = reinterpret_cast<const float4*>(sb + og * 8);
            float4 t0 = b4[0], t1 = b4[1];
            acc[0] = t0.x; acc[1] = t0.y; acc[2] = t0.z; acc[3] = t0.w; acc[4] = t1.x; acc[5] = t1.y; acc[6] = t1.z; acc[7] = t1.w;
            for (int tr = 0; tr < n_tr; ++tr) {
                const int combo = sbase[l + 3 * tr] * 25 + sbase[l + 3 * tr + 1] * 5 + sbase[l + 3 * tr + 2];
                const float4* r4 = reinterpret_cast<const float4*>(tab + (tr * 125 + combo) * C1 + og * 8);
                float4 a0 = r4[0], a1 = r4[1];
                acc[0] += a0.x; acc[1] += a0.y; acc[2] += a0.z; acc[3] += a0.w;
                acc[4] += a1.x; acc[5] += a1.y; acc[6] += a1.z; acc[7] += a1.w;
            }
            st8(y + ((size_t)b * SEQ_LEN + l) * ld + og * 8, acc);
            if (stats) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { float v = round_like<T>(acc[i]); s1[i] += v; s2[i] += v * v; }
            }
        }
    }
    if (stats) {   // lanes with the same og hold partials of the same 8 channels: fold them inside the warp, then across warps
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            for (int off = 16; off >= groups; off >>= 1) {
                s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], off);
                s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], off);
            }
        __syncthreads();
        if (lane < groups) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { red[(warp * groups + lane) * 16 + i] = s1[i]; red[(warp * groups + lane) * 16 + 8 + i] = s2[i]; }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < groups * 16; i += blockDim.x) {
            const int g = i / 16, j = i - g * 16;
            double tot = 0;
            for (int wv = 0; wv < 32; ++wv) tot += red[(wv * groups + g) * 16 + j];
            atomicAdd(&stats[(j < 8 ? 0 : C1) + g * 8 + (j & 7)], tot);
        }
    }
}

}  // namespace emb
