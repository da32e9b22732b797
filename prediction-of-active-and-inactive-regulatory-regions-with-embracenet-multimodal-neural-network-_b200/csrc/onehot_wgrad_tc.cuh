// Weight gradient of the FIRST Conv1d (one-hot input, CNN_pre.py:39 with in_channels = 4) on the tensor cores.
//
//   dW[o, c, t] = sum_{b, l} dy[b, l, o] * [base[b, l + t - pad] == c]                       
//
// The forward of this layer is a table lookup (onehot_conv_fwd_triple_kernel); its weight gradient is a genuine
// contraction over the B * 256 positions, and the one-hot operand never has to exist in global memory:
//
//   A (M x K)  = dy of one sample, [256 positions][64 channels] bf16, fetched by ONE TMA box (128B swizzle) and consumed
//                in place as an MN-major operand (M = channel, K = position), as every other wgrad of the engine does;
//   B (N x K)  = the sample's one-hot rows x_s[r] = onehot(base[r - pad]) (8 bf16 = 16 bytes per row: a, c, g, t, 0, 0, 0, 0),
//                EXPANDED IN SHARED MEMORY from the 256 base codes (one bulk copy of 256 bytes) by the generator warps.
//                Tap t of position k needs row k + t, so the [position][tap * 8 + channel] operand is the Toeplitz view
//                B[k][n] = x_flat[8 k + n] of that one dense array: an un-swizzled MN-major shared-memory descriptor with
//                SBO (stride between 8-wide N blocks) = 16 bytes and LBO (stride between 8-row K blocks) = 128 bytes
//                describes it exactly (the core matrices overlap in memory, which a descriptor is free to do).
//
// So ONE tcgen05.mma (128 x 128 x 16) per 16 positions accumulates ALL taps (16 tap slots x 8 channel slots) at once:
// 16 MMAs per sample, fp32 accumulation in TMEM across every sample of the CTA, one atomic flush of Cout * 4 * k values
// per CTA at the end.  dy crosses HBM exactly once (32 KB per sample per request, 5 stages in flight per SM), which is the
// roofline of this op; the previous per-base position-list kernel was bound by shared-memory reads at ~8x that time.
//
//   warps 0-3  generators (code bytes -> one-hot rows), warps 0-1 also the epilogue (TMEM lanes 0..63 = channels)
//   warp 4     producer   (TMA box of dy + bulk copy of the base codes, mbarrier complete_tx)
//   warp 5     MMA issuer (owns the TMEM allocation)
#pragma once
#include "gemm_tc.cuh"
#include "kernels_tma.cuh"

namespace emb {

constexpr int OHW_STAGES = 5;
constexpr int OHW_THREADS = 192;
constexpr int OHW_A_BYTES = SEQ_LEN * 128;              // [256 positions][64 channels] bf16
constexpr int OHW_B_ROWS = SEQ_LEN + 16;                // rows k + t, k < 256, t < 16
constexpr int OHW_B_BYTES = OHW_B_ROWS * 16;
constexpr int OHW_STAGE = (OHW_A_BYTES + OHW_B_BYTES + SEQ_LEN + 1023) / 1024 * 1024;     // + the 256 base codes; dy tiles stay 1024-byte aligned
constexpr int OHW_SMEM = OHW_STAGES * OHW_STAGE + 1024 + 256;

// un-swizzled (INTERLEAVE) shared-memory matrix descriptor, version 1
__device__ __forceinline__ uint64_t umma_desc_noswizzle(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__global__ void __launch_bounds__(OHW_THREADS, 1)
onehot_conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_dy, const uint8_t* __restrict__ bases, float* __restrict__ dw,
                            int B, int C1, int k, uint32_t idesc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int STAGE = OHW_STAGE;
    uint64_t* full_bar = (uint64_t*)(smem + OHW_STAGES * STAGE);
    uint64_t* empty_bar = full_bar + OHW_STAGES;
    uint64_t* codes_bar = empty_bar + OHW_STAGES;
    uint64_t* acc_bar = codes_bar + OHW_STAGES;
    uint32_t* tmem_slot = (uint32_t*)(acc_bar + 1);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int pad = (k - 1) / 2;
    const int n_mine = blockIdx.x < B ? (B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    // the halo rows (positions outside [0, 256)) and the tap slots beyond them are zero for every sample: clear B once
    for (int s = 0; s < OHW_STAGES; ++s) {
        uint4* b = (uint4*)(smem + s * STAGE + OHW_A_BYTES);
        for (int i = threadIdx.x; i < OHW_B_ROWS; i += OHW_THREADS) b[i] = make_uint4(0, 0, 0, 0);
    }
    fence_async_smem();
    if (threadIdx.x == 0) {
        for (int s = 0; s < OHW_STAGES; ++s) { mbar_init(&full_bar[s], 1 + 4); mbar_init(&empty_bar[s], 1); mbar_init(&codes_bar[s], 1); }
        mbar_init(acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ================= producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < n_mine; ++i) {
                const int b = blockIdx.x + i * gridDim.x;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* st = smem + stage * STAGE;
                mbar_expect_tx(&codes_bar[stage], SEQ_LEN);
                bulk_load(st + OHW_A_BYTES + OHW_B_BYTES, bases + (size_t)b * SEQ_LEN, SEQ_LEN, &codes_bar[stage]);
                mbar_expect_tx(&full_bar[stage], OHW_A_BYTES);
                tma_load_3d(st, &map_dy, &full_bar[stage], 0, 0, b);
                if (++stage == OHW_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 5) {
        // ================= MMA issuer =================
        // A: MN-major, 128B swizzle; the second 64-row block of the 128-row instruction re-reads the first (LBO = 0) and its
        // accumulator rows 64..127 are never looked at.  B: the Toeplitz view described at the top.
        const uint64_t da_base = umma_desc(0, 0, 1024), db_base = umma_desc_noswizzle(0, 128, 16);
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < n_mine; ++i) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * STAGE), sb = sa + OHW_A_BYTES;
            if (elect_one_sync()) {
                const uint64_t da0 = da_base | (uint64_t)((sa & 0x3FFFFu) >> 4), db0 = db_base | (uint64_t)((sb & 0x3FFFFu) >> 4);
#pragma unroll 4
                for (int s2 = 0; s2 < SEQ_LEN / 16; ++s2)          // 16 positions per step: 16 rows of 128 B (A), 16 rows of 16 B (B)
                    tc_mma_f16(tmem_base, da0 + (uint64_t)(s2 * 128), db0 + (uint64_t)(s2 * 16), idesc, (i | s2) ? 1u : 0u);
                tc_commit(&empty_bar[stage]);
            }
            __syncwarp();
            if (++stage == OHW_STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one_sync()) tc_commit(acc_bar);
        __syncwarp();
    } else {
        // ================= generators: 256 code bytes -> 256 one-hot rows of 16 bytes =================
        int stage = 0;
        uint32_t phase = 0;
        const int t = threadIdx.x;                        // 0..127: rows pad + t and pad + t + 128
        for (int i = 0; i < n_mine; ++i) {
            mbar_wait(&codes_bar[stage], phase);
            uint8_t* st = smem + stage * STAGE;
            const uint8_t* codes = st + OHW_A_BYTES + OHW_B_BYTES;
            uint4* rows = (uint4*)(st + OHW_A_BYTES) + pad;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t c = codes[t + h * 128];
                const uint32_t one = (c & 1) ? 0x3F800000u : 0x00003F80u;          // bf16 1.0 in the odd / even half of a word
                rows[t + h * 128] = make_uint4(c < 2 ? one : 0u, (c & ~1u) == 2 ? one : 0u, 0u, 0u);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full_bar[stage])) : "memory");
            if (++stage == OHW_STAGES) { stage = 0; phase ^= 1; }
        }
        // ================= epilogue: warps 0-1 own TMEM lanes 0..63 = output channels =================
        if (warp < 2 && n_mine > 0) {
            mbar_wait(acc_bar, 0);
            tc_fence_after();
            const int o = warp * 32 + lane;
            for (int tp = 0; tp < k; tp += 2) {           // 16 accumulator columns = two tap slots of 8 channels
                float v[16];
                tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(tp * 8), v);
                if (o < C1) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        atomicAdd(&dw[(o * 4 + c) * k + tp], v[c]);
                        if (tp + 1 < k) atomicAdd(&dw[(o * 4 + c) * k + tp + 1], v[8 + c]);
                    }
                }
            }
            tc_fence_before();
        }
    }
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
    }
}

inline bool onehot_wgrad_tc_ok(const void* dy, const uint8_t* bases, int C1, int k, int ld) {
    return C1 >= 8 && C1 <= 64 && (C1 % 8) == 0 && ld == C1 && k >= 1 && k <= 15 && (k & 1) && !((uintptr_t)dy & 15) && !((uintptr_t)bases & 15) &&
           !tuning().no_onehot_wgrad_tc;
}

// dw[C1][4][k] += ... (fp32 atomics; the caller zeroes it).  dy: [B, 256, ld] bf16.
inline int onehot_conv_wgrad_tc(const uint8_t* bases, const bf16* dy, float* dw, int B, int C1, int k, int ld, cudaStream_t st, int one_cta = 0) {
    int rc = tc_init();
    if (rc) return rc;
    if (first_on_device(1)) {
        cudaError_t e = cudaFuncSetAttribute(onehot_conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OHW_SMEM);
        if (e != cudaSuccess) return set_error(-3, "cudaFuncSetAttribute(onehot_conv_wgrad_tc_kernel): %s", cudaGetErrorString(e));
    }
    CUtensorMap map;
    rc = make_map(&map, dy, C1, SEQ_LEN, B, ld, (int64_t)SEQ_LEN * ld, 64, SEQ_LEN, 1);
    if (rc) return rc;
    const uint32_t idesc = make_idesc(1, 1, 128);
    const int grid = one_cta ? 1 : std::min(B, tc_num_sms());      // one CTA = one contributor: the deterministic-reduction mode
    onehot_conv_wgrad_tc_kernel<<<grid, OHW_THREADS, OHW_SMEM, st>>>(map, bases, dw, B, C1, k, idesc);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return set_error(-3, "onehot_conv_wgrad_tc launch failed: %s", cudaGetErrorString(err));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Forward of the first Conv1d through the same in-shared-memory one-hot operand.
//
//   y[b, l, o] = bias[o] + sum_{t, c} W[o, c, t] * [base[b, l + t - pad] == c]
//
// A (M x K)  = the Toeplitz view of the one-hot rows, now as the K-major operand: A[m = l][kk = t * 8 + slot] = x_flat[8 l + kk]
//              (un-swizzled K-major descriptor, LBO = 16 bytes between 8-wide K blocks = one row = one tap later, SBO = 128 bytes
//              between 8-row M blocks).  A row holds the one-hot of its base TWICE (slots 0-3 and 4-7).
// B (N x K)  = two matrices W'1[o][t * 8 + slot] = {hi, mid}(W[o, slot & 3, t]) and W'2 = {lo, 0} with hi = bf16(W),
//              mid = bf16(W - hi), lo = W - hi - mid: the fp32 master weight enters as an EXACT three-term bf16 split, every
//              product is exact and the fp32 accumulator therefore holds the gather-sum of the fp32 weights, as the lookup
//              kernels compute it (a two-term split left 0.2 % of the bf16-rounded outputs one ulp off, enough to re-route
//              max-pool gradients downstream); resident in shared memory for the whole kernel (32 KB).
// 16 tcgen05.mma (128 positions x 64 channels x 16; 8 per weight matrix) per half sample, one TMEM accumulator per half; two sets
// of four epilogue warps (one set per half: the kernel is bound by the latency of this epilogue chain, not by the MMAs) add the
// bias, round to bf16, stage their 32 x 64 tile in shared memory in the 128B-swizzle pattern and hand it to a TMA tensor store
// (4 KB per request), and accumulate the layer-0 BatchNorm statistics from the staged (rounded) values.
// Every output byte is written to HBM once, which is the only compulsory traffic of this op (the lookup kernels were bound
// by their 5 shared-memory table reads per output).
//
//   warps 0-3 / 8-11 epilogue of the first / second half sample (TMEM lane quarter = warp % 4), warp 4 producer (base codes),
//   warp 5 MMA issuer, warps 6-7 generators
// ---------------------------------------------------------------------------------------------
constexpr int OHF_STAGES = 4;
constexpr int OHF_THREADS = 384;
constexpr int OHF_W_BYTES = 2 * 64 * 128 * 2;
constexpr int OHF_ROWS_BYTES = OHW_B_ROWS * 16;
constexpr int OHF_STAGE = OHF_ROWS_BYTES + SEQ_LEN;
constexpr int OHF_TILE_BYTES = 32 * 128;
constexpr int OHF_SMEM = 1024 + 8 * OHF_TILE_BYTES + OHF_W_BYTES + OHF_STAGES * OHF_STAGE + 64 * 4 + 8 * 64 * 2 * 4 + 256;

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src_smem, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"((uint64_t)map), "r"(smem_u32(src_smem)), "r"(x), "r"(y), "r"(z) : "memory");
}

__global__ void __launch_bounds__(OHF_THREADS, 2)
onehot_conv_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_y, const uint8_t* __restrict__ bases, const float* __restrict__ w,
                          const float* __restrict__ bias, double* __restrict__ stats, int B, int C1, int k, uint32_t idesc,
                          const float* __restrict__ pool_scale, const float* __restrict__ pool_shift, bf16* __restrict__ pool_out, int pool_Lp) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* tiles = smem;                                          // [8 epilogue warps][32 rows][128 B], 128B-swizzled
    uint8_t* wsm = tiles + 8 * OHF_TILE_BYTES;                      // W'1, W'2 in the un-swizzled K-major core-matrix layout
    uint8_t* stages = wsm + OHF_W_BYTES;                            // [stage]{rows[272][16 B], codes[256]}
    float* sbias = (float*)(stages + OHF_STAGES * OHF_STAGE);       // [64]
    float* sred = sbias + 64;                                       // [8 warps][64 channels][2]; pooled (inference) mode: [64] scale | [64] shift
    uint64_t* codes_bar = (uint64_t*)(sred + 8 * 64 * 2);
    uint64_t* rows_full = codes_bar + OHF_STAGES;
    uint64_t* rows_empty = rows_full + OHF_STAGES;
    uint64_t* tfull = rows_empty + OHF_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int pad = (k - 1) / 2;
    const int n_mine = (int)blockIdx.x < B ? (B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    for (int i = threadIdx.x; i < OHF_STAGES * OHF_STAGE / 16; i += OHF_THREADS) ((uint4*)stages)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < 64 * 64; i += OHF_THREADS) {       // one (channel, tap, base) weight -> its hi, mid and lo slots
        const int o = i >> 6, t = (i >> 2) & 15, c = i & 3;
        float v = (o < C1 && t < k) ? w[((size_t)o * 4 + c) * k + t] : 0.f;
        const bf16 hi = __float2bfloat16_rn(v);
        const float r1 = v - __bfloat162float(hi);
        const bf16 mid = __float2bfloat16_rn(r1), lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
        bf16* dst = (bf16*)(wsm + (o >> 3) * 2048 + t * 128 + (o & 7) * 16);
        dst[c] = hi;
        dst[4 + c] = mid;
        dst[8192 + c] = lo;                                    // second matrix, 16 KB further
        dst[8192 + 4 + c] = __float2bfloat16_rn(0.f);
    }
    for (int i = threadIdx.x; i < 64; i += OHF_THREADS) sbias[i] = i < C1 ? bias[i] : 0.f;
    if (pool_out)
        for (int i = threadIdx.x; i < 64; i += OHF_THREADS) { sred[i] = i < C1 ? pool_scale[i] : 0.f; sred[64 + i] = i < C1 ? pool_shift[i] : 0.f; }
    fence_async_smem();
    if (threadIdx.x == 0) {
        for (int s = 0; s < OHF_STAGES; ++s) { mbar_init(&codes_bar[s], 1); mbar_init(&rows_full[s], 2); mbar_init(&rows_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < n_mine; ++i) {
                const int b = blockIdx.x + i * gridDim.x;
                mbar_wait(&rows_empty[stage], phase ^ 1);
                mbar_expect_tx(&codes_bar[stage], SEQ_LEN);
                bulk_load(stages + stage * OHF_STAGE + OHF_ROWS_BYTES, bases + (size_t)b * SEQ_LEN, SEQ_LEN, &codes_bar[stage]);
                if (++stage == OHF_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 5) {
        // ================= MMA issuer =================
        const uint64_t da_base = umma_desc_noswizzle(0, 16, 128), db0 = umma_desc_noswizzle(smem_u32(wsm), 128, 2048);
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < n_mine; ++i) {
            mbar_wait(&rows_full[stage], phase);
            const uint32_t rows = smem_u32(stages + stage * OHF_STAGE);
            for (int h = 0; h < 2; ++h) {
                mbar_wait(&tempty[h], (uint32_t)(i & 1) ^ 1);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint64_t da0 = da_base | (uint64_t)(((rows + (uint32_t)h * 128u * 16u) & 0x3FFFFu) >> 4);
#pragma unroll
                    for (int s2 = 0; s2 < 16; ++s2)                // 16 K elements = two taps: A + 2 rows (32 B), W' + 2 K blocks (256 B);
                        tc_mma_f16(tmem_base + (uint32_t)(h * 64), da0 + (uint64_t)((s2 & 7) * 2),      // steps 8..15: the same rows against W'2
                                   db0 + (uint64_t)((s2 & 7) * 16 + (s2 >> 3) * 1024), idesc, s2 ? 1u : 0u);
                    tc_commit(&tfull[h]);
                    if (h == 1) tc_commit(&rows_empty[stage]);
                }
                __syncwarp();
            }
            if (++stage == OHF_STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 6 || warp == 7) {
        // ================= generators: 256 code bytes -> 256 rows {onehot, onehot} of 16 bytes =================
        int stage = 0;
        uint32_t phase = 0;
        const int t = threadIdx.x - 6 * 32;               // 0..63
        for (int i = 0; i < n_mine; ++i) {
            mbar_wait(&codes_bar[stage], phase);
            uint8_t* st = stages + stage * OHF_STAGE;
            const uint8_t* codes = st + OHF_ROWS_BYTES;
            uint4* rows = (uint4*)st + pad;
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const uint32_t c = codes[t + h * 64];
                const uint32_t one = (c & 1) ? 0x3F800000u : 0x00003F80u;
                const uint32_t w0 = c < 2 ? one : 0u, w1 = (c & ~1u) == 2 ? one : 0u;
                rows[t + h * 64] = make_uint4(w0, w1, w0, w1);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&rows_full[stage])) : "memory");
            if (++stage == OHF_STAGES) { stage = 0; phase ^= 1; }
        }
    } else {
        // ================= epilogue: thread = position (TMEM lane), 64 channels per thread =================
        float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;             // BatchNorm partials of channels 2*lane, 2*lane + 1
        const int h = warp >> 3, q = warp & 3, ew = h * 4 + q;        // half sample, TMEM lane quarter, epilogue warp index
        uint8_t* tile = tiles + ew * OHF_TILE_BYTES;
        for (int i = 0; i < n_mine; ++i) {
            const int b = blockIdx.x + i * gridDim.x;
            mbar_wait(&tfull[h], (uint32_t)(i & 1));
            tc_fence_after();
            if (!pool_out) {
                if (lane == 0) bulk_wait_read<0>();                   // the previous store has drained the staging tile
                __syncwarp();
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float v[32];
                tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 64 + half * 32), v);
                tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 64 + half * 32 + 16), v + 16);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const int ch = half * 4 + c4;
                    uint32_t pk[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 bb = *reinterpret_cast<const float2*>(sbias + ch * 8 + 2 * j);
                        __nv_bfloat162 t2 = __floats2bfloat162_rn(v[c4 * 8 + 2 * j] + bb.x, v[c4 * 8 + 2 * j + 1] + bb.y);
                        pk[j] = *reinterpret_cast<uint32_t*>(&t2);
                    }
                    *reinterpret_cast<uint4*>(tile + lane * 128 + ((ch ^ (lane & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            }
            tc_fence_before();
            if (pool_out) {
                // Inference: the eight staged tiles together are the sample's whole [256][64] conv output.  Eval BatchNorm + ReLU +
                // MaxPool1d(10, 2) run on it here and only the pooled [Lp][C1] block goes to memory (thread = pooled row x 8 channels).
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tempty[h])) : "memory");
                asm volatile("bar.sync 2, 256;" ::: "memory");             // all eight epilogue warps have staged this sample
                const int nch = C1 >> 3, et = ew * 32 + lane;
                for (int item = et; item < pool_Lp * nch; item += 256) {
                    const int j = item / nch, ch = item - j * nch;
                    float sc[8], sh[8], m[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) { sc[c] = sred[ch * 8 + c]; sh[c] = sred[64 + ch * 8 + c]; m[c] = 0.f; }
#pragma unroll
                    for (int i2 = 0; i2 < 10; ++i2) {
                        const int r = 2 * j + i2, lr = r & 31;
                        const uint4 u = *reinterpret_cast<const uint4*>(tiles + (r >> 5) * OHF_TILE_BYTES + lr * 128 + ((ch ^ (lr & 7)) << 4));
                        const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            m[2 * c] = fmaxf(m[2 * c], fmaf(__uint_as_float(wv[c] << 16), sc[2 * c], sh[2 * c]));
                            m[2 * c + 1] = fmaxf(m[2 * c + 1], fmaf(__uint_as_float(wv[c] & 0xFFFF0000u), sc[2 * c + 1], sh[2 * c + 1]));
                        }
                    }
                    *reinterpret_cast<uint4*>(pool_out + ((size_t)b * pool_Lp + j) * C1 + ch * 8) =
                        make_uint4(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]), pack_bf16x2(m[4], m[5]), pack_bf16x2(m[6], m[7]));
                }
                asm volatile("bar.sync 2, 256;" ::: "memory");             // the tiles may be overwritten by the next sample
                continue;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tempty[h])) : "memory");
                tma_store_3d(&map_y, tile, 0, h * 128 + q * 32, b);
                bulk_commit();
            }
            if (stats) {
#pragma unroll 8
                for (int r = 0; r < 32; ++r) {
                    const uint32_t u = *reinterpret_cast<const uint32_t*>(tile + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4);
                    const float a = __uint_as_float(u << 16), c2 = __uint_as_float(u & 0xFFFF0000u);
                    s1a += a; s2a += a * a; s1b += c2; s2b += c2 * c2;
                }
            }
        }
        if (!pool_out && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");       // every tile is in global memory
        if (stats) {
            sred[(ew * 64 + 2 * lane) * 2] = s1a; sred[(ew * 64 + 2 * lane) * 2 + 1] = s2a;
            sred[(ew * 64 + 2 * lane + 1) * 2] = s1b; sred[(ew * 64 + 2 * lane + 1) * 2 + 1] = s2b;
        }
    }
    __syncthreads();
    if (stats && threadIdx.x < 128) {
        const int ch = threadIdx.x >> 1, which = threadIdx.x & 1;
        if (ch < C1) {
            double tot = 0;
            for (int wv = 0; wv < 8; ++wv) tot += (double)sred[(wv * 64 + ch) * 2 + which];
            atomicAdd(&stats[which * C1 + ch], tot);
        }
    }
    if (warp == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
    }
}

inline bool onehot_fwd_tc_ok(const void* y, const uint8_t* bases, int C1, int k, int ld) {
    return C1 >= 8 && C1 <= 64 && (C1 % 8) == 0 && ld == C1 && k >= 1 && k <= 15 && (k & 1) && !((uintptr_t)y & 15) && !((uintptr_t)bases & 15) &&
           !tuning().k1_lookup;
}

// y: [B, 256, ld] bf16; stats (nullable): [2][C1] doubles, accumulated (sum, sum of squares of the rounded outputs).
inline int onehot_conv_fwd_tc(const uint8_t* bases, const float* w, const float* bias, bf16* y, double* stats, int B, int C1, int k, int ld,
                              cudaStream_t st, const float* pool_scale = nullptr, const float* pool_shift = nullptr, bf16* pool_out = nullptr,
                              int pool_Lp = 0) {
    int rc = tc_init();
    if (rc) return rc;
    if (first_on_device(2)) {
        cudaError_t e = cudaFuncSetAttribute(onehot_conv_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OHF_SMEM);
        if (e != cudaSuccess) return set_error(-3, "cudaFuncSetAttribute(onehot_conv_fwd_tc_kernel): %s", cudaGetErrorString(e));
    }
    CUtensorMap map;
    rc = make_map(&map, y, C1, SEQ_LEN, B, ld, (int64_t)SEQ_LEN * ld, 64, 32, 1);
    if (rc) return rc;
    const uint32_t idesc = make_idesc(0, 0, 64);
    const int grid = std::min(B, 2 * tc_num_sms());
    onehot_conv_fwd_tc_kernel<<<grid, OHF_THREADS, OHF_SMEM, st>>>(map, bases, w, bias, pool_out ? nullptr : stats, B, C1, k, idesc, pool_scale, pool_shift,
                                                                   pool_out, pool_Lp);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return set_error(-3, "onehot_conv_fwd_tc launch failed: %s", cudaGetErrorString(err));
    return 0;
}

}  // namespace emb
