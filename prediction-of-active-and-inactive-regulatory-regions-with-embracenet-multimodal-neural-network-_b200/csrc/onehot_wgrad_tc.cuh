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
           !getenv("EMB_NO_ONEHOT_WGRAD_TC");
}

// dw[C1][4][k] += ... (fp32 atomics; the caller zeroes it).  dy: [B, 256, ld] bf16.
inline int onehot_conv_wgrad_tc(const uint8_t* bases, const bf16* dy, float* dw, int B, int C1, int k, int ld, cudaStream_t st) {
    int rc = tc_init();
    if (rc) return rc;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(onehot_conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OHW_SMEM);
        if (e != cudaSuccess) return set_error(-3, "cudaFuncSetAttribute(onehot_conv_wgrad_tc_kernel): %s", cudaGetErrorString(e));
        attr = true;
    }
    CUtensorMap map;
    rc = make_map(&map, dy, C1, SEQ_LEN, B, ld, (int64_t)SEQ_LEN * ld, 64, SEQ_LEN, 1);
    if (rc) return rc;
    const uint32_t idesc = make_idesc(1, 1, 128);
    const int grid = std::min(B, tc_num_sms());
    onehot_conv_wgrad_tc_kernel<<<grid, OHW_THREADS, OHW_SMEM, st>>>(map, bases, dw, B, C1, k, idesc);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return set_error(-3, "onehot_conv_wgrad_tc launch failed: %s", cudaGetErrorString(err));
    return 0;
}

}  // namespace emb
