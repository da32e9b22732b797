// Inference form of the FIRST conv layer in ONE kernel:
//     a0 = MaxPool1d(10, 2)( ReLU( BatchNorm_eval( Conv1d_{4 -> C1}(onehot(bases)) ) ) )        (CNN_pre.py:39-49 in eval mode, in_channels = 4)
//
// The training-shaped forward (onehot_conv_fwd_tc_kernel) writes the pre-pooling output y0 [B, 256, C1] and the pooling kernel
// reads it back: at inference that round trip is 2 x 16 KB per region for 8 KB of pooled output (arch S), and the two kernels
// were 0.77 of the 2.06 ms a 65 536-region forward took.  Here the product is TRANSPOSED, as in conv_pool_tc.cuh, so that a TMEM
// lane is an output channel and an epilogue thread pools along its own registers -- with one more step, because C1 (16 / 32 / 64)
// would leave most of the 128 lanes empty:
//
//   the 256 positions are cut into nb = 128 / C1 BLOCKS of Pq = ceil(Lp / nb) pooled rows (2 Pq + 8 positions, overlapping by the
//   8 positions a window reaches over), and block q is computed into lanes [q C1, (q + 1) C1):
//       D[q C1 + o, n] = sum_{t, c} W[o, c, t] * onehot(base[2 q Pq + n + t - pad])[c]
//   A (M x K) = the weights as a "tall" K-major array [128 - C1 zero rows | C1 weight rows | 128 - C1 zero rows]; the 128-row
//               window that starts (128 - C1) - q C1 rows down has the weights exactly in rows [q C1, (q + 1) C1) and zeros
//               elsewhere, so every block uses the SAME shared-memory weights through a shifted descriptor;
//   B (N x K) = the Toeplitz view of the sample's one-hot rows (onehot_wgrad_tc.cuh), started 2 q Pq rows down.
//   nb x ceil(k / 2) tcgen05.mma (128 x N x 16) per sample (two tap slots of 8 K elements each); the weight array holds
//   Kw / 8 = 8 (k <= 7) or 16 (k <= 15) tap slots.
//
// The fp32 master weight enters as a two-term bf16 split {hi, mid} against the one-hot stored twice per row (2^-17 relative;
// the training forward needs the exact three-term split because one bf16 ulp re-routes max-pool gradients; inference does not).
// Rounding points otherwise equal conv_pool_tc.cuh: fp32 accumulator, conv bias folded into the BatchNorm shift, one bf16 rounding
// of the pooled value.
//
//   warp 0 producer (base codes, bulk copy)   warp 1 MMA issuer   warps 2-3 generators (codes -> one-hot rows)
//   warps 4-19 epilogue: TMEM lane quarter = warp % 4; `parts` warps per quarter share a sample's pooled rows and 4 / parts warp
//   sets take samples round-robin; 512 / N accumulators are in flight (the MMA -> epilogue -> MMA hand-over costs ~1 us)
#pragma once
#include "onehot_wgrad_tc.cuh"
#include "conv_pool_tc.cuh"

namespace emb {

constexpr int OHP_STAGES = 6;                               // one-hot row arrays in flight (freed by the MMAs that read them)
constexpr int OHP_CODES = 16;                               // base-code buffers in flight (freed by the generators): the 256-byte loads are pure
                                                            // DRAM latency, and with one buffer per row array the MMA warp starved (r2 ncu: 63 %
                                                            // of the kernel's stall samples were epilogue warps waiting for an accumulator)
constexpr int OHP_ROWS = 288;                               // one-hot rows per stage: 2 (nb - 1) Pq + N + 16 <= 288 for every C1
constexpr int OHP_STAGE = OHP_ROWS * 16;
constexpr int OHP_DATA_BYTES = OHP_STAGES * OHP_STAGE + OHP_CODES * SEQ_LEN;
constexpr int OHP_EPI_WARPS = 16;
constexpr int OHP_THREADS = 128 + OHP_EPI_WARPS * 32;
constexpr int OHP_MAX_BUF = 8;

struct OhpParams {
    int B, C1, k, Lp;
    int nb, Pq, N, Kw;               // position blocks, pooled rows per block, MMA N (positions per block, padded), K per block
    int NS, nbuf, parts;             // TMEM column stride of one accumulator, accumulators in flight, epilogue warps per quarter sharing a sample
    int w_bytes, sbo;                // tall weight array: bytes (rounded), bytes between 8-row groups
    uint32_t idesc;
    int dbg;                         // EMB_CONV_DEBUG bisection: 1 = epilogue does not pool, 8 = test_wait spin instead of try_wait
};

// C1T = C1 = the output row stride (16 / 32 / 64); KT = the kernel size when it is 5 / 11 / 15, 0 = run-time (the MMA-issuing thread's
// loop then unrolls into descriptor adds with immediates: see conv_pool_tc.cuh)
template <int C1T, int KT>
__global__ void __launch_bounds__(OHP_THREADS, 1)
onehot_conv_pool_tc_kernel(const uint8_t* __restrict__ bases, const float* __restrict__ w, const float* __restrict__ bias,
                           const float* __restrict__ scale, const float* __restrict__ shift, bf16* __restrict__ out, const OhpParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* wsm = smem;                                            // [256 - C1 rows][Kw] bf16, un-swizzled K-major core matrices
    uint8_t* stages = wsm + p.w_bytes;                              // [stage] rows[288][16 B]
    uint8_t* codes_sm = stages + OHP_STAGES * OHP_STAGE;            // [code buffer][256]
    uint64_t* codes_full = (uint64_t*)(codes_sm + OHP_CODES * SEQ_LEN);
    uint64_t* codes_empty = codes_full + OHP_CODES;
    uint64_t* rows_full = codes_empty + OHP_CODES;
    uint64_t* rows_empty = rows_full + OHP_STAGES;
    uint64_t* tfull = rows_empty + OHP_STAGES;
    uint64_t* tempty = tfull + OHP_MAX_BUF;
    uint32_t* tmem_slot = (uint32_t*)(tempty + OHP_MAX_BUF);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int pad = (p.k - 1) / 2;
    const int n_mine = (int)blockIdx.x < p.B ? (p.B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int parts = p.parts, sets = (OHP_EPI_WARPS / 4) / parts;  // row parts per sample; warp sets take samples round-robin

    // zero rows of the weight array, the halo rows and the unused tap slots read as zero for the whole kernel
    for (int i = threadIdx.x; i < (p.w_bytes + OHP_DATA_BYTES) / 16; i += OHP_THREADS) ((uint4*)smem)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    {
        const int slots = p.Kw >> 3;                                // tap slots of 8 K elements: {onehot hi-part, onehot mid-part}
        for (int i = threadIdx.x; i < p.C1 * slots * 4; i += OHP_THREADS) {
            const int o = i / (slots * 4), t = (i >> 2) % slots, c = i & 3;
            const float v = t < p.k ? w[((size_t)o * 4 + c) * p.k + t] : 0.f;
            const bf16 hi = __float2bfloat16_rn(v);
            const bf16 mid = __float2bfloat16_rn(v - __bfloat162float(hi));
            const int R = 128 - p.C1 + o;
            bf16* dst = (bf16*)(wsm + (R >> 3) * p.sbo + t * 128 + (R & 7) * 16);
            dst[c] = hi;
            dst[4 + c] = mid;
        }
    }
    fence_async_smem();
    if (threadIdx.x == 0) {
        for (int s = 0; s < OHP_STAGES; ++s) { mbar_init(&rows_full[s], 2); mbar_init(&rows_empty[s], 1); }
        for (int s = 0; s < OHP_CODES; ++s) { mbar_init(&codes_full[s], 1); mbar_init(&codes_empty[s], 2); }
        for (int a = 0; a < OHP_MAX_BUF; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4 * parts); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= producer: the samples' 256 base codes, up to OHP_CODES samples ahead =================
        if (lane == 0) {
            int cs = 0;
            uint32_t phase = 0;
            for (int i = 0; i < n_mine; ++i) {
                const int b = blockIdx.x + i * gridDim.x;
                POOL_WAIT(&codes_empty[cs], phase ^ 1);
                mbar_expect_tx(&codes_full[cs], SEQ_LEN);
                bulk_load(codes_sm + cs * SEQ_LEN, bases + (size_t)b * SEQ_LEN, SEQ_LEN, &codes_full[cs]);
                if (++cs == OHP_CODES) { cs = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        const uint64_t dw_base = umma_desc_noswizzle(0, 128, (uint32_t)p.sbo), dx_base = umma_desc_noswizzle(0, 16, 128);
        const int ksteps = (p.k + 1) >> 1;                  // 16 K elements = two tap slots; slots >= k hold zero weights and are skipped
        int stage = 0, buf = 0;
        uint32_t phase = 0, use = 0;
        for (int i = 0; i < n_mine; ++i) {
            POOL_WAIT(&rows_full[stage], phase);
            POOL_WAIT(&tempty[buf], (use & 1u) ^ 1u);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint32_t rows = smem_u32(stages + stage * OHP_STAGE);
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * p.NS);
                if (KT) {
                    constexpr int NB = 128 / C1T, PQ = ((SEQ_LEN - 10) / 2 + 1 + NB - 1) / NB, KW = KT <= 7 ? 64 : 128;
                    const uint32_t wa = smem_u32(wsm) + (uint32_t)((128 - C1T) >> 3) * (uint32_t)(KW * 16);
                    const uint64_t dw = dw_base | (uint64_t)((wa & 0x3FFFFu) >> 4), dx = dx_base | (uint64_t)((rows & 0x3FFFFu) >> 4);
#pragma unroll
                    for (int q = 0; q < NB; ++q)
#pragma unroll
                        for (int s = 0; s < (KT + 1) / 2; ++s)
                            tc_mma_f16(d_tmem, dw - (uint64_t)(q * (C1T / 8) * KW) + (uint64_t)(s * 16), dx + (uint64_t)(q * 2 * PQ + s * 2), p.idesc,
                                       (q | s) ? 1u : 0u);
                } else {
                    for (int q = 0; q < p.nb; ++q) {
                        const uint32_t wa = smem_u32(wsm) + (uint32_t)((128 - p.C1 - q * p.C1) >> 3) * (uint32_t)p.sbo;
                        const uint32_t xa = rows + (uint32_t)(2 * q * p.Pq) * 16u;
                        const uint64_t dw = dw_base | (uint64_t)((wa & 0x3FFFFu) >> 4), dx = dx_base | (uint64_t)((xa & 0x3FFFFu) >> 4);
                        for (int s = 0; s < ksteps; ++s)            // 16 K elements = two tap slots: weights + 2 K blocks (256 B), rows + 2 (32 B)
                            tc_mma_f16(d_tmem, dw + (uint64_t)(s * 16), dx + (uint64_t)(s * 2), p.idesc, (q | s) ? 1u : 0u);
                    }
                }
                tc_commit(&tfull[buf]);
                tc_commit(&rows_empty[stage]);
            }
            __syncwarp();
            if (++stage == OHP_STAGES) { stage = 0; phase ^= 1; }
            if (++buf == p.nbuf) { buf = 0; ++use; }
        }
    } else if (warp < 4) {
        // ================= generators: 256 code bytes -> 256 rows {onehot, onehot} of 16 bytes =================
        int stage = 0, cs = 0;
        uint32_t phase = 0, cphase = 0;
        const int t = threadIdx.x - 64;                   // 0..63
        for (int i = 0; i < n_mine; ++i) {
            POOL_WAIT(&codes_full[cs], cphase);
            POOL_WAIT(&rows_empty[stage], phase ^ 1);      // the MMAs that read this row array OHP_STAGES samples ago are complete
            const uint8_t* codes = codes_sm + cs * SEQ_LEN;
            uint4* rows = (uint4*)(stages + stage * OHP_STAGE) + pad;
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const uint32_t c = codes[t + h * 64];
                const uint32_t one = (c & 1) ? 0x3F800000u : 0x00003F80u;          // bf16 1.0 in the odd / even half of a word
                const uint32_t w0 = c < 2 ? one : 0u, w1 = (c & ~1u) == 2 ? one : 0u;
                rows[t + h * 64] = make_uint4(w0, w1, w0, w1);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&rows_full[stage]); mbar_arrive(&codes_empty[cs]); }
            if (++stage == OHP_STAGES) { stage = 0; phase ^= 1; }
            if (++cs == OHP_CODES) { cs = 0; cphase ^= 1; }
        }
    } else {
        // ================= epilogue: thread = (position block, channel); its registers walk the block's positions =================
        // The pair loop is pool_walk of conv_pool_tc.cuh.
        const int e = warp - 4, wq = e & 3, sub = e >> 2;
        const int set = sub % sets, part = sub / sets;
        const int li = wq * 32 + lane, qb = li / p.C1, ch = li - qb * p.C1;
        const float sc = scale[ch];
        const float shb = fmaf(bias[ch], sc, shift[ch]);
        const int per = (p.Pq + parts - 1) / parts;
        const int lo = min(part * per, p.Pq), hi = min(lo + per, p.Pq);          // pooled rows [lo, hi) of the block, block-local
        const int valid = max(0, min(p.Pq, p.Lp - qb * p.Pq));                   // the last block may be short
        const unsigned n_eff = (unsigned)max(0, min(hi, valid) - lo);
        const int P_first = lo + 4, P_end = hi + 4;                              // the pair that completes window j is pair j + 4
        const int c_first = (2 * lo) & ~15;
        const int k_first = (c_first >> 1) - P_first;
        const long long ld = p.C1;
        for (int i = set; i < n_mine; i += sets) {
            const int b = blockIdx.x + i * gridDim.x;
            const int buf = i % p.nbuf;
            POOL_WAIT(&tfull[buf], (uint32_t)(i / p.nbuf) & 1u);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(buf * p.NS);
            bf16* o = out + ((long long)b * p.Lp + qb * p.Pq + lo + k_first) * ld + ch;
            if (!(p.dbg & 1)) pool_walk<C1T>(t_addr, c_first, 2 * P_end, k_first, n_eff, o, ld, sc, shb);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

inline OhpParams onehot_pool_params(int B, int C1, int k, int Lp) {
    OhpParams p = {};
    p.B = B; p.C1 = C1; p.k = k; p.Lp = Lp;
    p.nb = 128 / C1;
    p.Pq = cdiv(Lp, p.nb);
    p.N = round_up(2 * p.Pq + 8, 16);
    p.Kw = k <= 7 ? 64 : 128;
    p.NS = p.N;                                  // accumulators packed back to back: the hand-over latency (~1 us) wants as many in flight as fit
    p.nbuf = std::min(OHP_MAX_BUF, 512 / p.NS);
    p.parts = 1;                                 // measured (profiles/r02_infer_parts.sh): whole samples per warp beat shared ones
    if (tuning().pool_parts == 1 || tuning().pool_parts == 2 || tuning().pool_parts == 4) p.parts = tuning().pool_parts;
    p.sbo = p.Kw * 16;
    p.w_bytes = round_up(((256 - C1) / 8) * p.sbo, 1024);
    p.idesc = make_idesc(0, 0, p.N);
    p.dbg = tuning().conv_debug;
    return p;
}

inline bool onehot_pool_tc_ok(const uint8_t* bases, int C1, int k, int ld_out, int Lp) {
    if (!(C1 == 16 || C1 == 32 || C1 == 64) || ld_out != C1 || k < 1 || k > 15 || !(k & 1) || ((uintptr_t)bases & 15)) return false;
    if (Lp != (SEQ_LEN - 10) / 2 + 1) return false;
    const OhpParams p = onehot_pool_params(1, C1, k, Lp);
    return 2 * (p.nb - 1) * p.Pq + p.N + (p.Kw >> 3) <= OHP_ROWS && p.N <= 256;
}

// out: [B, Lp, C1] bf16 (channels-last, the next layer's input); scale / shift: eval-mode BatchNorm per channel
inline int onehot_conv_pool_tc(const uint8_t* bases, const float* w, const float* bias, const float* scale, const float* shift, bf16* out,
                               int B, int C1, int k, int Lp, cudaStream_t st) {
    int rc = tc_init();
    if (rc) return rc;
    const OhpParams p = onehot_pool_params(B, C1, k, Lp);
    const size_t smem = 1024 + (size_t)p.w_bytes + OHP_DATA_BYTES + 512;
    using Kern = void (*)(const uint8_t*, const float*, const float*, const float*, const float*, bf16*, const OhpParams);
#define EMB_OHP_ROW(C) {onehot_conv_pool_tc_kernel<C, 0>, onehot_conv_pool_tc_kernel<C, 5>, onehot_conv_pool_tc_kernel<C, 11>, onehot_conv_pool_tc_kernel<C, 15>}
    static const Kern kerns[3][4] = {EMB_OHP_ROW(16), EMB_OHP_ROW(32), EMB_OHP_ROW(64)};
#undef EMB_OHP_ROW
    if (first_on_device(4)) {
        for (int i = 0; i < 12; ++i) {
            cudaError_t e = cudaFuncSetAttribute(kerns[i / 4][i % 4], cudaFuncAttributeMaxDynamicSharedMemorySize, tc_max_smem());
            if (e != cudaSuccess) return set_error(-3, "cudaFuncSetAttribute(onehot_conv_pool_tc_kernel): %s", cudaGetErrorString(e));
        }
    }
    if (smem > (size_t)tc_max_smem()) return set_error(-5, "onehot_conv_pool_tc: shared memory");
    const int grid = std::min(B, tc_num_sms());
    kerns[C1 == 16 ? 0 : C1 == 32 ? 1 : 2][k == 5 ? 1 : k == 11 ? 2 : k == 15 ? 3 : 0]<<<grid, OHP_THREADS, smem, st>>>(bases, w, bias, scale, shift, out, p);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return set_error(-3, "onehot_conv_pool_tc launch failed: %s", cudaGetErrorString(err));
    return 0;
}

}  // namespace emb
