// Shared device/host helpers for libembrace_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdlib.h>
#include <string.h>

namespace emb {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// error plumbing: every ABI entry returns a code, the text is kept thread-local
// ---------------------------------------------------------------------------------------------
extern thread_local std::string g_last_error;
int set_error(int code, const char* fmt, ...);

#define EMB_CUDA_OK(expr)                                                                              \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess)                                                                         \
            return emb::set_error(-3, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                                  __LINE__);                                                           \
    } while (0)

#define EMB_CHECK_LAUNCH()                                                                          \
    do {                                                                                            \
        cudaError_t _e = cudaGetLastError();                                                        \
        if (_e != cudaSuccess)                                                                      \
            return emb::set_error(-3, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),   \
                                  __FILE__, __LINE__);                                              \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Tuning switches.  Read ONCE per process from the environment (first use), never on the launch path; every switch can
// also be set through the ABI (emb_set_option(name, value), include/embrace_b200.h lists them).  All default to the
// production configuration; the non-default values exist for A/B measurements and as verification fall-backs.
// ---------------------------------------------------------------------------------------------
struct Tuning {
    int k1_pairs = 0;            // EMB_K1_PAIRS            K1 lookup forward: tap-pair tables instead of tap-triple tables
    int k1_lookup = 0;           // EMB_K1_LOOKUP           K1 forward as the shared-memory gather-sum even in bf16 mode (no tensor core)
    int no_onehot_wgrad_tc = 0;  // EMB_NO_ONEHOT_WGRAD_TC  K1 backward as the position-list histogram
    int epi_stats = 0;           // EMB_EPI_STATS           BatchNorm statistics inside the conv GEMM epilogue (no separate pass)
    int no_tma_k2 = 0;           // EMB_NO_TMA_K2           pooling backward without the bulk-copy kernels / arg-max codes
    int prof_dump = 0;           // EMB_PROF_DUMP           emb_profile_read prints one line per timed GEMM launch
    int conv_reuse = 1;          // EMB_CONV_REUSE          0: every conv on the per-tap kernel, 1: tap-reuse kernel where it applies
    int conv_debug = 0;          // EMB_CONV_DEBUG          tap-reuse kernel: skip loads / MMAs / epilogue (timing experiments)
    int conv_no_resident = 0;    // EMB_CONV_NO_RESIDENT    tap-reuse kernel: always stream the weights
    int conv_tps1 = 0;           // EMB_CONV_TPS1           tap-reuse kernel: one tap per streamed weight stage
    int min_kiters = 8;          // EMB_MIN_KITERS          split-K: minimum K blocks per CTA
    int wgrad_taps = 0;          // EMB_WGRAD_TAPS          conv wgrad taps per CTA (0: as many as TMEM holds, 1: per-tap kernel)
    int wgrad_ntile = 128;       // EMB_WGRAD_NT            conv wgrad N tile
    int wgrad_fuse_taps = 1;     // EMB_WGRAD_FUSE_TAPS     conv wgrad with Cin = 64: up to four taps per tcgen05.mma (verified r2; 0 = one MMA per tap)
    int deterministic = 0;       // EMB_DETERMINISTIC       fixed-order reductions: no split-K, one CTA per reduction column block
    int k2_wide = 1;             // EMB_K2_WIDE             8-channel (16-byte) forward pooling kernel where C % 8 == 0
    int infer_fuse = 2;          // EMB_INFER_FUSE          eval forward without the pre-pooling tensor: 2 = transposed conv + in-register pooling
                                 //                         (conv_pool_tc.cuh), 1 = pooling in the row-major GEMM epilogue (four epilogue warps, through
                                 //                         shared memory: measured 2.3x slower than 0 = unfused)
    int infer_fuse_k1 = 1;       // EMB_INFER_FUSE_K1       with infer_fuse = 2: the first (one-hot) conv layer through onehot_pool_tc.cuh as well
    int pool_parts = 0;          // EMB_POOL_PARTS          pooled inference kernels: epilogue warps sharing one tile (0 = heuristic; 1, 2, 4)
    int tc_min_mflop = 0;        // EMB_TC_MIN_MFLOP        Linear GEMMs below this many MFLOP run on the SIMT kernel (0: tensor cores whenever the shape allows)
    int fork = 1;                // EMB_FORK                independent branches of the step on side streams (parallel graph branches)
};
struct TuningName { const char* env; const char* name; int Tuning::*field; };
inline const TuningName* tuning_names(int* n) {
    static const TuningName t[] = {
        {"EMB_K1_PAIRS", "k1_pairs", &Tuning::k1_pairs}, {"EMB_K1_LOOKUP", "k1_lookup", &Tuning::k1_lookup},
        {"EMB_NO_ONEHOT_WGRAD_TC", "no_onehot_wgrad_tc", &Tuning::no_onehot_wgrad_tc}, {"EMB_EPI_STATS", "epi_stats", &Tuning::epi_stats},
        {"EMB_NO_TMA_K2", "no_tma_k2", &Tuning::no_tma_k2}, {"EMB_PROF_DUMP", "prof_dump", &Tuning::prof_dump},
        {"EMB_CONV_REUSE", "conv_reuse", &Tuning::conv_reuse}, {"EMB_CONV_DEBUG", "conv_debug", &Tuning::conv_debug},
        {"EMB_CONV_NO_RESIDENT", "conv_no_resident", &Tuning::conv_no_resident}, {"EMB_CONV_TPS1", "conv_tps1", &Tuning::conv_tps1},
        {"EMB_MIN_KITERS", "min_kiters", &Tuning::min_kiters}, {"EMB_WGRAD_TAPS", "wgrad_taps", &Tuning::wgrad_taps},
        {"EMB_WGRAD_NT", "wgrad_ntile", &Tuning::wgrad_ntile}, {"EMB_WGRAD_FUSE_TAPS", "wgrad_fuse_taps", &Tuning::wgrad_fuse_taps},
        {"EMB_DETERMINISTIC", "deterministic", &Tuning::deterministic}, {"EMB_K2_WIDE", "k2_wide", &Tuning::k2_wide},
        {"EMB_FORK", "fork", &Tuning::fork}, {"EMB_TC_MIN_MFLOP", "tc_min_mflop", &Tuning::tc_min_mflop}, {"EMB_INFER_FUSE", "infer_fuse", &Tuning::infer_fuse},
        {"EMB_INFER_FUSE_K1", "infer_fuse_k1", &Tuning::infer_fuse_k1}, {"EMB_POOL_PARTS", "pool_parts", &Tuning::pool_parts},
    };
    *n = (int)(sizeof t / sizeof t[0]);
    return t;
}
inline Tuning& tuning() {
    static Tuning t = [] {
        Tuning v;
        int n = 0;
        const TuningName* names = tuning_names(&n);
        for (int i = 0; i < n; ++i) {
            const char* e = getenv(names[i].env);
            if (e) v.*(names[i].field) = (*e == 0) ? 1 : atoi(e);       // an empty value means "on"
        }
        return v;
    }();
    return t;
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline int64_t round_up64(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// activation element type: float (EMB_PREC_FP32) or bf16 (EMB_PREC_BF16)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float load_act(const void* p, int dtype, size_t i) {
    return dtype ? __bfloat162float(((const bf16*)p)[i]) : ((const float*)p)[i];
}
__device__ __forceinline__ void store_act(void* p, int dtype, size_t i, float v) {
    if (dtype) ((bf16*)p)[i] = __float2bfloat16_rn(v);
    else ((float*)p)[i] = v;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10: counter-based generator, keyed by (seed); counter = (element index, stream, step).
// Draws therefore do not depend on how a batch is partitioned across GPUs or thread blocks.
// ---------------------------------------------------------------------------------------------
struct RngState {
    uint64_t seed;
    uint32_t step;
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

enum RngStream : uint32_t {
    RNG_FFNN_DROP = 0x10, RNG_CNN_DROP = 0x20, RNG_POST_DROP = 0x30, RNG_EMBRACE = 0x40, RNG_MODAL_COIN = 0x50,
    RNG_MODAL_ROWS = 0x51
};

__device__ __forceinline__ uint4 rng_raw(const RngState& s, uint32_t stream, uint64_t elem) {
    uint4 ctr = make_uint4((uint32_t)elem, (uint32_t)(elem >> 32), stream, s.step);
    uint2 key = make_uint2((uint32_t)s.seed, (uint32_t)(s.seed >> 32));
    return philox4x32_10(ctr, key);
}
// Element e of a stream uses word (e & 3) of the Philox block with counter (e >> 2): four draws per block, and the value
// depends only on (seed, step, stream, e) -- not on tiling, back end or data-parallel partition.
__device__ __forceinline__ uint32_t rng_word(const uint4& r, uint32_t i) { return i == 0 ? r.x : i == 1 ? r.y : i == 2 ? r.z : r.w; }
__device__ __forceinline__ float u32_to_unit_f32(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }     // [0,1), 24 bits
__device__ __forceinline__ double u32_to_unit_f64(uint32_t w) { return (double)w * (1.0 / 4294967296.0); }           // [0,1), 32 bits
__device__ __forceinline__ float rng_uniform_f32(const RngState& s, uint32_t stream, uint64_t elem) {
    return u32_to_unit_f32(rng_word(rng_raw(s, stream, elem >> 2), (uint32_t)elem & 3u));
}
// the modality choice compares in fp64 (EmbraceNetMultimodal.py:84 via torch.multinomial); 32 random bits are ample for it
__device__ __forceinline__ double rng_uniform_f64(const RngState& s, uint32_t stream, uint64_t elem) {
    return u32_to_unit_f64(rng_word(rng_raw(s, stream, elem >> 2), (uint32_t)elem & 3u));
}

// Dropout draws of the CNN stack ([B, C, Lpool] in the reference's layout): ONE Philox block serves a channel PAIR x four
// pooled positions j = 4q .. 4q+3 as eight 16-bit uniforms (word j & 3: low half = even channel, high half = odd channel).
// The generator dominated the instruction count of the pooling kernel with one 32-bit draw per element (two blocks per
// thread and four positions); a 16-bit keep/drop decision quantises p to 1/65536.
__device__ __forceinline__ uint4 rng_cnn_block(const RngState& s, uint32_t stream, uint64_t global_row, int C, int cpair, int Lp, int jq) {
    const uint64_t q_per_row = (uint64_t)((Lp + 3) >> 2), pairs = (uint64_t)((C + 1) >> 1);
    return rng_raw(s, stream, (global_row * pairs + cpair) * q_per_row + jq);
}
__device__ __forceinline__ float rng_cnn_u16(const uint4& r, uint32_t jj, uint32_t odd) {
    const uint32_t w = rng_word(r, jj);
    return ((float)(odd ? (w >> 16) : (w & 0xFFFFu)) + 0.5f) * (1.0f / 65536.0f);      // (0, 1), exact in fp32
}
__device__ __forceinline__ float rng_cnn_uniform(const RngState& s, uint32_t stream, uint64_t global_row, int C, int c, int Lp, int j) {
    return rng_cnn_u16(rng_cnn_block(s, stream, global_row, C, c >> 1, Lp, j >> 2), (uint32_t)j & 3u, (uint32_t)c & 1u);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace emb
