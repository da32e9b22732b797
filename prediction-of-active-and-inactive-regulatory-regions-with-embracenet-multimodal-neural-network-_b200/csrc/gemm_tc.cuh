// tcgen05 / TMEM / TMA GEMM back end (bf16 operands, fp32 accumulation in tensor memory), sm_100a.
//
// One warp-specialised kernel serves every GEMM-shaped op of the step:
//   nn.Linear fwd / dgrad / wgrad and the Conv1d (layers >= 1) implicit GEMMs fwd / dgrad / wgrad.
// Nothing is materialised for the convolutions: a 3-D TMA tensor map over the channels-last activation
// [B, L, C] fetches, for tap t, the box starting at l = t - pad; rows that fall outside [0, L) are zero
// filled by the TMA unit, which is exactly the "same" zero padding.  Operands that are contiguous along
// M/N instead of K (weights in dgrad, both operands in wgrad) are consumed in place through MN-major
// shared-memory descriptors, so no transposed copies exist either.
//
//   warp 0     TMA producer   (cp.async.bulk.tensor.3d -> 128B-swizzled smem stages, mbarrier complete_tx)
//   warp 1     MMA issuer     (tcgen05.mma.cta_group::1.kind::f16, 128 x N x 16, accumulators in TMEM)
//   warps 2-5  epilogue       (tcgen05.ld 32x32b -> registers -> the same fused epilogues as the SIMT path)
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include "common.cuh"
#include "gemm_simt.cuh"

namespace emb {

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a mis-programmed pipeline traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}
// cute::elect_one_sync(): exactly one lane of the (converged) warp gets a non-zero result
__device__ __forceinline__ uint32_t elect_one_sync() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, %1;\n\t"
        "@px mov.s32 %0, 1;\n\t}"
        : "+r"(pred) : "r"(0xFFFFFFFFu));
    return pred;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// the same load without the wait: several can be in flight before one tc_ld_wait() (the registers are undefined until then)
__device__ __forceinline__ void tc_ld16_nowait(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): SWIZZLE_128B, version 1 (Blackwell)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// ---------------------------------------------------------------------------------------------
// kernel parameters
// ---------------------------------------------------------------------------------------------
struct TcOperand {
    int mn_major;       // 0: K-major tile [rows][64 k] ; 1: MN-major blocks of [k rows][64 mn]
    int boxes;          // TMA boxes per stage (MN-major: one per 64-wide block)
    int box_bytes;      // bytes one box deposits (expect_tx)
    int block_bytes;    // smem distance between consecutive boxes of a stage (>= box_bytes, multiple of 1024)
    int stage_bytes;    // boxes * block_bytes
    int base[3];        // TMA coordinates = base + tap*ctap + j*cj + tile*ctile + blk*cblk
    int ctap[3], cj[3], ctile[3], cblk[3];
    int kstep_bytes;    // descriptor start-address advance per UMMA K step (16 elements)
    int lbo_bytes;
    int dst_off;        // byte offset of the TMA destination inside a block (guard rows in front of a tap-shifted tile)
};

struct TcParams {
    TcOperand a, b;
    int M, N;               // logical output extent (masking)
    int n_logical;          // N handed to the epilogue (index arithmetic of replayed draws / idx)
    int n_taps, n_inner;    // K loop = n_taps x n_inner stages (per K split)
    int j_total;            // wgrad: total row blocks over all splits (guards the last split)
    int split_k;            // blockIdx.z = tap_z * split_k + ks when tap_in_z
    int tap_in_z;           // wgrad conv: the tap is a grid coordinate, not a loop
    int n_off_per_tap;      // wgrad conv: logical n = tap * Cin + n
    int k_steps;            // UMMA K steps per stage
    int n_tile;             // UMMA N (multiple of 16, <= 256)
    int rows_per_group;     // epilogue row mapping: tile row r -> group r / rpg, element r % rpg
    int groups_per_tile;
    int stages;
    int zero_smem;          // MN-major K rows beyond the box must read as zero
    uint32_t idesc;
    uint32_t tmem_cols;     // total allocation = 2 accumulator stages
    int acc_stride;         // TMEM columns per accumulator stage
    int acc_stages;         // 2: epilogue of tile i overlaps the main loop of tile i+1; 1: all of TMEM is one tile's
    int taps_per_cta;       // multi-tap wgrad: this many taps accumulate side by side in TMEM from ONE staged operand pair;
                            // tap t reads the B tile through a descriptor shifted by (tap - pad) rows of 128 bytes
    int tap_pad;
    int fuse_taps;          // multi-tap wgrad with n_tile == 64 (experimental, EMB_WGRAD_FUSE_TAPS): this many taps per tcgen05.mma -- the 64-wide
                            // N blocks of ONE instruction are the same B tile 128 bytes (one row = one tap) apart, i.e. LBO = 128
    int grid_m, grid_n, grid_z, total_tiles;
};

constexpr int TC_THREADS = 192;
constexpr int TC_MAX_STAGES = 6;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// 16 consecutive columns of one output row: vectorised paths for the two epilogues that carry almost all of the
// traffic (plain bias[+ReLU] store and the ReLU/Dropout mask of a dgrad), generic per-element path otherwise.
__device__ __forceinline__ void tc_epilogue_row16(const Epilogue& ep, int m, int n0, int n_off, int M, int N, int n_logical, float* v) {
    const bool full16 = n0 + 16 <= N;
    const bool vec = full16 && ep.out_dtype == 1 && (ep.ldo & 7) == 0;
    if (vec && ep.mode == EPI_LINEAR && ep.drop_p == 0.f && ep.flatC == 0) {
        if (ep.bias) {
            const float4* b4 = reinterpret_cast<const float4*>(ep.bias + n0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 b = __ldg(b4 + i);
                v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
            }
        }
        if (ep.relu) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        uint4* dst = reinterpret_cast<uint4*>((bf16*)ep.out + (size_t)m * ep.ldo + n0);
        dst[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        dst[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
        return;
    }
    if (vec && ep.mode == EPI_LINEAR && ep.drop_p > 0.f && ep.flatC == 0 && ep.drop_u == nullptr && (N & 3) == 0) {
        // nn.Linear -> ReLU -> Dropout with the counter-based generator: 4 Philox blocks cover the 16 columns
        if (ep.bias) {
            const float4* b4 = reinterpret_cast<const float4*>(ep.bias + n0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 b = __ldg(b4 + i);
                v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
            }
        }
        const float inv_keep = 1.f / (1.f - ep.drop_p);
        const RngState rs = *ep.rng;
        const uint64_t e0 = (uint64_t)(ep.row_offset + m) * N + n0;      // multiple of 4
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint4 r = rng_raw(rs, ep.rng_stream, (e0 >> 2) + i);
            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float x = v[4 * i + j];
                if (ep.relu) x = fmaxf(x, 0.f);
                v[4 * i + j] = (u32_to_unit_f32(w[j]) >= ep.drop_p) ? x * inv_keep : 0.f;
            }
        }
        uint4* dst = reinterpret_cast<uint4*>((bf16*)ep.out + (size_t)m * ep.ldo + n0);
        dst[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        dst[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
        return;
    }
    if (vec && ep.mode == EPI_EMBRACE && ep.emb_u == nullptr && (N & 15) == 0 && (ep.ld_d0 & 7) == 0) {
        // docking_1 epilogue: d1 = relu(acc + b1); idx = (u > cum0[m]); e = idx ? d1 : d0; idx stored as bytes
        const float4* b4 = reinterpret_cast<const float4*>(ep.bias + n0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 b = __ldg(b4 + i);
            v[4 * i] = fmaxf(v[4 * i] + b.x, 0.f); v[4 * i + 1] = fmaxf(v[4 * i + 1] + b.y, 0.f);
            v[4 * i + 2] = fmaxf(v[4 * i + 2] + b.z, 0.f); v[4 * i + 3] = fmaxf(v[4 * i + 3] + b.w, 0.f);
        }
        const uint4* d4 = reinterpret_cast<const uint4*>((const bf16*)ep.d0 + (size_t)m * ep.ld_d0 + n0);
        const uint4 da = d4[0], db = d4[1];
        const uint32_t dw[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
        // u = w / 2^32 > cum0  <=>  w > floor(cum0 * 2^32)   (integer compare; cum0 == 1 can never be exceeded)
        const unsigned long long thr = (unsigned long long)floor(ep.cum0[m] * 4294967296.0);
        const RngState rs = *ep.rng;
        const uint64_t e0 = (uint64_t)(ep.row_offset + m) * N + n0;
        uint32_t idw[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint4 r = rng_raw(rs, RNG_EMBRACE, (e0 >> 2) + i);
            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = 4 * i + j;
                const int id = (unsigned long long)w[j] > thr;
                const uint32_t pair = dw[k >> 1];
                const float d0v = __uint_as_float((k & 1) ? (pair & 0xFFFF0000u) : (pair << 16));
                v[k] = id ? v[k] : d0v;
                idw[i] |= (uint32_t)id << (8 * j);
            }
        }
        uint4* dst = reinterpret_cast<uint4*>((bf16*)ep.out + (size_t)m * ep.ldo + n0);
        dst[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        dst[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
        *reinterpret_cast<uint4*>(ep.idx_out + (size_t)m * N + n0) = make_uint4(idw[0], idw[1], idw[2], idw[3]);
        return;
    }
    if (vec && ep.mode == EPI_EMBRACE_BWD && (N & 15) == 0 && (ep.ld_e & 7) == 0) {
        const uint4* e4 = reinterpret_cast<const uint4*>((const bf16*)ep.e + (size_t)m * ep.ld_e + n0);
        const uint4 ea = e4[0], eb = e4[1];
        const uint32_t ew[8] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
        const uint4 iq = *reinterpret_cast<const uint4*>(ep.idx + (size_t)m * N + n0);
        const uint32_t iw[4] = {iq.x, iq.y, iq.z, iq.w};
        float a0[16], a1[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const uint32_t h = (k & 1) ? (ew[k >> 1] >> 16) : (ew[k >> 1] & 0xFFFFu);
            const bool pos = h != 0 && !(h & 0x8000u);
            const int id = (iw[k >> 2] >> (8 * (k & 3))) & 0xFF;
            const float g = pos ? v[k] : 0.f;
            a0[k] = id == 0 ? g : 0.f;
            a1[k] = id == 1 ? g : 0.f;
        }
        uint4* d0p = reinterpret_cast<uint4*>((bf16*)ep.out + (size_t)m * ep.ldo + n0);
        uint4* d1p = reinterpret_cast<uint4*>((bf16*)ep.out2 + (size_t)m * ep.ldo + n0);
        d0p[0] = make_uint4(pack_bf16x2(a0[0], a0[1]), pack_bf16x2(a0[2], a0[3]), pack_bf16x2(a0[4], a0[5]), pack_bf16x2(a0[6], a0[7]));
        d0p[1] = make_uint4(pack_bf16x2(a0[8], a0[9]), pack_bf16x2(a0[10], a0[11]), pack_bf16x2(a0[12], a0[13]), pack_bf16x2(a0[14], a0[15]));
        d1p[0] = make_uint4(pack_bf16x2(a1[0], a1[1]), pack_bf16x2(a1[2], a1[3]), pack_bf16x2(a1[4], a1[5]), pack_bf16x2(a1[6], a1[7]));
        d1p[1] = make_uint4(pack_bf16x2(a1[8], a1[9]), pack_bf16x2(a1[10], a1[11]), pack_bf16x2(a1[12], a1[13]), pack_bf16x2(a1[14], a1[15]));
        return;
    }
    if (vec && ep.mode == EPI_MASKGRAD && (ep.ld_ref & 7) == 0) {
        const uint4* r4 = reinterpret_cast<const uint4*>((const bf16*)ep.ref + (size_t)m * ep.ld_ref + n0);
        uint4 ra = r4[0], rb = r4[1];
        const uint32_t rw[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            // bf16 > 0  <=>  sign bit clear and magnitude non-zero
            uint32_t lo = rw[i] & 0xFFFFu, hi = rw[i] >> 16;
            v[2 * i] = (lo != 0 && !(lo & 0x8000u)) ? v[2 * i] * ep.scale : 0.f;
            v[2 * i + 1] = (hi != 0 && !(hi & 0x8000u)) ? v[2 * i + 1] * ep.scale : 0.f;
        }
        uint4* dst = reinterpret_cast<uint4*>((bf16*)ep.out + (size_t)m * ep.ldo + n0);
        dst[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        dst[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
        return;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        int n = n0 + i;
        if (n < N) epilogue_apply(ep, m, n_off + n, M, n_logical, v[i]);
    }
}

// Column sums over the 32 rows a warp holds (16 columns per lane): a transpose-reduce butterfly -- each step halves the
// number of columns a lane carries while doubling the rows folded into them (8 + 4 + 2 + 1 + 1 = 16 shuffles instead of
// 16 x 5).  On return lane l holds the total of column ((l >> 1) & 15) ... precisely: column index col16_of_lane(l).
__device__ __forceinline__ float warp_colsum16(const float* v, int lane) {
    float a[8];
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float send = hi ? v[i] : v[i + 8], keep = hi ? v[i + 8] : v[i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    float b[4];
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float send = hi ? a[i] : a[i + 4], keep = hi ? a[i + 4] : a[i];
            b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    float c2[2];
    {
        const bool hi = lane & 4;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float send = hi ? b[i] : b[i + 2], keep = hi ? b[i + 2] : b[i];
            c2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    float d;
    {
        const bool hi = lane & 2;
        const float send = hi ? c2[0] : c2[1], keep = hi ? c2[1] : c2[0];
        d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    return d;      // column = 8 * bit4 + 4 * bit3 + 2 * bit2 + bit1 of the lane index
}
__device__ __forceinline__ int col16_of_lane(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }

// BatchNorm statistics of a conv forward tile chunk: v = the 16 fp32 accumulators of this lane's row (bias not yet added).
// s_warp: this warp's PRIVATE [2][n_stat] accumulators (plain read-modify-write by one lane per column: no atomics).
__device__ __forceinline__ void tc_bn_stats16(const Epilogue& ep, float* s_warp, int n_stat, int n0, int N, bool row_ok, const float* v, int lane) {
    float r[16], q[16];
    const bool full = n0 + 16 <= N;
    float b[16];
    if (full && ep.bias) {
        const float4* b4 = reinterpret_cast<const float4*>(ep.bias + n0);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float4 t = __ldg(b4 + i); b[4 * i] = t.x; b[4 * i + 1] = t.y; b[4 * i + 2] = t.z; b[4 * i + 3] = t.w; }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) b[i] = (ep.bias && n0 + i < N) ? __ldg(ep.bias + n0 + i) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float x = (row_ok && n0 + i < N) ? __bfloat162float(__float2bfloat16_rn(v[i] + b[i])) : 0.f;
        r[i] = x;
        q[i] = x * x;
    }
    const float s1 = warp_colsum16(r, lane), s2 = warp_colsum16(q, lane);
    if (!(lane & 1)) {
        const int n = n0 + col16_of_lane(lane);
        if (n < N) { s_warp[n] += s1; s_warp[n_stat + n] += s2; }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// EPI_POOL: the inference forward never writes the pre-pooling conv output.  A conv tile holds WHOLE samples (rows = positions),
// so BatchNorm (eval: one scale/shift per channel) + ReLU + MaxPool1d(10, 2) run on the tile: the 128 epilogue threads stage
// relu(bn(.)) of a 16-column chunk in shared memory ([128 rows][20 floats]: the 80-byte row stride keeps the 16-byte stores of
// a quarter warp on different banks), meet at a named barrier, and then produce the pooled rows -- thread = (pooled row,
// channel pair), 10 reads each -- straight into the next layer's input.  Arithmetic and rounding points are those of the
// unfused path (bf16 conv output, fp32 BatchNorm apply, one bf16 rounding of the pooled value), so both produce the same bits.
// Two buffers alternate, so one barrier per chunk suffices.
// ---------------------------------------------------------------------------------------------
constexpr int TC_POOL_ROW = 20;
constexpr int TC_POOL_SCRATCH_BYTES = 2 * 128 * TC_POOL_ROW * 4;

__device__ __forceinline__ void tc_pool_chunk16(const Epilogue& ep, float* buf, int tid, const float* v, int n0, int N, int rows_per_sample,
                                                int samples_in_tile, int sample0, int Bn) {
    float z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int n = n0 + i;
        float t = 0.f;
        if (n < N) {
            const float y = __bfloat162float(__float2bfloat16_rn(v[i] + (ep.bias ? __ldg(ep.bias + n) : 0.f)));
            t = fmaxf(fmaf(y, __ldg(ep.pool_scale + n), __ldg(ep.pool_shift + n)), 0.f);
        }
        z[i] = t;
    }
    float4* dst = reinterpret_cast<float4*>(buf + tid * TC_POOL_ROW);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = make_float4(z[4 * i], z[4 * i + 1], z[4 * i + 2], z[4 * i + 3]);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int Lp = ep.pool_Lp;
    const int n_out = samples_in_tile * Lp * 8;
    for (int o = tid; o < n_out; o += 128) {
        const int cp = o & 7, pj = o >> 3;
        const int g = pj / Lp, j = pj - g * Lp;
        const int sample = sample0 + g, n = n0 + 2 * cp;
        if (sample >= Bn || n >= N) continue;
        const float* src = buf + (g * rows_per_sample + 2 * j) * TC_POOL_ROW + 2 * cp;
        float2 m = *reinterpret_cast<const float2*>(src);
#pragma unroll
        for (int i = 1; i < 10; ++i) {
            const float2 t = *reinterpret_cast<const float2*>(src + i * TC_POOL_ROW);
            m.x = fmaxf(m.x, t.x);
            m.y = fmaxf(m.y, t.y);
        }
        *reinterpret_cast<uint32_t*>((bf16*)ep.pool_out + ((size_t)sample * Lp + j) * ep.pool_ld + n) = pack_bf16x2(m.x, m.y);
    }
}

// Persistent: one CTA per SM walks the tile list; two TMEM accumulator stages let the epilogue of tile i
// overlap the main loop of tile i+1.
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcParams p,
               const Epilogue ep) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int stage_bytes = p.a.stage_bytes + p.b.stage_bytes;
    uint64_t* full_bar = (uint64_t*)(smem + (size_t)p.stages * stage_bytes);
    uint64_t* empty_bar = full_bar + TC_MAX_STAGES;
    uint64_t* tfull_bar = empty_bar + TC_MAX_STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);
    float* epi_scratch = (float*)((uint8_t*)full_bar + 256);      // [4 warps][32][33] transpose tiles of the atomic epilogue

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

    const bool bn_on = ep.bn_stats != nullptr && ep.mode == EPI_LINEAR;       // the transpose scratch holds [4 warps][2][N] accumulators (N <= 512)
    if (bn_on) for (int i = threadIdx.x; i < 8 * p.N; i += TC_THREADS) epi_scratch[i] = 0.f;
    if (p.zero_smem) {
        uint4 z = make_uint4(0, 0, 0, 0);
        for (int i = threadIdx.x; i < p.stages * stage_bytes / 16; i += TC_THREADS) ((uint4*)smem)[i] = z;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tile t -> (tile_n fastest, tile_m, z); z = tap_z * split_k + ks
#define TC_DECODE_TILE(t)                                                                         \
    const int tile_n = (t) % p.grid_n;                                                            \
    const int tile_m = ((t) / p.grid_n) % p.grid_m;                                               \
    const int zz = (t) / (p.grid_n * p.grid_m);                                                   \
    const int tap_z = p.tap_in_z ? zz / p.split_k : 0;                                            \
    const int ks = p.tap_in_z ? zz - tap_z * p.split_k : zz;                                      \
    const int n_inner = p.j_total > 0 ? max(0, min(p.n_inner, p.j_total - ks * p.n_inner)) : p.n_inner; \
    const int n_iters = (p.tap_in_z ? 1 : p.n_taps) * n_inner;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            // The producer is ONE thread: its per-stage instruction count bounds small-tile GEMMs (measured ~900 cycles per
            // stage with coordinates recomputed from the parameter block), so everything loop-invariant lives in
            // registers and TMA coordinates advance by deltas.
            const uint32_t tx = (uint32_t)(p.a.boxes * p.a.box_bytes + p.b.boxes * p.b.box_bytes);
            const int a_boxes = p.a.boxes, b_boxes = p.b.boxes;
            const uint32_t a_block = p.a.block_bytes, b_block = p.b.block_bytes;
            const int acj0 = p.a.cj[0], acj1 = p.a.cj[1], acj2 = p.a.cj[2], act0 = p.a.ctap[0], act1 = p.a.ctap[1], act2 = p.a.ctap[2];
            const int bcj0 = p.b.cj[0], bcj1 = p.b.cj[1], bcj2 = p.b.cj[2], bct0 = p.b.ctap[0], bct1 = p.b.ctap[1], bct2 = p.b.ctap[2];
            const int acb0 = p.a.cblk[0], acb1 = p.a.cblk[1], acb2 = p.a.cblk[2], bcb0 = p.b.cblk[0], bcb1 = p.b.cblk[1], bcb2 = p.b.cblk[2];
            const uint32_t smem0 = smem_u32(smem), full0 = smem_u32(full_bar), a_stage = p.a.stage_bytes;
            const uint32_t a_off = p.a.dst_off, b_off = p.b.dst_off;
            const uint64_t ma = (uint64_t)&map_a, mb = (uint64_t)&map_b;
            const int n_stages = p.stages;
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                TC_DECODE_TILE(t)
                const int tap_first = p.tap_in_z ? tap_z : 0;
                const int j_first = ks * p.n_inner;
                int a0 = p.a.base[0] + tap_first * act0 + j_first * acj0 + tile_m * p.a.ctile[0];
                int a1 = p.a.base[1] + tap_first * act1 + j_first * acj1 + tile_m * p.a.ctile[1];
                int a2 = p.a.base[2] + tap_first * act2 + j_first * acj2 + tile_m * p.a.ctile[2];
                int b0 = p.b.base[0] + tap_first * bct0 + j_first * bcj0 + tile_n * p.b.ctile[0];
                int b1 = p.b.base[1] + tap_first * bct1 + j_first * bcj1 + tile_n * p.b.ctile[1];
                int b2 = p.b.base[2] + tap_first * bct2 + j_first * bcj2 + tile_n * p.b.ctile[2];
                int jj = 0;
                for (int it = 0; it < n_iters; ++it) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    const uint32_t bar = full0 + 8u * stage;
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx) : "memory");
                    const uint32_t sa = smem0 + (uint32_t)stage * (uint32_t)stage_bytes + a_off;
                    const uint32_t sb = smem0 + (uint32_t)stage * (uint32_t)stage_bytes + a_stage + b_off;
                    for (int blk = 0; blk < a_boxes; ++blk)
                        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                                     ::"r"(sa + blk * a_block), "l"(ma), "r"(bar), "r"(a0 + blk * acb0), "r"(a1 + blk * acb1), "r"(a2 + blk * acb2) : "memory");
                    for (int blk = 0; blk < b_boxes; ++blk)
                        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                                     ::"r"(sb + blk * b_block), "l"(mb), "r"(bar), "r"(b0 + blk * bcb0), "r"(b1 + blk * bcb1), "r"(b2 + blk * bcb2) : "memory");
                    if (++jj == n_inner) {          // next tap: rewind the chunk coordinate, step the tap coordinate
                        jj = 0;
                        a0 += act0 - (n_inner - 1) * acj0; a1 += act1 - (n_inner - 1) * acj1; a2 += act2 - (n_inner - 1) * acj2;
                        b0 += bct0 - (n_inner - 1) * bcj0; b1 += bct1 - (n_inner - 1) * bcj1; b2 += bct2 - (n_inner - 1) * bcj2;
                    } else {
                        a0 += acj0; a1 += acj1; a2 += acj2;
                        b0 += bcj0; b1 += bcj1; b2 += bcj2;
                    }
                    if (++stage == n_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // The whole warp walks the loop (uniform control flow keeps the descriptors in uniform registers) and ONE elected
        // lane issues: inside a divergent `if (lane == 0)` every tcgen05.mma cost an ELECT / R2UR / BRA.U.ANY waterfall
        // (~25 SASS instructions, ~150 cycles), which made N <= 128 tiles issue-bound.
        const uint64_t da_desc = umma_desc(0, p.a.lbo_bytes, 1024), db_desc = umma_desc(0, p.b.lbo_bytes, 1024);
        const uint32_t da_hi = (uint32_t)(da_desc >> 32), db_hi = (uint32_t)(db_desc >> 32);
        const uint32_t da_lbo = (uint32_t)da_desc, db_lbo = (uint32_t)db_desc;         // LBO field, bits 16-29
        const uint32_t a_kstep = (uint32_t)p.a.kstep_bytes >> 4, b_kstep = (uint32_t)p.b.kstep_bytes >> 4;
        const uint32_t smem0 = smem_u32(smem), a_stage = p.a.stage_bytes, idesc = p.idesc;
        const int k_steps = p.k_steps, n_stages = p.stages, multi = p.taps_per_cta;
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            TC_DECODE_TILE(t)
            (void)tile_n; (void)tile_m;
            if (n_iters == 0) continue;
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);      // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_stride);
            for (int it = 0; it < n_iters; ++it) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem0 + (uint32_t)stage * (uint32_t)stage_bytes;
                const uint32_t sb = sa + a_stage;
                if (elect_one_sync()) {
                    if (multi > 1) {
                        // all taps of this CTA's group from one staged (dy, x) pair: x is read through row-shifted descriptors
                        const int tap0 = tap_z * multi;
                        const int nt = min(multi, p.n_taps - tap0);
                        const uint64_t da0 = ((uint64_t)da_hi << 32) | (da_lbo | ((sa & 0x3FFFFu) >> 4));
                        if (p.fuse_taps > 1) {
                            for (int tt = 0; tt < nt; tt += p.fuse_taps) {
                                const int g = min(p.fuse_taps, nt - tt);
                                const uint32_t sbt = sb + p.b.dst_off + (uint32_t)((tap0 + tt - p.tap_pad) * 128);
                                const uint64_t db0 = ((uint64_t)db_hi << 32) | ((8u << 16) | ((sbt & 0x3FFFFu) >> 4));        // LBO = 128 bytes
                                const uint32_t idg = (idesc & ~(0x3Fu << 17)) | ((uint32_t)(g * 64 >> 3) << 17);               // N = g * 64
                                const uint32_t d_t = d_tmem + (uint32_t)(tt * p.n_tile);
                                for (int s2 = 0; s2 < k_steps; ++s2)
                                    tc_mma_f16(d_t, da0 + (uint64_t)(s2 * a_kstep), db0 + (uint64_t)(s2 * b_kstep), idg, (it | s2) ? 1u : 0u);
                            }
                        } else
                        for (int tt = 0; tt < nt; ++tt) {
                            const uint32_t sbt = sb + p.b.dst_off + (uint32_t)((tap0 + tt - p.tap_pad) * 128);
                            const uint64_t db0 = ((uint64_t)db_hi << 32) | (db_lbo | ((sbt & 0x3FFFFu) >> 4));
                            const uint32_t d_t = d_tmem + (uint32_t)(tt * p.n_tile);
                            for (int s2 = 0; s2 < k_steps; ++s2)
                                tc_mma_f16(d_t, da0 + (uint64_t)(s2 * a_kstep), db0 + (uint64_t)(s2 * b_kstep), idesc, (it | s2) ? 1u : 0u);
                        }
                    } else {
                        const uint64_t da0 = ((uint64_t)da_hi << 32) | (da_lbo | ((sa & 0x3FFFFu) >> 4));
                        const uint64_t db0 = ((uint64_t)db_hi << 32) | (db_lbo | ((sb & 0x3FFFFu) >> 4));
                        tc_mma_f16(d_tmem, da0, db0, idesc, it ? 1u : 0u);
                        if (k_steps == 4) {
                            tc_mma_f16(d_tmem, da0 + a_kstep, db0 + b_kstep, idesc, 1u);
                            tc_mma_f16(d_tmem, da0 + 2 * a_kstep, db0 + 2 * b_kstep, idesc, 1u);
                            tc_mma_f16(d_tmem, da0 + 3 * a_kstep, db0 + 3 * b_kstep, idesc, 1u);
                        } else {
#pragma unroll 4
                            for (int s2 = 1; s2 < k_steps; ++s2)
                                tc_mma_f16(d_tmem, da0 + (uint64_t)(s2 * a_kstep), db0 + (uint64_t)(s2 * b_kstep), idesc, 1u);
                        }
                    }
                    tc_commit(&empty_bar[stage]);          // frees the smem stage when these MMAs retire
                }
                __syncwarp();
                if (++stage == n_stages) { stage = 0; phase ^= 1; }
            }
            if (elect_one_sync()) tc_commit(&tfull_bar[acc]);                // accumulator complete
            __syncwarp();
            if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ================= epilogue (warps 2..5 own TMEM lane quarters warp%4) =================
        const int q = warp & 3;
        const int r = q * 32 + lane;                                   // tile row == TMEM lane
        const int grp = r / p.rows_per_group;
        const int r_in = r - grp * p.rows_per_group;
        const bool r_ok = r < p.groups_per_tile * p.rows_per_group;
        int acc = 0, pool_chunk = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            TC_DECODE_TILE(t)
            (void)ks;
            if (n_iters == 0) continue;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const int m = (tile_m * p.groups_per_tile + grp) * p.rows_per_group + r_in;
            const bool row_ok = r_ok && (m < p.M);
            const int n_base = tile_n * p.n_tile;
            const int tpc = p.taps_per_cta > 1 ? p.taps_per_cta : 1;
            const int tap0 = p.tap_in_z ? tap_z * tpc : 0;
            const int nt = p.tap_in_z ? min(tpc, p.n_taps - tap0) : 1;
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.acc_stride);
            const bool coalesced_atomics = ep.mode == EPI_ATOMIC && ep.map == MAP_ROWMAJOR && p.rows_per_group == 128;
            for (int t = 0; t < nt; ++t) {
                const int n_off = p.tap_in_z ? (tap0 + t) * p.n_off_per_tap : 0;
                if (coalesced_atomics) {
                    // Split-K / wgrad accumulation.  A thread owns one output ROW of the accumulator, so a direct atomicAdd makes
                    // every warp instruction touch 32 different rows (32 L2 transactions; measured ~45 us per 128 x 128 tile,
                    // independent of K).  The 32 x 32 block is transposed through shared memory instead, so that one
                    // instruction adds 32 consecutive columns of one row: a single 128-byte transaction.
                    float* sc = epi_scratch + q * (32 * 33);
                    const int m0 = tile_m * 128 + q * 32;
                    for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
                        if (n_base + c0 >= p.N) break;                      // warp-uniform
                        float v[32];
                        tc_ld16(t_addr + (uint32_t)(t * p.n_tile + c0), v);
                        if (c0 + 16 < p.n_tile) tc_ld16(t_addr + (uint32_t)(t * p.n_tile + c0 + 16), v + 16);
                        else {
#pragma unroll
                            for (int i = 16; i < 32; ++i) v[i] = 0.f;
                        }
#pragma unroll
                        for (int i = 0; i < 32; ++i) sc[lane * 33 + i] = v[i];
                        __syncwarp();
                        const int n = n_base + c0 + lane;
                        const bool n_ok = n < p.N && c0 + lane < p.n_tile;
                        float* dst = (float*)ep.out + (size_t)m0 * ep.ldo + n_off + n;
                        const int rows = min(32, p.M - m0);
                        for (int rr = 0; rr < rows; ++rr)
                            if (n_ok) atomicAdd(dst + (size_t)rr * ep.ldo, sc[rr * 33 + lane]);
                        __syncwarp();
                    }
                    continue;
                }
                if (ep.mode == EPI_POOL) {
                    for (int c0 = 0; c0 < p.n_tile; c0 += 16, ++pool_chunk) {      // the two staging buffers alternate ACROSS tiles as well
                        if (n_base + c0 >= p.N) break;                      // uniform over the four epilogue warps
                        float v[16];
                        tc_ld16(t_addr + (uint32_t)c0, v);
                        tc_pool_chunk16(ep, epi_scratch + (pool_chunk & 1) * (128 * TC_POOL_ROW), r, v, n_base + c0, p.N, p.rows_per_group,
                                        p.groups_per_tile, tile_m * p.groups_per_tile, p.M / p.rows_per_group);
                    }
                    continue;
                }
                for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
                    if (n_base + c0 >= p.N) break;                          // warp-uniform
                    float v[16];
                    tc_ld16(t_addr + (uint32_t)(t * p.n_tile + c0), v);
                    if (bn_on) tc_bn_stats16(ep, epi_scratch + q * 2 * p.N, p.N, n_base + c0, p.N, row_ok, v, lane);
                    if (row_ok) tc_epilogue_row16(ep, m, n_base + c0, n_off, p.M, p.N, p.n_logical, v);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);   // accumulator may be overwritten
            if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1; }
        }
    }
#undef TC_DECODE_TILE
    tc_fence_before();
    __syncthreads();
    if (bn_on) for (int i = threadIdx.x; i < 2 * p.N; i += TC_THREADS)
        atomicAdd(&ep.bn_stats[i], (double)epi_scratch[i] + (double)epi_scratch[2 * p.N + i] + (double)epi_scratch[4 * p.N + i] + (double)epi_scratch[6 * p.N + i]);
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn& encode_fn() { static EncodeTiledFn f = nullptr; return f; }
inline int& tc_max_smem() { static int v = 0; return v; }
inline int& tc_num_sms() { static int v = 148; return v; }
// conv wgrad tiling (tunable for experiments): taps accumulated per CTA (<=1: one tap per CTA) and the N tile
inline int tc_min_kiters() { return tuning().min_kiters; }
inline int tc_wgrad_taps() { return tuning().wgrad_taps; }   // 0: as many taps as TMEM holds; 1: per-tap kernel
inline int tc_wgrad_ntile() { return tuning().wgrad_ntile; }

// true the first time it is called for (slot, current device): cudaFuncSetAttribute is per device
inline bool first_on_device(int slot) {
    static bool done[8][64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return true;
    if (done[slot][dev]) return false;
    done[slot][dev] = true;
    return true;
}

inline int tc_init() {
    // per device: cudaFuncSetAttribute applies to the current device's instance of the kernel
    static bool done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (encode_fn() && dev >= 0 && dev < 64 && done[dev]) return 0;
    if (!encode_fn()) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (err != cudaSuccess || !fn) return set_error(-3, "cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(err));
        encode_fn() = (EncodeTiledFn)fn;
    }
    int smem = 0;
    cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    tc_max_smem() = smem;
    cudaDeviceGetAttribute(&tc_num_sms(), cudaDevAttrMultiProcessorCount, dev);
    cudaError_t err = cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return set_error(-3, "cudaFuncSetAttribute(tc_gemm_kernel): %s", cudaGetErrorString(err));
    if (dev >= 0 && dev < 64) done[dev] = true;
    return 0;
}

// bf16 tensor [d2][d1][d0] (d0 contiguous), strides in ELEMENTS for d1 and d2
inline int make_map(CUtensorMap* m, const void* ptr, int d0, int d1, int d2, int64_t s1, int64_t s2, int b0, int b1, int b2) {
    cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
    cuuint64_t strides[2] = {(cuuint64_t)s1 * 2, (cuuint64_t)s2 * 2};
    cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2};
    cuuint32_t es[3] = {1, 1, 1};
    if (((uintptr_t)ptr & 15) || (strides[0] & 15) || (strides[1] & 15)) return set_error(-1, "TMA operand not 16-byte aligned");
    CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(-3, "cuTensorMapEncodeTiled failed (%d) dims=%d,%d,%d box=%d,%d,%d", (int)r, d0, d1, d2, b0, b1, b2);
    return 0;
}

enum TcKind { TC_LINEAR_FWD = 0, TC_LINEAR_DGRAD = 1, TC_LINEAR_WGRAD = 2, TC_CONV_FWD = 3, TC_CONV_DGRAD = 4, TC_CONV_WGRAD = 5 };

struct TcProblem {
    int kind;
    const bf16* a; int lda;     // see tc_gemm() for the meaning per kind
    const bf16* b; int ldb;
    int M, N, K;                // linear: out[M,N] = A[M,K] B[N,K]^T (fwd); see per-kind notes
    int B, L, Cin, Cout, taps, pad;   // conv geometry (L = positions of both the conv input and output)
    int wgrad_tap_stride;             // TC_CONV_WGRAD: > 0 -> the target is [taps][Cout][Cin] (coalesced atomics), else dW[Cout][Cin][taps] via ep.map
};

__host__ __device__ inline uint32_t make_idesc_dev(int a_mn, int b_mn, int n_tile) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a_mn ? 1 : 0) << 15) | ((uint32_t)(b_mn ? 1 : 0) << 16) |
           ((uint32_t)(n_tile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

inline uint32_t make_idesc(int a_mn, int b_mn, int n_tile) {
    uint32_t d = 0;
    d |= 1u << 4;                    // C format F32
    d |= 1u << 7;                    // A format BF16
    d |= 1u << 10;                   // B format BF16
    d |= (uint32_t)(a_mn ? 1 : 0) << 15;
    d |= (uint32_t)(b_mn ? 1 : 0) << 16;
    d |= (uint32_t)(n_tile >> 3) << 17;
    d |= (uint32_t)(128 >> 4) << 24; // M = 128
    return d;
}

inline void kmajor_operand(TcOperand& o, int rows) {
    o.mn_major = 0; o.boxes = 1; o.box_bytes = rows * 128; o.block_bytes = round_up(rows * 128, 1024);
    o.stage_bytes = o.block_bytes; o.kstep_bytes = 32; o.lbo_bytes = 16;
}
inline void mnmajor_operand(TcOperand& o, int mn_extent, int k_rows_box, int k_rows_stage) {
    o.mn_major = 1; o.boxes = cdiv(mn_extent, 64); o.box_bytes = k_rows_box * 128; o.block_bytes = k_rows_stage * 128;
    o.stage_bytes = o.boxes * o.block_bytes; o.kstep_bytes = 16 * 128; o.lbo_bytes = o.block_bytes;
}

inline bool tc_shape_ok(int n, int k_elems_contig_a, int k_elems_contig_b) {
    return n >= 16 && (n % 16) == 0 && (k_elems_contig_a % 8) == 0 && (k_elems_contig_b % 8) == 0;
}

// Launch one GEMM on the tensor cores.  Operand conventions (all bf16, leading dimensions in elements):
//  TC_LINEAR_FWD    a = X [M, K] (lda), b = W [N, K] (ldb)                       out[m,n] = sum_k X[m,k] W[n,k]
//  TC_LINEAR_DGRAD  a = G [M, K] (lda; K = out units), b = W [K, N] (ldb; N = in) out[m,n] = sum_k G[m,k] W[k,n]
//  TC_LINEAR_WGRAD  a = G [K, M] (lda; K = batch rows, M = out), b = X [K, N] (ldb) out[m,n] = sum_k G[k,m] X[k,n]
//  TC_CONV_FWD      a = act [B, L, Cin] (lda), b = Wf [taps, Cout, Cin] (ldb)     M = B*L rows, N = Cout
//  TC_CONV_DGRAD    a = dy [B, L, Cout] (lda), b = Wf [taps, Cout, Cin] (ldb)     M = B*L rows, N = Cin
//  TC_CONV_WGRAD    a = dy [B, L, Cout] (lda), b = act [B, L, Cin] (ldb)          M = Cout, N = Cin per tap (logical n = tap*Cin + c)
inline int tc_gemm(const TcProblem& pr, const Epilogue& ep, cudaStream_t st, int* launches = nullptr) {
    int rc = tc_init();
    if (rc) return rc;
    TcParams p = {};
    CUtensorMap ma, mb;
    const int kind = pr.kind;
    const bool conv = kind >= TC_CONV_FWD;
    int n_tile, grid_m, grid_n, grid_z = 1;
    p.split_k = 1;
    p.rows_per_group = 128; p.groups_per_tile = 1;
    if (kind == TC_LINEAR_FWD || kind == TC_CONV_FWD || kind == TC_LINEAR_DGRAD || kind == TC_CONV_DGRAD) {
        const int N = pr.N;
        n_tile = N <= 256 ? round_up(N, 16) : 256;
        // few row tiles (small batch, data-parallel shards): narrower N tiles until the launch has about one tile per SM -- a
        // 1024 x 1024 x 4096 docking GEMM is 32 tiles of 128 x 256 (48 us on 32 SMs) or 128 tiles of 128 x 64
        if (!conv)
            while (n_tile >= 128 && (n_tile % 32) == 0 && cdiv(pr.M, 128) * cdiv(N, n_tile) < tc_num_sms() / 2) n_tile /= 2;
        grid_n = cdiv(N, n_tile);
        p.M = pr.M; p.N = N; p.n_logical = N;
        if (conv) {
            const int bt = std::max(1, 128 / pr.L);
            if (pr.L > 128) return set_error(-5, "conv GEMM: L > 128 not supported");
            p.rows_per_group = pr.L; p.groups_per_tile = bt;
            grid_m = cdiv(pr.B, bt);
            const int Ca = kind == TC_CONV_FWD ? pr.Cin : pr.Cout;    // channels of the A activation
            rc = make_map(&ma, pr.a, Ca, pr.L, pr.B, pr.lda, (int64_t)pr.L * pr.lda, 64, pr.L, bt);
            if (rc) return rc;
            kmajor_operand(p.a, bt * pr.L);
            p.a.block_bytes = p.a.stage_bytes = round_up(128 * 128, 1024);   // UMMA reads 128 rows
            p.a.cj[0] = 64;
            p.a.base[1] = kind == TC_CONV_FWD ? -pr.pad : pr.pad;
            p.a.ctap[1] = kind == TC_CONV_FWD ? 1 : -1;
            p.a.ctile[2] = bt;
            p.n_taps = pr.taps;
            p.n_inner = cdiv(Ca, 64);
            // weights Wf [taps][Cout][Cin]
            if (kind == TC_CONV_FWD) {
                rc = make_map(&mb, pr.b, pr.Cin, pr.Cout, pr.taps, pr.ldb, (int64_t)pr.Cout * pr.ldb, 64, n_tile, 1);
                if (rc) return rc;
                kmajor_operand(p.b, n_tile);
                p.b.cj[0] = 64; p.b.ctile[1] = n_tile; p.b.ctap[2] = 1;
            } else {
                rc = make_map(&mb, pr.b, pr.Cin, pr.Cout, pr.taps, pr.ldb, (int64_t)pr.Cout * pr.ldb, 64, 64, 1);
                if (rc) return rc;
                mnmajor_operand(p.b, n_tile, 64, 64);
                p.b.ctile[0] = n_tile; p.b.cblk[0] = 64; p.b.cj[1] = 64; p.b.ctap[2] = 1;
            }
            p.k_steps = 4;
        } else {
            grid_m = cdiv(pr.M, 128);
            rc = make_map(&ma, pr.a, pr.K, pr.M, 1, pr.lda, (int64_t)pr.M * pr.lda, 64, 128, 1);
            if (rc) return rc;
            kmajor_operand(p.a, 128);
            p.a.cj[0] = 64; p.a.ctile[1] = 128;
            p.n_taps = 1;
            p.n_inner = cdiv(pr.K, 64);
            if (kind == TC_LINEAR_FWD) {
                rc = make_map(&mb, pr.b, pr.K, N, 1, pr.ldb, (int64_t)N * pr.ldb, 64, n_tile, 1);
                if (rc) return rc;
                kmajor_operand(p.b, n_tile);
                p.b.cj[0] = 64; p.b.ctile[1] = n_tile;
            } else {
                rc = make_map(&mb, pr.b, N, pr.K, 1, pr.ldb, (int64_t)pr.K * pr.ldb, 64, 64, 1);
                if (rc) return rc;
                mnmajor_operand(p.b, n_tile, 64, 64);
                p.b.ctile[0] = n_tile; p.b.cblk[0] = 64; p.b.cj[1] = 64;
            }
            p.k_steps = 4;
        }
    } else {
        // wgrad: both operands MN-major, K = rows of the batch
        const int N = pr.N;                       // Cin (conv) or in-features (linear)
        n_tile = N <= 128 ? round_up(N, 16) : 128;
        grid_n = cdiv(N, n_tile);
        grid_m = cdiv(pr.M, 128);
        p.M = pr.M; p.N = N;
        int rows_box, j_total, conv_tap_groups = 1;
        (void)rows_box;
        if (conv && (tc_wgrad_taps() == 1 || pr.taps < 2 || pr.pad > 7)) {
            if (pr.L > 128) return set_error(-5, "conv GEMM: L > 128 not supported");
            const int bt = std::max(1, 128 / pr.L);
            rows_box = bt * pr.L;
            j_total = cdiv(pr.B, bt);
            rc = make_map(&ma, pr.a, pr.Cout, pr.L, pr.B, pr.lda, (int64_t)pr.L * pr.lda, 64, pr.L, bt);
            if (rc) return rc;
            rc = make_map(&mb, pr.b, pr.Cin, pr.L, pr.B, pr.ldb, (int64_t)pr.L * pr.ldb, 64, pr.L, bt);
            if (rc) return rc;
            mnmajor_operand(p.a, 128, rows_box, 128);
            mnmajor_operand(p.b, n_tile, rows_box, 128);
            p.a.ctile[0] = 128; p.a.cblk[0] = 64; p.a.cj[2] = bt;
            p.b.ctile[0] = n_tile; p.b.cblk[0] = 64; p.b.cj[2] = bt; p.b.base[1] = -pr.pad; p.b.ctap[1] = 1;
            p.tap_in_z = 1; p.n_taps = pr.taps; p.n_off_per_tap = pr.wgrad_tap_stride > 0 ? pr.wgrad_tap_stride : pr.Cin;
            p.n_logical = pr.taps * pr.Cin;
            p.zero_smem = rows_box < 128;
            p.k_steps = 8;
            conv_tap_groups = pr.taps;
        } else if (conv) {
            // Multi-tap wgrad: one CTA = (128 output channels) x (n_tile input channels) x (a group of T taps, T * n_tile <= 512
            // TMEM columns).  The per-tap kernel above is bound by L2 -> SM operand traffic (it re-fetches dy and x for each
            // of the 15 taps); here ONE staged (dy, x) pair feeds T taps.  K rows of a stage, S = L + pad per sample:
            //   dy_s[g*S + l]       = dy[g, l]   (box starts at l = 0; rows L..S-1 are out of bounds -> zero-filled)
            //   x_s [g*S + pad + l] = x[g, l]    (box starts at l = -pad; the zero halo is SHARED by consecutive samples)
            // so tap t is dy_s against x_s shifted by t rows: one descriptor offset, no reload.
            const int S = pr.L + pr.pad;
            if (S > 240) return set_error(-5, "conv wgrad: L + pad > 240 not supported");
            const int bt = std::max(1, 144 / S);              // whole samples per K block
            const int R = round_up(bt * S, 16);               // K rows per stage (rows beyond the box stay zero)
            rows_box = bt * S;
            j_total = cdiv(pr.B, bt);
            const int nt_max = tc_wgrad_ntile();
            n_tile = N <= nt_max ? round_up(N, 16) : nt_max;
            grid_n = cdiv(N, n_tile);
            rc = make_map(&ma, pr.a, pr.Cout, pr.L, pr.B, pr.lda, (int64_t)pr.L * pr.lda, 64, S, bt);
            if (rc) return rc;
            rc = make_map(&mb, pr.b, pr.Cin, pr.L, pr.B, pr.ldb, (int64_t)pr.L * pr.ldb, 64, S, bt);
            if (rc) return rc;
            mnmajor_operand(p.a, 128, rows_box, R);
            mnmajor_operand(p.b, n_tile, rows_box, R + 16);   // tap t reads rows t .. t + R - 1 (t <= 2 * pad <= 14)
            p.a.ctile[0] = 128; p.a.cblk[0] = 64; p.a.cj[2] = bt; p.a.base[1] = 0;
            p.b.ctile[0] = n_tile; p.b.cblk[0] = 64; p.b.cj[2] = bt; p.b.base[1] = -pr.pad;
            const int want = tc_wgrad_taps() > 1 ? tc_wgrad_taps() : 512 / n_tile;
            p.taps_per_cta = std::max(1, std::min(std::min(pr.taps, want), 512 / n_tile));
            p.tap_pad = 0;
            p.fuse_taps = (n_tile == 64 && p.b.boxes == 1 && tuning().wgrad_fuse_taps) ? std::min(4, p.taps_per_cta) : 1;
            p.tap_in_z = 1; p.n_taps = pr.taps; p.n_off_per_tap = pr.wgrad_tap_stride > 0 ? pr.wgrad_tap_stride : pr.Cin;
            p.n_logical = pr.taps * pr.Cin;
            p.zero_smem = 1;
            p.k_steps = R / 16;
            conv_tap_groups = cdiv(pr.taps, p.taps_per_cta);
        } else {
            rows_box = 128;
            j_total = cdiv(pr.K, 128);
            rc = make_map(&ma, pr.a, pr.M, pr.K, 1, pr.lda, (int64_t)pr.K * pr.lda, 64, 128, 1);
            if (rc) return rc;
            rc = make_map(&mb, pr.b, N, pr.K, 1, pr.ldb, (int64_t)pr.K * pr.ldb, 64, 128, 1);
            if (rc) return rc;
            mnmajor_operand(p.a, 128, 128, 128);
            mnmajor_operand(p.b, n_tile, 128, 128);
            p.a.ctile[0] = 128; p.a.cblk[0] = 64; p.a.cj[1] = 128;
            p.b.ctile[0] = n_tile; p.b.cblk[0] = 64; p.b.cj[1] = 128;
            p.n_taps = 1; p.n_logical = N;
        }
        if (!conv) p.k_steps = 8;
        p.j_total = j_total;
        const int tiles = grid_m * grid_n * (conv ? conv_tap_groups : 1);
        // split-K: every split adds a full tile of fp32 atomics, so keep at least tc_min_kiters() K blocks per CTA
        int split = std::max(1, std::min(std::max(1, j_total / tc_min_kiters()), tc_num_sms() / std::max(1, tiles)));
        if (tuning().deterministic) split = 1;      // one CTA per output tile: the fp32 accumulation has a fixed order
        p.n_inner = cdiv(j_total, split);
        split = cdiv(j_total, p.n_inner);
        p.split_k = split;
        grid_z = split * (conv ? conv_tap_groups : 1);
    }
    p.n_tile = n_tile;
    p.idesc = make_idesc(p.a.mn_major, p.b.mn_major, n_tile);
    p.acc_stride = n_tile <= 16 ? 16 : n_tile <= 32 ? 32 : n_tile <= 64 ? 64 : n_tile <= 128 ? 128 : 256;
    p.acc_stages = 2;
    if (p.taps_per_cta > 1) {
        const int cols = p.taps_per_cta * n_tile;
        p.acc_stride = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
        p.acc_stages = p.acc_stride <= 256 ? 2 : 1;
    }
    p.tmem_cols = std::max(32, p.acc_stages * p.acc_stride);
    p.grid_m = grid_m; p.grid_n = grid_n; p.grid_z = grid_z;
    p.total_tiles = grid_m * grid_n * grid_z;
    const int stage_bytes = p.a.stage_bytes + p.b.stage_bytes;
    const int epi_scratch_bytes = std::max(4 * 32 * 33 * 4, ep.mode == EPI_POOL ? TC_POOL_SCRATCH_BYTES : 0);
    const int budget = tc_max_smem() - 2048 - epi_scratch_bytes;
    p.stages = std::min(TC_MAX_STAGES, budget / stage_bytes);
    if (p.stages < 2) return set_error(-5, "tc_gemm: stage of %d bytes does not fit twice in shared memory", stage_bytes);
    const size_t smem = (size_t)p.stages * stage_bytes + 1024 + 256 + epi_scratch_bytes;
    const int grid = std::min(p.total_tiles, tc_num_sms());
    tc_gemm_kernel<<<grid, TC_THREADS, smem, st>>>(ma, mb, p, ep);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return set_error(-3, "tc_gemm launch failed: %s", cudaGetErrorString(err));
    if (launches) ++*launches;
    return 0;
}

}  // namespace emb
