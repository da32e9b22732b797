// Inference conv layer in ONE kernel:  a = MaxPool1d(10, 2)( ReLU( BatchNorm_eval( Conv1d(x) ) ) )      (CNN_pre.py:39-49 in eval mode)
//
// The training-shaped forward writes the pre-pooling conv output y (bf16, [B, L, Cout]) and reads it back in the pooling
// kernel; at inference that round trip is most of the time (arch S, 65 536 regions: 6.2 of 10 GB of traffic).  Fusing the
// pooling into the epilogue of the row-major conv GEMM (rows = positions) was measured 2.3x SLOWER than not fusing: pooling
// runs ACROSS rows there, i.e. across threads, through shared memory and named barriers, on four epilogue warps.
//
// This kernel computes the TRANSPOSED product instead,
//     D[cout, position] = sum_{tap, cin} W[tap][cout][cin] * x[position + tap - pad][cin]
// A = the weights (K-major, 128 cout rows, resident in shared memory for every tap), B = the activation tile (K-major, the
// same staged tile with its zero halo that conv_tc.cuh uses; tap t is the tile read through a descriptor advanced by t rows),
// so a TMEM LANE is an output channel and the TMEM COLUMNS are the positions of whole samples.  An epilogue thread then owns
// one channel: BatchNorm is two registers, and the max-pool window slides along the thread's own registers -- no exchange
// between threads at all.  Eight epilogue warps (two per TMEM lane quarter, each taking half of the tile's pooled rows) hide
// the latency of that sequential walk; each pooled row leaves as 32 channels x 2 bytes per warp, channels-last, which is the
// next layer's input layout.  Arithmetic and rounding points equal the unfused forward (bf16 conv output, fp32 BatchNorm,
// one bf16 rounding of the pooled value) except that the conv bias is folded into the BatchNorm shift, i.e. the conv output is
// not rounded to bf16 on the way: one rounding FEWER than the unfused forward.
//
//   warp 0 TMA producer      warp 1 MMA issuer      warps 2-9 epilogue (TMEM lane quarter = warp % 4)
#pragma once
#include "gemm_tc.cuh"

namespace emb {

// The pooling walk of ONE epilogue thread (= one output channel) over TMEM columns [c_first, c_end) of its lane: columns are
// positions, two per pair; per pair two fma (BatchNorm with the conv bias folded into the shift) and one max; the window of pooled
// row j = max(pair maxima j .. j + 4, 0) (ReLU once per pooled element, not per position).  The four ring slots are named
// registers (a chunk is 8 pairs, a multiple of 4): no moves.  `k` is the pooled row the first pair completes (negative while the
// window is still filling), rows 0 <= k < n_eff are stored at o[k * LD]; `o` arrives pointing at row k.
// A chunk whose eight rows are all stored takes the FAST form: no predicates, and with the row stride LD a template constant
// the eight stores are immediate offsets from one pointer (the first version spent two thirds of its instructions on a 64-bit
// pointer add, a counter and a range test per pair).  Up to four 16-column TMEM loads are in flight before the single wait.
// experiment (EMB_CONV_DEBUG & 8): non-blocking test_wait spin instead of the potentially suspending try_wait
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
#define POOL_WAIT(bar, parity) do { if (p.dbg & 8) mbar_wait_spin(bar, parity); else mbar_wait(bar, parity); } while (0)

template <int LD>
__device__ __forceinline__ void pool_walk(uint32_t t_addr, int c_first, int c_end, int k, unsigned n_eff, bf16* o, long long ld_rt,
                                          float sc, float shb) {
    const long long ld = LD ? (long long)LD : ld_rt;
    float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
#define EMB_POOL_VALUE(V, PP)                                                                                 \
        const float pm = fmaxf(fmaf(V[2 * (PP)], sc, shb), fmaf(V[2 * (PP) + 1], sc, shb));                   \
        const bf16 rv = __float2bfloat16_rn(fmaxf(fmaxf(fmaxf(w0, w1), fmaxf(w2, w3)), fmaxf(pm, 0.f)));
#define EMB_POOL_FAST(V, PP, SLOT) { EMB_POOL_VALUE(V, PP) o[(PP) * ld] = rv; SLOT = pm; }
#define EMB_POOL_SLOW(V, PP, SLOT) { EMB_POOL_VALUE(V, PP) if ((unsigned)(k + (PP)) < n_eff) o[(PP) * ld] = rv; SLOT = pm; }
#define EMB_POOL_CHUNK(V)                                                                                     \
        if (k >= 0 && k + 7 < (int)n_eff) {                                                                   \
            EMB_POOL_FAST(V, 0, w0) EMB_POOL_FAST(V, 1, w1) EMB_POOL_FAST(V, 2, w2) EMB_POOL_FAST(V, 3, w3)   \
            EMB_POOL_FAST(V, 4, w0) EMB_POOL_FAST(V, 5, w1) EMB_POOL_FAST(V, 6, w2) EMB_POOL_FAST(V, 7, w3)   \
        } else {                                                                                              \
            EMB_POOL_SLOW(V, 0, w0) EMB_POOL_SLOW(V, 1, w1) EMB_POOL_SLOW(V, 2, w2) EMB_POOL_SLOW(V, 3, w3)   \
            EMB_POOL_SLOW(V, 4, w0) EMB_POOL_SLOW(V, 5, w1) EMB_POOL_SLOW(V, 6, w2) EMB_POOL_SLOW(V, 7, w3)   \
        }                                                                                                     \
        k += 8;                                                                                               \
        o += 8 * ld;
    for (int c16 = c_first; c16 < c_end; c16 += 64) {
        float va[16], vb[16], vc[16], vd[16];
        const bool hb = c16 + 16 < c_end, hc = c16 + 32 < c_end, hd = c16 + 48 < c_end;
        tc_ld16_nowait(t_addr + (uint32_t)c16, va);
        if (hb) tc_ld16_nowait(t_addr + (uint32_t)(c16 + 16), vb);
        if (hc) tc_ld16_nowait(t_addr + (uint32_t)(c16 + 32), vc);
        if (hd) tc_ld16_nowait(t_addr + (uint32_t)(c16 + 48), vd);
        tc_ld_wait();
        { EMB_POOL_CHUNK(va) }
        if (hb) { EMB_POOL_CHUNK(vb) }
        if (hc) { EMB_POOL_CHUNK(vc) }
        if (hd) { EMB_POOL_CHUNK(vd) }
    }
#undef EMB_POOL_CHUNK
#undef EMB_POOL_SLOW
#undef EMB_POOL_FAST
#undef EMB_POOL_VALUE
}

struct TcPoolParams {
    int Bn, L, S, bt, Lp;            // samples, conv positions, staged rows per sample (L + pad), samples per tile, pooled length
    int taps, pad, Cout;
    int parts;                       // epilogue warps per lane quarter that share one tile's pooled rows (1, 2 or 4)
    int n_chunks, k_steps_last;      // K chunks of 64 input channels; UMMA K steps of the last one
    int total_tiles;
    int a_slot_bytes, a_box_bytes;   // activation slot ([rows][64 ch], 128B swizzle) and the bytes one TMA box deposits
    int a_slots;                     // activation slots in flight: tiles are small (one or two boxes), so the prefetch must be several tiles deep
    int w_bytes;                     // all weight blocks + one slack block for the last MMA's over-read, rounded to 1 KB
    int w_slot_bytes;                // one (chunk, tap) weight block: Cout rows x 128 B.  The MMA reads 128 rows from there: for Cout < 128 the
                                     // rows behind are the next block's (or the barriers') bytes -- they only feed TMEM lanes >= Cout, which nobody reads
    uint32_t idesc;
    int ld_out;
    int dbg;                         // EMB_CONV_DEBUG bisection: 1 = epilogue does not pool, 8 = test_wait spin instead of try_wait
};

constexpr int TCP_EPI_WARPS = 16;            // sets x parts x 4 lane quarters
constexpr int TCP_ACCS = 4;                  // TMEM accumulators of 128 columns in flight.  An accumulator hand-over (tcgen05.commit -> epilogue
                                             // wake-up -> arrive -> MMA-warp wake-up) was measured at ~1 us with nothing else to do (r2 bisection,
                                             // profiles/r02_infer_bisect.txt): with two accumulators that latency, not the MMAs or the pooling, set the
                                             // tile period
constexpr int TCP_THREADS = 64 + TCP_EPI_WARPS * 32;
constexpr int TCP_MAX_SLOTS = 8;

// LD = Cout (the output row stride and the rows of one weight block) and TAPS = the kernel size when they are the search space's
// values (CNN_pre.py:22-27), 0 = run-time.  The specialisation is for the MMA-issuing thread as much as for the epilogue: r2's
// bisection (profiles/r02_infer_bisect.txt) showed the kernel taking 250 us with the MMAs AND the pooling switched off -- the
// time of the one elected thread walking its tap loop (constant-bank reloads, descriptor multiplies) at single-thread latency.
// With both constants the loop unrolls into descriptor adds with immediates, three instructions per tcgen05.mma.
template <int LD, int TAPS>
__global__ void __launch_bounds__(TCP_THREADS, 1)
tc_conv_pool_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const TcPoolParams p,
                    const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift, bf16* __restrict__ out,
                    bf16* __restrict__ scratch) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_w = smem + p.a_slots * p.a_slot_bytes;
    uint64_t* bars = (uint64_t*)(smem_w + p.w_bytes);
    uint64_t* a_full = bars;                        // [8]
    uint64_t* a_empty = a_full + TCP_MAX_SLOTS;      // [8]
    uint64_t* w_full = a_empty + TCP_MAX_SLOTS;      // [1]
    uint64_t* tfull = w_full + 1;        // [TCP_ACCS]
    uint64_t* tempty = tfull + TCP_ACCS; // [TCP_ACCS]
    uint32_t* tmem_slot = (uint32_t*)(tempty + TCP_ACCS);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

    {   // halo / tail rows of the activation slots must read as zero
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (int i = threadIdx.x; i < p.a_slots * p.a_slot_bytes / 16; i += TCP_THREADS) ((uint4*)smem)[i] = z;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < TCP_MAX_SLOTS; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < TCP_ACCS; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4 * p.parts); }
        mbar_init(w_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TCP_ACCS * 128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            const uint64_t mx = (uint64_t)&map_x, mw = (uint64_t)&map_w;
            const uint32_t sa0 = smem_u32(smem), sw0 = smem_u32(smem_w);
            {   // every tap of the weights, once: box {64 cin, Cout rows, 1 tap}
                const uint32_t bar = smem_u32(w_full);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(p.taps * p.n_chunks * p.w_slot_bytes)) : "memory");
                for (int c = 0; c < p.n_chunks; ++c)
                    for (int t = 0; t < p.taps; ++t)
                        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                                     ::"r"(sw0 + (uint32_t)((c * p.taps + t) * p.w_slot_bytes)), "l"(mw), "r"(bar), "r"(c * 64), "r"(0), "r"(t) : "memory");
            }
            int as = 0;
            uint32_t aphase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x)
                for (int c = 0; c < p.n_chunks; ++c) {
                    POOL_WAIT(&a_empty[as], aphase ^ 1);
                    const uint32_t abar = smem_u32(&a_full[as]);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(abar), "r"((uint32_t)p.a_box_bytes) : "memory");
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                                 ::"r"(sa0 + (uint32_t)(as * p.a_slot_bytes)), "l"(mx), "r"(abar), "r"(c * 64), "r"(-p.pad), "r"(tile * p.bt) : "memory");
                    if (++as == p.a_slots) { as = 0; aphase ^= 1; }
                }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: D[cout][position] += W[tap][cout][:] . x[position + tap][:] =================
        const uint64_t dproto = umma_desc(0, 16, 1024);
        const uint32_t d_hi = (uint32_t)(dproto >> 32), d_lo16 = (uint32_t)dproto;
        const uint32_t sa0 = smem_u32(smem), sw0 = smem_u32(smem_w), idesc = p.idesc;
        int as = 0, acc = 0;
        uint32_t aphase = 0, acc_phase = 0;
        POOL_WAIT(w_full, 0);
        tc_fence_after();
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            POOL_WAIT(&tempty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 128);
            uint32_t accumulate = 0;
            for (int c = 0; c < p.n_chunks; ++c) {
                POOL_WAIT(&a_full[as], aphase);
                tc_fence_after();
                const int ks = (c == p.n_chunks - 1) ? p.k_steps_last : 4;
                if (elect_one_sync()) {
                    uint64_t dx = ((uint64_t)d_hi << 32) | (d_lo16 | (((sa0 + (uint32_t)(as * p.a_slot_bytes)) & 0x3FFFFu) >> 4));
                    uint64_t dw = ((uint64_t)d_hi << 32) | (d_lo16 | (((sw0 + (uint32_t)(c * p.taps) * (uint32_t)p.w_slot_bytes) & 0x3FFFFu) >> 4));
                    if (TAPS) {
#pragma unroll
                        for (int t = 0; t < TAPS; ++t)
#pragma unroll
                            for (int s2 = 0; s2 < 4; ++s2)
                                if (s2 < ks)
                                    tc_mma_f16(d_tmem, dw + (uint64_t)(t * (LD * 8) + 2 * s2), dx + (uint64_t)(8 * t + 2 * s2), idesc,
                                               (t | s2) ? 1u : accumulate);
                    } else {
#pragma unroll 1
                        for (int t = 0; t < p.taps; ++t) {
                            for (int s2 = 0; s2 < ks; ++s2) {
                                tc_mma_f16(d_tmem, dw + (uint64_t)(2 * s2), dx + (uint64_t)(2 * s2), idesc, accumulate);
                                accumulate = 1;
                            }
                            dx += 8;                               // next tap: the activation tile one row (128 bytes) further
                            dw += (uint64_t)(p.w_slot_bytes >> 4);
                        }
                    }
                    tc_commit(&a_empty[as]);
                }
                __syncwarp();
                accumulate = 1;
                if (++as == p.a_slots) { as = 0; aphase ^= 1; }
            }
            if (elect_one_sync()) tc_commit(&tfull[acc]);
            __syncwarp();
            if (++acc == TCP_ACCS) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ================= epilogue: thread = output channel; its registers walk the positions =================
        // (pool_walk above.)  S is even, so every sample starts on an even column and pairs never straddle samples; pairs that
        // precede a segment only pass through the ring.
        const int q = warp & 3, sub = (warp - 2) >> 2;                    // sub 0 .. TCP_EPI_WARPS / 4 - 1
        const int sets = (TCP_EPI_WARPS / 4) / p.parts;                   // warp sets take tiles round-robin; `parts` warps per quarter share a tile's rows
        const int set = sub % sets, part = sub / sets;
        const int ch = q * 32 + lane;
        const bool ch_ok = ch < p.Cout;
        const float sc = ch_ok ? scale[ch] : 0.f;
        const float shb = ch_ok ? fmaf(bias[ch], sc, shift[ch]) : 0.f;
        (void)scratch;
        const bool quarter_live = q * 32 < p.Cout;                        // a whole warp beyond Cout only keeps the barrier protocol going
        for (int t = set, tile = blockIdx.x + set * gridDim.x; tile < p.total_tiles; t += sets, tile += sets * gridDim.x) {
            const int acc = t % TCP_ACCS;
            POOL_WAIT(&tfull[acc], (uint32_t)(t / TCP_ACCS) & 1u);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 128);
            const int sample0 = tile * p.bt;
            const int n_samples = min(p.bt, p.Bn - sample0);
            const int PR = n_samples * p.Lp;                              // pooled rows of this tile, flattened (sample, j)
            const int per = (PR + p.parts - 1) / p.parts;
            const int pr_lo = min(PR, part * per), pr_hi = min(PR, pr_lo + per);
            int pr = quarter_live ? pr_lo : pr_hi;
            while (pr < pr_hi) {
                const int g = pr / p.Lp, j_lo = pr - g * p.Lp;
                const int n = min(p.Lp - j_lo, pr_hi - pr);               // this segment: pooled rows j_lo .. j_lo + n - 1 of sample g
                const int P_first = ((g * p.S) >> 1) + j_lo + 4, P_end = P_first + n;      // the pair that completes window j is pair j + 4
                const int c_first = (2 * (P_first - 4)) & ~15;
                const int k = (c_first >> 1) - P_first;                   // the pooled row the first pair completes (negative: still filling)
                bf16* o = out + ((long long)(sample0 + g) * p.Lp + j_lo + k) * (long long)p.ld_out + ch;
                if (!(p.dbg & 1)) pool_walk<LD>(t_addr, c_first, 2 * P_end, k, ch_ok ? (unsigned)n : 0u, o, p.ld_out, sc, shb);
                pr += n;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCP_ACCS * 128) : "memory");
    }
}

inline size_t tc_conv_pool_wbytes(int taps, int Cin, int Cout) { return round_up64((int64_t)taps * cdiv(Cin, 64) * Cout * 128 + 128 * 128, 1024); }
inline size_t tc_conv_pool_smem(int pad, int taps, int Cin, int Cout, int slots) {
    const int a_rows = round_up(128 + 2 * pad, 8);
    return (size_t)slots * a_rows * 128 + tc_conv_pool_wbytes(taps, Cin, Cout) + 1024 + 512;
}

// x [B, L, Cin] bf16 (ldx), w = Wf [taps][Cout][ldw] bf16, out [B, Lp, ld_out] bf16
inline bool tc_conv_pool_ok(int L, int pad, int taps, int Cin, int Cout, int Lp, int ldx, int ld_out) {
    if (L > 128 || taps < 1 || taps > 15 || Cin < 16 || (Cin % 8) || Cout < 8 || Cout > 128 || (Cout % 8) || Lp < 1) return false;
    if (ldx != Cin || ld_out != Cout) return false;
    return tc_conv_pool_smem(pad, taps, Cin, Cout, 2) <= (size_t)tc_max_smem();
}

// scratch: at least 148 * TCP_THREADS bf16 elements nobody else uses during the launch (the unused pre-pooling buffer of the layer)
inline int tc_conv_pool(const bf16* x, const bf16* w, int ldw, const float* bias, const float* scale, const float* shift, bf16* out, bf16* scratch,
                        int B, int L, int Lp, int Cin, int Cout, int taps, int pad, cudaStream_t st) {
    int rc = tc_init();
    if (rc) return rc;
    using Kern = void (*)(const CUtensorMap, const CUtensorMap, const TcPoolParams, const float*, const float*, const float*, bf16*, bf16*);
#define EMB_TCP_ROW(LD) {tc_conv_pool_kernel<LD, 5>, tc_conv_pool_kernel<LD, 11>, tc_conv_pool_kernel<LD, 15>}
    static const Kern kerns[4][3] = {EMB_TCP_ROW(32), EMB_TCP_ROW(64), EMB_TCP_ROW(96), EMB_TCP_ROW(128)};
#undef EMB_TCP_ROW
    static const Kern generic = tc_conv_pool_kernel<0, 0>;
    if (first_on_device(3)) {
        for (int i = 0; i <= 12; ++i) {
            cudaError_t err = cudaFuncSetAttribute(i < 12 ? kerns[i / 3][i % 3] : generic, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_max_smem());
            if (err != cudaSuccess) return set_error(-3, "cudaFuncSetAttribute(tc_conv_pool_kernel): %s", cudaGetErrorString(err));
        }
    }
    TcPoolParams p = {};
    p.Bn = B; p.L = L; p.S = round_up(L + pad, 2);          // even: every sample starts on an even TMEM column
    p.bt = 1 + (128 - L) / p.S; p.Lp = Lp; p.taps = taps; p.pad = pad; p.Cout = Cout;
    p.n_chunks = cdiv(Cin, 64);
    p.k_steps_last = cdiv(Cin - 64 * (p.n_chunks - 1), 16);
    p.total_tiles = cdiv(B, p.bt);
    const int a_rows = round_up(128 + 2 * pad, 8);
    p.a_slot_bytes = a_rows * 128;
    p.a_box_bytes = p.bt * p.S * 128;
    if (p.bt * p.S > a_rows) return set_error(-5, "tc_conv_pool: tile rows exceed the activation slot");
    p.idesc = make_idesc(0, 0, 128);
    p.ld_out = Cout;
    p.dbg = tuning().conv_debug;
    p.parts = 1;                                 // measured (profiles/r02_infer_parts.sh): whole tiles per warp beat shared ones
    if (tuning().pool_parts == 1 || tuning().pool_parts == 2 || tuning().pool_parts == 4) p.parts = tuning().pool_parts;
    CUtensorMap mx, mw;
    rc = make_map(&mx, x, Cin, L, B, Cin, (int64_t)L * Cin, 64, p.S, p.bt);
    if (rc) return rc;
    rc = make_map(&mw, w, Cin, Cout, taps, ldw, (int64_t)Cout * ldw, 64, Cout, 1);
    if (rc) return rc;
    p.w_slot_bytes = Cout * 128;
    p.w_bytes = (int)tc_conv_pool_wbytes(taps, Cin, Cout);
    p.a_slots = 2;
    while (p.a_slots < TCP_MAX_SLOTS && tc_conv_pool_smem(pad, taps, Cin, Cout, p.a_slots + 1) <= (size_t)tc_max_smem()) ++p.a_slots;
    const size_t smem = tc_conv_pool_smem(pad, taps, Cin, Cout, p.a_slots);
    const int grid = std::min(p.total_tiles, tc_num_sms());
    const int ti = taps == 5 ? 0 : taps == 11 ? 1 : taps == 15 ? 2 : -1;
    const Kern kf = ((Cout % 32) == 0 && Cout <= 128 && ti >= 0) ? kerns[Cout / 32 - 1][ti] : generic;
    kf<<<grid, TCP_THREADS, smem, st>>>(mx, mw, p, bias, scale, shift, out, scratch);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return set_error(-3, "tc_conv_pool launch failed: %s", cudaGetErrorString(err));
    return 0;
}

}  // namespace emb
