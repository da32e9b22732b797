// Conv1d (layers >= 1) forward / dgrad implicit GEMM with TAP REUSE, tcgen05 / TMEM / TMA, sm_100a.
//
// The per-tap kernel in gemm_tc.cuh re-fetches the activation tile once per tap: 15 taps read (almost) the same
// 128 rows 15 times, and with narrow layers (N = 64..96) the GEMM is bound by L2 -> SM operand traffic, not by the
// tensor pipe.  Here the activation tile of a K chunk (64 channels) is staged ONCE, with its zero halo, and tap t
// is the same shared-memory tile read through a UMMA descriptor whose start address is advanced by t rows of
// 128 bytes (the 128B swizzle is a function of the absolute shared-memory address, so a row-shifted start
// address addresses the shifted rows correctly; csrc/probe.cuh checks this on the hardware).
//
//   smem A slot (K-major, SWIZZLE_128B):  [pad zero rows][sample 0: L rows][pad zero rows][sample 1: L rows] ... [zeros]
//       one TMA box {64 ch, L + pad, bt samples} starting at l = -pad deposits it; the l < 0 rows are zero-filled by
//       the TMA unit, the rows behind the box are zeroed once and never written.  Consecutive samples SHARE a halo.
//   D row r  <->  sample r / S, position r % S   (S = L + pad; rows with r % S >= L are discarded)
//   fwd  : D[r] += A[r + t]          * W[t]       dgrad: D[r] += A[r + 2*pad - t] * W[t]^T
//
// Weights: streamed per (chunk, tap) through a ring of small stages, or -- when all taps fit next to two A slots
// (conv1 of arch L: 15 x 96 x 64 bf16 = 180 KB) -- loaded once per CTA and kept resident for every tile.
//
//   warp 0     TMA producer      warp 1     MMA issuer      warps 2-5  epilogue (shared with gemm_tc.cuh)
#pragma once
#include "gemm_tc.cuh"

namespace emb {

struct TcConvParams {
    int M, N;                 // output rows (B * L) and columns
    int Bn, L, S, bt;         // samples, positions, smem rows per sample (L + pad), samples per tile
    int taps, pad, dgrad;
    int n_chunks, k_steps_last;
    int n_tile, grid_m, grid_n, total_tiles;
    int a_slot_bytes, a_box_bytes;
    int b_slot_bytes, b_boxes, b_box_bytes, b_stages, b_resident;
    int tps;                  // streamed weights: taps per pipeline stage (one barrier round trip per `tps` taps)
    int b_tail, b_tail_slot_bytes, b_tail_box_bytes;   // dgrad, resident: the last K chunk has < 64 rows and uses its own (smaller) box
    uint32_t b_tail_lbo_bytes;
    uint32_t b_kstep16, b_lbo_bytes;     // descriptor advance per UMMA K step (in 16-byte units), LBO
    uint32_t idesc, tmem_cols;
    int acc_stride;
    int debug;                // experiments: 1 skip MMAs, 2 skip epilogue body, 4 skip stores only
};

constexpr int TCV_A_SLOTS = 2;
constexpr int TCV_MAX_B_STAGES = 8;

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv_reuse_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const __grid_constant__ CUtensorMap map_b2, const TcConvParams p, const Epilogue ep) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int n_bslots = p.b_resident ? p.taps * p.n_chunks : p.b_stages * p.tps;
    uint8_t* smem_b = smem + TCV_A_SLOTS * p.a_slot_bytes;
    const uint32_t b_tail_base = (uint32_t)((p.n_chunks - 1) * p.taps) * (uint32_t)p.b_slot_bytes;     // resident tail slots start here
    const size_t b_total = p.b_resident && p.b_tail ? (size_t)b_tail_base + (size_t)p.taps * p.b_tail_slot_bytes : (size_t)n_bslots * p.b_slot_bytes;
    uint64_t* bars = (uint64_t*)(smem_b + b_total);
    uint64_t* a_full = bars;                       // [2]
    uint64_t* a_empty = a_full + TCV_A_SLOTS;      // [2]
    uint64_t* b_full = a_empty + TCV_A_SLOTS;      // [8]  (resident: b_full[0] covers every weight box)
    uint64_t* b_empty = b_full + TCV_MAX_B_STAGES; // [8]
    uint64_t* tfull_bar = b_empty + TCV_MAX_B_STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);
    float* s_stats = (float*)((uint8_t*)bars + 512);           // [2][N] BatchNorm accumulators (only with ep.bn_stats)
    const bool bn_on = ep.bn_stats != nullptr && ep.mode == EPI_LINEAR;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

    if (bn_on) for (int i = threadIdx.x; i < 8 * p.N; i += TC_THREADS) s_stats[i] = 0.f;
    {   // halo / tail rows of the A slots must read as zero
        uint4 z = make_uint4(0, 0, 0, 0);
        for (int i = threadIdx.x; i < TCV_A_SLOTS * p.a_slot_bytes / 16; i += TC_THREADS) ((uint4*)smem)[i] = z;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < TCV_A_SLOTS; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < TCV_MAX_B_STAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
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

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            const uint64_t ma = (uint64_t)&map_a, mb = (uint64_t)&map_b;
            const uint32_t sa0 = smem_u32(smem), sb0 = smem_u32(smem_b);
            const uint32_t b_tx = (uint32_t)(p.b_boxes * p.b_box_bytes);
            auto load_b = [&](uint32_t dst, uint32_t bar, int chunk, int tap, int tile_n) {
                if (!p.dgrad) {       // Wf [taps][Cout][Cin], K-major: box {64 cin, n_tile cout, 1}
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                                 ::"r"(dst), "l"(mb), "r"(bar), "r"(chunk * 64), "r"(tile_n * p.n_tile), "r"(tap) : "memory");
                } else {              // same tensor, MN-major: one box {64 cin, 64 cout rows, 1} per 64 output columns
                    for (int blk = 0; blk < p.b_boxes; ++blk)
                        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                                     ::"r"(dst + (uint32_t)(blk * p.b_box_bytes)), "l"(mb), "r"(bar), "r"(tile_n * p.n_tile + blk * 64),
                                       "r"(chunk * 64), "r"(tap) : "memory");
                }
            };
            if (p.b_resident) {
                const uint32_t bar = smem_u32(&b_full[0]);
                const int full_chunks = p.b_tail ? p.n_chunks - 1 : p.n_chunks;
                const uint32_t tail_tx = p.b_tail ? (uint32_t)(p.taps * p.b_boxes * p.b_tail_box_bytes) : 0u;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b_tx * (uint32_t)(full_chunks * p.taps) + tail_tx) : "memory");
                for (int c = 0; c < full_chunks; ++c)
                    for (int t = 0; t < p.taps; ++t) load_b(sb0 + (uint32_t)((c * p.taps + t) * p.b_slot_bytes), bar, c, t, 0);
                if (p.b_tail) {          // dgrad only: MN-major boxes {64 cin, 16 * k_steps_last cout rows, 1 tap} from the second tensor map
                    const uint64_t mb2 = (uint64_t)&map_b2;
                    const int c = p.n_chunks - 1;
                    for (int t = 0; t < p.taps; ++t)
                        for (int blk = 0; blk < p.b_boxes; ++blk)
                            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                                         ::"r"(sb0 + b_tail_base + (uint32_t)(t * p.b_tail_slot_bytes + blk * p.b_tail_box_bytes)), "l"(mb2), "r"(bar),
                                           "r"(blk * 64), "r"(c * 64), "r"(t) : "memory");
                }
            }
            int as = 0, bs = 0;
            uint32_t aphase = 0, bphase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int tile_n = tile % p.grid_n, tile_m = tile / p.grid_n;
                for (int c = 0; c < p.n_chunks; ++c) {
                    mbar_wait(&a_empty[as], aphase ^ 1);
                    const uint32_t abar = smem_u32(&a_full[as]);
                    if (p.debug & 8) mbar_arrive(&a_full[as]);
                    else {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(abar), "r"((uint32_t)p.a_box_bytes) : "memory");
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                                 ::"r"(sa0 + (uint32_t)(as * p.a_slot_bytes)), "l"(ma), "r"(abar), "r"(c * 64), "r"(-p.pad), "r"(tile_m * p.bt) : "memory");
                    }
                    if (++as == TCV_A_SLOTS) { as = 0; aphase ^= 1; }
                    if (!p.b_resident) {
                        for (int t = 0; t < p.taps; t += p.tps) {
                            const int nt = min(p.tps, p.taps - t);
                            mbar_wait(&b_empty[bs], bphase ^ 1);
                            const uint32_t bar = smem_u32(&b_full[bs]);
                            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b_tx * (uint32_t)nt) : "memory");
                            for (int tt = 0; tt < nt; ++tt)
                                load_b(sb0 + (uint32_t)((bs * p.tps + tt) * p.b_slot_bytes), bar, c, t + tt, tile_n);
                            if (++bs == p.b_stages) { bs = 0; bphase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // The whole warp walks the loop (uniform control flow, so ptxas keeps descriptors in uniform registers); one
        // elected lane issues.  With the issuing code inside a divergent `if (lane == 0)` every tcgen05.mma cost ~25
        // SASS instructions (ELECT / R2UR / BRA.U.ANY waterfall): ~150 cycles per MMA, three times the 48-56 cycles a
        // 128 x 96 x 16 MMA needs.
        const uint32_t da_hi = (uint32_t)(umma_desc(0, 16, 1024) >> 32), db_hi = (uint32_t)(umma_desc(0, p.b_lbo_bytes, 1024) >> 32);
        const uint32_t da_lo16 = (uint32_t)umma_desc(0, 16, 1024), db_lo16 = (uint32_t)umma_desc(0, p.b_lbo_bytes, 1024);   // LBO field (bits 16-29)
        const uint32_t db_lo16_tail = (uint32_t)umma_desc(0, p.b_tail_lbo_bytes, 1024);
        const uint32_t sa0 = smem_u32(smem), sb0 = smem_u32(smem_b), idesc = p.idesc, bk = p.b_kstep16;
        const int taps = p.taps, n_chunks = p.n_chunks, resident = p.b_resident, b_stages = p.b_stages;
        const uint32_t a_slot = p.a_slot_bytes, b_slot = p.b_slot_bytes;
        const int shift0 = p.dgrad ? 2 * p.pad * 8 : 0, dshift = p.dgrad ? -8 : 8;      // in 16-byte descriptor units
        int as = 0, bs = 0, acc = 0;
        uint32_t aphase = 0, bphase = 0, acc_phase = 0;
        if (resident) { mbar_wait(&b_full[0], 0); tc_fence_after(); }
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_stride);
            uint32_t accumulate = 0;
            for (int c = 0; c < n_chunks; ++c) {
                mbar_wait(&a_full[as], aphase);
                tc_fence_after();
                uint32_t a_lo = (((sa0 + (uint32_t)as * a_slot) & 0x3FFFFu) >> 4) + (uint32_t)shift0;
                const int ks = (c == n_chunks - 1) ? p.k_steps_last : 4;
                const bool tail = resident && p.b_tail && c == n_chunks - 1;
                uint32_t b_res = ((tail ? sb0 + b_tail_base : sb0 + (uint32_t)(c * taps) * b_slot) & 0x3FFFFu) >> 4;
                const uint32_t b_res_step = (tail ? (uint32_t)p.b_tail_slot_bytes : b_slot) >> 4;
                const uint32_t db_lo = tail ? db_lo16_tail : db_lo16;
                if (resident) {
                    // no per-tap barriers: every MMA of this chunk (taps x K steps) goes out of ONE elected region, so the
                    // issuing thread spends a handful of uniform-datapath instructions per MMA (the per-tap loop below costs
                    // ~100 cycles per tap, which bounds the narrow N = 64..96 layers whose MMAs take only 50 cycles)
                    if (!(p.debug & 1) && elect_one_sync()) {
                        uint64_t da = ((uint64_t)da_hi << 32) | (da_lo16 | a_lo);
                        uint64_t db = ((uint64_t)db_hi << 32) | (db_lo | b_res);
                        const int64_t da_step = (int64_t)dshift;
                        const uint64_t db_step = (uint64_t)b_res_step;
                        if (ks == 4) {
#pragma unroll 1
                            for (int t = 0; t < taps; ++t) {
                                tc_mma_f16(d_tmem, da, db, idesc, accumulate);
                                tc_mma_f16(d_tmem, da + 2, db + bk, idesc, 1u);
                                tc_mma_f16(d_tmem, da + 4, db + 2 * bk, idesc, 1u);
                                tc_mma_f16(d_tmem, da + 6, db + 3 * bk, idesc, 1u);
                                accumulate = 1;
                                da += da_step; db += db_step;
                            }
                        } else {
#pragma unroll 1
                            for (int t = 0; t < taps; ++t) {
                                for (int s2 = 0; s2 < ks; ++s2) {
                                    tc_mma_f16(d_tmem, da + (uint64_t)(2 * s2), db + (uint64_t)(s2 * bk), idesc, accumulate);
                                    accumulate = 1;
                                }
                                da += da_step; db += db_step;
                            }
                        }
                    }
                    __syncwarp();
                    accumulate = 1;
                } else {
                    // streamed weights: one barrier round trip per stage of `tps` taps
                    const int tps = p.tps;
                    for (int t = 0; t < taps; t += tps) {
                        const int nt = min(tps, taps - t);
                        mbar_wait(&b_full[bs], bphase);
                        tc_fence_after();
                        if (!(p.debug & 1) && elect_one_sync()) {
                            uint64_t da = ((uint64_t)da_hi << 32) | (da_lo16 | a_lo);
                            uint64_t db = ((uint64_t)db_hi << 32) | (db_lo16 | (((sb0 + (uint32_t)(bs * tps) * b_slot) & 0x3FFFFu) >> 4));
                            for (int tt = 0; tt < nt; ++tt) {
                                tc_mma_f16(d_tmem, da, db, idesc, accumulate);
                                if (ks > 1) tc_mma_f16(d_tmem, da + 2, db + bk, idesc, 1u);
                                if (ks > 2) tc_mma_f16(d_tmem, da + 4, db + 2 * bk, idesc, 1u);
                                if (ks > 3) tc_mma_f16(d_tmem, da + 6, db + 3 * bk, idesc, 1u);
                                accumulate = 1;
                                da += (int64_t)dshift;
                                db += (uint64_t)(b_slot >> 4);
                            }
                            tc_commit(&b_empty[bs]);
                        }
                        __syncwarp();
                        accumulate = 1;
                        a_lo += (uint32_t)(dshift * nt);
                        if (++bs == b_stages) { bs = 0; bphase ^= 1; }
                    }
                }
                if (elect_one_sync()) tc_commit(&a_empty[as]);
                __syncwarp();
                if (++as == TCV_A_SLOTS) { as = 0; aphase ^= 1; }
            }
            if (elect_one_sync()) tc_commit(&tfull_bar[acc]);
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ================= epilogue (warps 2..5 own TMEM lane quarters warp%4) =================
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int grp = r / p.S, r_in = r - grp * p.S;
        const bool r_ok = grp < p.bt && r_in < p.L;
        int acc = 0, pool_chunk = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int tile_n = tile % p.grid_n, tile_m = tile / p.grid_n;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const int sample = tile_m * p.bt + grp;
            const int m = sample * p.L + r_in;
            const bool row_ok = r_ok && sample < p.Bn;
            const int n_base = tile_n * p.n_tile;
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.acc_stride);
            if (ep.mode == EPI_POOL) {
                for (int c0 = 0; c0 < p.n_tile; c0 += 16, ++pool_chunk) {         // the two staging buffers alternate ACROSS tiles as well
                    if (n_base + c0 >= p.N) break;                          // uniform over the four epilogue warps
                    float v[16];
                    tc_ld16(t_addr + (uint32_t)c0, v);
                    tc_pool_chunk16(ep, s_stats + (pool_chunk & 1) * (128 * TC_POOL_ROW), r, v, n_base + c0, p.N, p.S, p.bt, tile_m * p.bt, p.Bn);
                }
            } else
            if (!(p.debug & 2)) for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
                if (n_base + c0 >= p.N) break;
                float v[16];
                tc_ld16(t_addr + (uint32_t)c0, v);
                if (bn_on) tc_bn_stats16(ep, s_stats + q * 2 * p.N, p.N, n_base + c0, p.N, row_ok, v, lane);
                if (row_ok && !(p.debug & 4)) tc_epilogue_row16(ep, m, n_base + c0, 0, p.M, p.N, p.N, v);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (bn_on) for (int i = threadIdx.x; i < 2 * p.N; i += TC_THREADS)
        atomicAdd(&ep.bn_stats[i], (double)s_stats[i] + (double)s_stats[2 * p.N + i] + (double)s_stats[4 * p.N + i] + (double)s_stats[6 * p.N + i]);
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

inline int tc_conv_mode() { return tuning().conv_reuse; }

// Is the tap-reuse kernel applicable / preferable for this conv problem?  (fwd and dgrad only)
inline bool tc_conv_reuse_ok(const TcProblem& pr) {
    if (tc_conv_mode() == 0) return false;
    if (pr.kind != TC_CONV_FWD && pr.kind != TC_CONV_DGRAD) return false;
    if (pr.L > 128 || pr.taps < 2 || pr.taps > 15) return false;
    const int S = pr.L + pr.pad;
    const int bt_reuse = 1 + (128 - pr.L) / S, bt_plain = std::max(1, 128 / pr.L);
    if (tc_conv_mode() == 2) return true;
    // the shared halo costs MMA rows when many short samples share a tile (L = 25: 4 instead of 5 samples);
    // wide layers are already tensor-bound there, so keep the per-tap kernel unless the row utilisation is equal
    return bt_reuse * 10 >= bt_plain * 9;
}

inline int tc_conv_reuse(const TcProblem& pr, const Epilogue& ep, cudaStream_t st) {
    int rc = tc_init();
    if (rc) return rc;
    if (first_on_device(0)) {
        cudaError_t err = cudaFuncSetAttribute(tc_conv_reuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_max_smem());
        if (err != cudaSuccess) return set_error(-3, "cudaFuncSetAttribute(tc_conv_reuse_kernel): %s", cudaGetErrorString(err));
    }
    TcConvParams p = {};
    const bool dgrad = pr.kind == TC_CONV_DGRAD;
    const int Ca = dgrad ? pr.Cout : pr.Cin;          // channels of the staged activation (the GEMM K per tap)
    const int N = dgrad ? pr.Cin : pr.Cout;
    p.M = pr.B * pr.L; p.N = N; p.Bn = pr.B; p.L = pr.L; p.taps = pr.taps; p.pad = pr.pad; p.dgrad = dgrad ? 1 : 0;
    p.S = pr.L + pr.pad;
    p.bt = 1 + (128 - pr.L) / p.S;
    p.n_chunks = cdiv(Ca, 64);
    p.k_steps_last = cdiv(Ca - 64 * (p.n_chunks - 1), 16);
    p.n_tile = N <= 256 ? round_up(N, 16) : 256;
    p.grid_n = cdiv(N, p.n_tile);
    p.grid_m = cdiv(pr.B, p.bt);
    p.total_tiles = p.grid_m * p.grid_n;
    const int a_rows = round_up(128 + 2 * pr.pad, 8);
    p.a_slot_bytes = a_rows * 128;
    p.a_box_bytes = p.bt * p.S * 128;
    if (p.bt * p.S > a_rows) return set_error(-5, "tc_conv_reuse: tile rows exceed the A slot");
    CUtensorMap ma, mb;
    rc = make_map(&ma, pr.a, Ca, pr.L, pr.B, pr.lda, (int64_t)pr.L * pr.lda, 64, p.S, p.bt);
    if (rc) return rc;
    if (!dgrad) {
        rc = make_map(&mb, pr.b, pr.Cin, pr.Cout, pr.taps, pr.ldb, (int64_t)pr.Cout * pr.ldb, 64, p.n_tile, 1);
        p.b_boxes = 1; p.b_box_bytes = p.n_tile * 128; p.b_slot_bytes = round_up(p.n_tile * 128, 1024);
        p.b_kstep16 = 2; p.b_lbo_bytes = 16;
    } else {
        rc = make_map(&mb, pr.b, pr.Cin, pr.Cout, pr.taps, pr.ldb, (int64_t)pr.Cout * pr.ldb, 64, 64, 1);
        p.b_boxes = cdiv(p.n_tile, 64); p.b_box_bytes = 64 * 128; p.b_slot_bytes = p.b_boxes * p.b_box_bytes;
        p.b_kstep16 = (16 * 128) >> 4; p.b_lbo_bytes = p.b_box_bytes;
    }
    if (rc) return rc;
    CUtensorMap mb2 = mb;
    p.idesc = make_idesc(0, dgrad ? 1 : 0, p.n_tile);
    p.acc_stride = p.n_tile <= 32 ? 32 : p.n_tile <= 64 ? 64 : p.n_tile <= 128 ? 128 : 256;
    p.tmem_cols = 2 * p.acc_stride;
    p.debug = tuning().conv_debug;
    const int stats_bytes = ep.mode == EPI_POOL ? TC_POOL_SCRATCH_BYTES      // staging tiles of the fused pooling epilogue
                                                : ep.bn_stats ? round_up(8 * N * 4, 128) : 0;      // [4 epilogue warps][2][N] floats
    const int budget = tc_max_smem() - 2048 - stats_bytes - TCV_A_SLOTS * p.a_slot_bytes;
    int all_b = p.taps * p.n_chunks * p.b_slot_bytes;
    if (dgrad && p.k_steps_last < 4 && p.grid_n == 1 && all_b > budget) {
        // the last K chunk holds only 16 * k_steps_last cout rows: give it its own, smaller box so that every tap of W fits
        const int rows = 16 * p.k_steps_last;
        const int tail_box = rows * 128, tail_slot = p.b_boxes * tail_box;
        const int all_tail = p.taps * ((p.n_chunks - 1) * p.b_slot_bytes + tail_slot);
        if (all_tail <= budget) {
            rc = make_map(&mb2, pr.b, pr.Cin, pr.Cout, pr.taps, pr.ldb, (int64_t)pr.Cout * pr.ldb, 64, rows, 1);
            if (rc) return rc;
            p.b_tail = 1; p.b_tail_box_bytes = tail_box; p.b_tail_slot_bytes = tail_slot; p.b_tail_lbo_bytes = tail_box;
            all_b = all_tail;
        }
    }
    p.b_resident = (p.grid_n == 1 && all_b <= budget && !tuning().conv_no_resident) ? 1 : 0;
    if (!p.b_resident) p.b_tail = 0;
    p.tps = 1;
    if (!p.b_resident && !tuning().conv_tps1)
        for (int t : {5, 3, 2})               // the largest group that still leaves three stages in flight
            if (t <= p.taps && budget / (t * p.b_slot_bytes) >= 3 && t * p.b_slot_bytes <= 64 * 1024) { p.tps = t; break; }
    p.b_stages = std::min(TCV_MAX_B_STAGES, budget / (p.tps * p.b_slot_bytes));
    if (!p.b_resident && p.b_stages < 2) return set_error(-5, "tc_conv_reuse: weight stage of %d bytes does not fit twice", p.b_slot_bytes);
    const size_t smem = (size_t)TCV_A_SLOTS * p.a_slot_bytes + (size_t)(p.b_resident ? all_b : p.b_stages * p.tps * p.b_slot_bytes) + 1024 + 512 + stats_bytes;
    const int grid = std::min(p.total_tiles, tc_num_sms());
    tc_conv_reuse_kernel<<<grid, TC_THREADS, smem, st>>>(ma, mb, mb2, p, ep);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return set_error(-3, "tc_conv_reuse launch failed: %s", cudaGetErrorString(err));
    return 0;
}

// every tensor-core GEMM of the engine goes through here
inline int tc_dispatch(const TcProblem& pr, const Epilogue& ep, cudaStream_t st) {
    if (tc_conv_reuse_ok(pr)) return tc_conv_reuse(pr, ep, st);
    return tc_gemm(pr, ep, st);
}

}  // namespace emb
