// Data-parallel exchange steps as kernels over NVLink peer memory (one process per GPU).
//
// What couples the row shards of a data-parallel train step, and the kernel that carries it here:
//   BatchNorm batch statistics  [sum, sumsq] fwd / [sum dz, sum dz*xhat] bwd, 2*C doubles per conv layer
//                               -> one-shot all-reduce INSIDE the finalize kernels (bn_finalize_dp_kernel, bn_bwd_finalize_dp_kernel)
//   class weights of the loss   global positive count -> the same one-shot exchange inside ce_loss_dp_kernel
//   parameter gradients         -> dp_reduce_opt_kernel: reduce-scatter by peer LOADS (rank r sums slice r of every rank's
//                               gradient arena in fixed rank order), the optimizer rule on that slice only (the Adam moments
//                               of a slice live on its owner: optimizer state and work are sharded), and the all-gather of
//                               the updated PARAMETERS by peer STORES -- one kernel, no gradient ever travels twice
// No NCCL call and no host callback is on the step: every exchange is a few-microsecond latency-bound message pattern
// (8 ranks x <= 8 KB) or one pass over 1/W of the arena, all of it capturable in the step's CUDA graph.
//
// Protocol.  Every rank owns a DpComm block (device memory, mapped into every peer through CUDA IPC or, for engines that
// share a process, by plain pointer).  A step has a fixed list of sync points; `epoch` counts training steps (device
// counter, advanced by a kernel so that graph replays see it).  At sync point s a rank PUSHES its contribution into slot
// [epoch & 1][s][rank] of every rank's block (itself included), fences (system scope) and then release-stores `epoch` into
// flags[s][rank] of every block; it then acquire-polls the W flags of its OWN block (local memory) and reads the W slots in
// rank order, so every rank computes bit-identical sums.  Slots are double-buffered by epoch parity; a slot is rewritten
// two epochs later, by which time its readers have passed at least one full later barrier.
// Every wait happens in a ONE-CTA kernel of <= 256 threads (a spinning CTA holds almost no SM resources, so two engines that
// share one GPU -- the single-device test -- cannot starve each other), is bounded, and traps instead of hanging.
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace emb {

constexpr int DP_MAX_WORLD = 8;
constexpr int DP_MAX_SYNC = 12;           // 2 per conv layer (<= 8) + loss + gradients-ready + parameters-written
constexpr int DP_SLOT_DOUBLES = 1024;     // 2 * C, C <= 512 (the search space's widest conv layer)
constexpr int DP_SYNC_LOSS = 8, DP_SYNC_GRADS_READY = 9, DP_SYNC_PARAMS_DONE = 10;

struct DpComm {
    unsigned int flags[DP_MAX_SYNC][32];                                    // [sync][source rank], one 128-byte line per sync point
    double slots[2][DP_MAX_SYNC][DP_MAX_WORLD][DP_SLOT_DOUBLES];
};

struct DpCtx {
    int rank, world;
    DpComm* comm[DP_MAX_WORLD];           // comm[rank] is this rank's own block
    const unsigned int* epoch;            // device counter of training steps (own memory)
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// bounded wait (about 20 s with the back-off): a peer that never arrives becomes a launch error, not a hung GPU
__device__ __forceinline__ void dp_wait_flag(const unsigned int* f, unsigned int ep) {
    for (unsigned int spin = 0; spin < (1u << 24); ++spin) {
        if ((int)(ld_acquire_sys(f) - ep) >= 0) return;
        if (spin > 64) __nanosleep(spin > 4096 ? 1000 : 100);
    }
    printf("libembrace_sm100: data-parallel peer did not arrive (flag %p, epoch %u)\n", (const void*)f, ep);
    __trap();
}

__global__ void dp_epoch_advance_kernel(unsigned int* epoch) { *epoch += 1; }

// Whole-CTA collective: out[i] = sum over ranks of vals[i], i < n <= DP_SLOT_DOUBLES.  vals may live in shared or global
// memory; out may alias vals.  Deterministic: the W contributions are added in rank order on every rank.
__device__ __forceinline__ void dp_allreduce_small(const DpCtx& dp, int sync, const double* vals, double* out, int n) {
    const unsigned int ep = *dp.epoch;
    const int par = (int)(ep & 1u);
    for (int q = 0; q < dp.world; ++q) {
        double* dst = dp.comm[q]->slots[par][sync][dp.rank];
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = vals[i];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < dp.world) {
        st_release_sys(&dp.comm[threadIdx.x]->flags[sync][dp.rank], ep);
        dp_wait_flag(&dp.comm[dp.rank]->flags[sync][threadIdx.x], ep);
    }
    __syncthreads();
    const DpComm* me = dp.comm[dp.rank];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int q = 0; q < dp.world; ++q) s += __ldcg(&me->slots[par][sync][q][i]);
        out[i] = s;
    }
    __syncthreads();
}

// Barrier without payload (one CTA, >= world threads): "my writes before this kernel are visible" -> wait for everyone's.
__global__ void __launch_bounds__(32) dp_barrier_kernel(const DpCtx dp, int sync) {
    const unsigned int ep = *dp.epoch;
    __threadfence_system();
    if ((int)threadIdx.x < dp.world) {
        st_release_sys(&dp.comm[threadIdx.x]->flags[sync][dp.rank], ep);
        dp_wait_flag(&dp.comm[dp.rank]->flags[sync][threadIdx.x], ep);
    }
}

// SyncBN forward: all-reduce [sum | sumsq] (2*C doubles), then what bn_finalize_kernel does (scale/shift, mean/rstd, running
// statistics with the GLOBAL element count n).  stats is overwritten with the global sums.
__global__ void __launch_bounds__(256)
bn_finalize_dp_kernel(const DpCtx dp, int sync, double* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                      float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ scale, float* __restrict__ shift,
                      float* __restrict__ mean_out, float* __restrict__ rstd_out, double n, int C) {
    dp_allreduce_small(dp, sync, stats, stats, 2 * C);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double mu = stats[c] / n;
        double var = stats[C + c] / n - mu * mu;
        if (var < 0) var = 0;
        double rstd = 1.0 / sqrt(var + (double)BN_EPS);
        float sc = gamma[c] * (float)rstd;
        scale[c] = sc;
        shift[c] = beta[c] - (float)mu * sc;
        mean_out[c] = (float)mu;
        rstd_out[c] = (float)rstd;
        running_mean[c] = (1.f - BN_MOMENTUM) * running_mean[c] + BN_MOMENTUM * (float)mu;
        double unb = n > 1 ? var * n / (n - 1) : var;
        running_var[c] = (1.f - BN_MOMENTUM) * running_var[c] + BN_MOMENTUM * (float)unb;
    }
}

// SyncBN backward: all-reduce [sum dz | sum dz*xhat]; the global sums stay in bstats (the apply pass reads them).  dgamma /
// dbeta are written by rank 0 only (the gradient reduction sums the arenas of all ranks; the other ranks' entries stay zero).
__global__ void __launch_bounds__(256)
bn_bwd_finalize_dp_kernel(const DpCtx dp, int sync, double* __restrict__ bstats, float* __restrict__ dgamma, float* __restrict__ dbeta, int C) {
    dp_allreduce_small(dp, sync, bstats, bstats, 2 * C);
    if (dp.rank == 0)
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            dbeta[c] = (float)bstats[c];
            dgamma[c] = (float)bstats[C + c];
        }
}

// global positive count for get_loss_weights_from_labels (utils.py:121-140 sees the whole batch): one CTA
__global__ void __launch_bounds__(256)
dp_count_positives_kernel(const DpCtx dp, const int32_t* __restrict__ labels, int B, int64_t* __restrict__ n_pos_global) {
    __shared__ double s_val;
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int npos = 0;
    for (int i = threadIdx.x; i < B; i += blockDim.x) npos += labels[i] == 1;
    npos = __reduce_add_sync(0xffffffffu, npos);
    if ((threadIdx.x & 31) == 0 && npos) atomicAdd(&s_cnt, npos);
    __syncthreads();
    if (threadIdx.x == 0) s_val = (double)s_cnt;
    __syncthreads();
    dp_allreduce_small(dp, DP_SYNC_LOSS, &s_val, &s_val, 1);
    if (threadIdx.x == 0) *n_pos_global = (int64_t)(s_val + 0.5);
}

// ---------------------------------------------------------------------------------------------
// Gradient reduce-scatter + optimizer + parameter all-gather in one kernel.
// Rank r owns arena elements [lo, hi).  For each of them: g = sum_q grads_q[i] (peer loads, rank order), the optimizer
// rule of opt_step_kernel on (params_r[i], m[i], v[i]), then the new parameter is stored into EVERY rank's params arena.
// Preceded by dp_barrier_kernel(GRADS_READY) (every rank's backward has finished) and followed by
// dp_barrier_kernel(PARAMS_DONE) (every rank's stores have landed; nobody still reads my gradients).
// write_grads (tests / the split emb_backward + emb_opt_step API): also store the summed gradient into every rank's arena.
// ---------------------------------------------------------------------------------------------
struct DpArenas {
    float* params[DP_MAX_WORLD];
    float* grads[DP_MAX_WORLD];
};

__device__ __forceinline__ float opt_rule(const OptScalars& s, float pi, float gi, float& mi, float& vi) {
    if (s.kind == 1) pi *= (1.f - s.lr * s.wd);   // AdamW
    else gi = fmaf(s.wd, pi, gi);
    if (s.kind == 3) {                            // RMSprop
        vi = s.alpha * vi + (1.f - s.alpha) * gi * gi;
        pi -= s.lr * gi / (sqrtf(vi) + s.eps);
    } else {
        mi = s.b1 * mi + (1.f - s.b1) * gi;
        vi = s.b2 * vi + (1.f - s.b2) * gi * gi;
        if (s.kind == 2) {                        // Nadam
            float denom = sqrtf(vi / s.bc2) + s.eps;
            pi -= s.nadam_c_g * gi / denom;
            pi -= s.nadam_c_m * mi / denom;
        } else {
            float denom = sqrtf(vi) / s.bc2_sqrt + s.eps;
            pi -= (s.lr / s.bc1) * mi / denom;
        }
    }
    return pi;
}

__global__ void __launch_bounds__(256)
dp_reduce_opt_kernel(const int rank, const int world, const DpArenas ar, float* __restrict__ m, float* __restrict__ v,
                     const OptScalars* __restrict__ sp, long long lo, long long hi, int do_opt, int write_grads) {
    const OptScalars s = *sp;
    const long long n4 = (hi - lo) >> 2;          // lo, hi are multiples of 4 (every tensor is 16-byte aligned in the arena)
    for (long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += (long long)gridDim.x * blockDim.x) {
        const long long i = lo + 4 * i4;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = 0; q < world; ++q) {
            const float4 t = __ldcg(reinterpret_cast<const float4*>(ar.grads[q] + i));
            g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
        }
        if (write_grads)
            for (int q = 0; q < world; ++q) __stcg(reinterpret_cast<float4*>(ar.grads[q] + i), g);
        if (do_opt) {
            float4 p = *reinterpret_cast<const float4*>(ar.params[rank] + i);
            float4 mm = *reinterpret_cast<const float4*>(m + i), vv = *reinterpret_cast<const float4*>(v + i);
            p.x = opt_rule(s, p.x, g.x, mm.x, vv.x);
            p.y = opt_rule(s, p.y, g.y, mm.y, vv.y);
            p.z = opt_rule(s, p.z, g.z, mm.z, vv.z);
            p.w = opt_rule(s, p.w, g.w, mm.w, vv.w);
            *reinterpret_cast<float4*>(m + i) = mm;
            *reinterpret_cast<float4*>(v + i) = vv;
            for (int q = 0; q < world; ++q) __stcg(reinterpret_cast<float4*>(ar.params[q] + i), p);
        }
    }
    __threadfence_system();
}

}  // namespace emb
