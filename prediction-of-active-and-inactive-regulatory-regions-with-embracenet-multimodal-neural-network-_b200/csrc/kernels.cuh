// HBM-bound kernels of the EmbraceNet step (everything that is not a GEMM).
// Activations are channels-last [B, L, C] with leading dimension ld >= C; T is float or bf16.
#pragma once
#include "common.cuh"

namespace emb {

constexpr int SEQ_LEN = 256;
constexpr int POOL_K = 10;
constexpr int POOL_S = 2;
constexpr float BN_EPS = 1e-5f;
constexpr float BN_MOMENTUM = 0.1f;

// ---------------------------------------------------------------------------------------------
// input staging: x_ffnn fp32 [B,F] -> T [B, ld] (zero padded columns)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void cast_rows_kernel(const float* __restrict__ x, T* __restrict__ out, int B, int F, int ld) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * ld) return;
    int b = i / ld, f = i - (size_t)b * ld;
    out[i] = from_f<T>(f < F ? x[(size_t)b * F + f] : 0.f);
}

// ---------------------------------------------------------------------------------------------
// K1: first Conv1d over one-hot input == gather-sum of weight rows (CNN_pre.py:39, in_channels=4)
//   y[b,l,o] = bias[o] + sum_tap W[o][base[b,l+tap-p]][tap]      (taps falling outside [0,256) add 0)
// One CTA per sample (grid-strided); the [k][4][C1] table and the sample's bases live in shared memory.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
onehot_conv_fwd_kernel(const uint8_t* __restrict__ bases, const float* __restrict__ w, const float* __restrict__ bias,
                       T* __restrict__ y, int B, int C1, int k, int ld, int round_w) {
    extern __shared__ float smem[];
    float* tab = smem;                         // [k][4][C1]
    float* sb = tab + k * 4 * C1;              // [C1]
    uint8_t* sbase = (uint8_t*)(sb + C1);      // [256 + 2p], 4 = "outside"
    const int p = (k - 1) / 2;
    for (int i = threadIdx.x; i < k * 4 * C1; i += blockDim.x) {
        int tap = i / (4 * C1), r = i - tap * 4 * C1, c = r / C1, o = r - c * C1;
        float v = w[((size_t)o * 4 + c) * k + tap];
        tab[i] = round_w ? __bfloat162float(__float2bfloat16_rn(v)) : v;
    }
    for (int i = threadIdx.x; i < C1; i += blockDim.x) sb[i] = bias[i];
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < SEQ_LEN + 2 * p; i += blockDim.x) {
            int l = i - p;
            sbase[i] = (l >= 0 && l < SEQ_LEN) ? bases[(size_t)b * SEQ_LEN + l] : 4;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < SEQ_LEN * C1; i += blockDim.x) {
            int l = i / C1, o = i - l * C1;
            float acc = sb[o];
            for (int tap = 0; tap < k; ++tap) {
                int c = sbase[l + tap];
                if (c < 4) acc += tab[(tap * 4 + c) * C1 + o];
            }
            y[((size_t)b * SEQ_LEN + l) * ld + o] = from_f<T>(acc);
        }
    }
}

// K1 backward: dW[o][c][tap] = sum_{b,l} dy[b,l,o] * [base[b,l+tap-p] == c]; dbias[o] = sum dy.
// Each thread owns up to 4 (o,tap) pairs and keeps 4 per-base accumulators for each in registers;
// a CTA sweeps several samples before it touches global memory (one atomicAdd per owned entry).
constexpr int OHB_MAXP = 4;
template <typename T>
__global__ void __launch_bounds__(256)
onehot_conv_bwd_kernel(const uint8_t* __restrict__ bases, const T* __restrict__ dy, float* __restrict__ dw,
                       float* __restrict__ dbias, int B, int C1, int k, int ld) {
    __shared__ uint8_t sbase[SEQ_LEN + 16];
    const int p = (k - 1) / 2;
    const int npairs = C1 * k;
    float acc[OHB_MAXP][4];
#pragma unroll
    for (int q = 0; q < OHB_MAXP; ++q)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[q][c] = 0.f;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < SEQ_LEN + 2 * p; i += blockDim.x) {
            int l = i - p;
            sbase[i] = (l >= 0 && l < SEQ_LEN) ? bases[(size_t)b * SEQ_LEN + l] : 4;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < OHB_MAXP; ++q) {
            int pair = threadIdx.x + q * 256;
            if (pair >= npairs) break;
            int tap = pair / C1, o = pair - tap * C1;     // o fastest: coalesced dy reads
            const T* g = dy + (size_t)b * SEQ_LEN * ld + o;
            for (int l = 0; l < SEQ_LEN; ++l) {
                int c = sbase[l + tap];
                float v = to_f(g[(size_t)l * ld]);
                acc[q][0] += (c == 0) ? v : 0.f;
                acc[q][1] += (c == 1) ? v : 0.f;
                acc[q][2] += (c == 2) ? v : 0.f;
                acc[q][3] += (c == 3) ? v : 0.f;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < OHB_MAXP; ++q) {
        int pair = threadIdx.x + q * 256;
        if (pair >= npairs) break;
        int tap = pair / C1, o = pair - tap * C1;
#pragma unroll
        for (int c = 0; c < 4; ++c) atomicAdd(&dw[((size_t)o * 4 + c) * k + tap], acc[q][c]);
        if (dbias && tap == p) atomicAdd(&dbias[o], acc[q][0] + acc[q][1] + acc[q][2] + acc[q][3]);  // centre tap sees every l
    }
}

// ---------------------------------------------------------------------------------------------
// BatchNorm1d batch statistics over (B,L): per-channel sum and sum of squares (fp32 partials, fp64 merge)
// stats[0..C) += sum, stats[C..2C) += sumsq
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
bn_stats_kernel(const T* __restrict__ y, double* __restrict__ stats, int64_t R, int C, int ld) {
    __shared__ float s1[8][33], s2[8][33];
    int c = blockIdx.x * 32 + threadIdx.x;
    float a = 0.f, q = 0.f;
    if (c < C)
        for (int64_t r = (int64_t)blockIdx.y * 8 + threadIdx.y; r < R; r += (int64_t)gridDim.y * 8) {
            float v = to_f(y[r * ld + c]);
            a += v;
            q += v * v;
        }
    s1[threadIdx.y][threadIdx.x] = a;
    s2[threadIdx.y][threadIdx.x] = q;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        double da = 0, dq = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { da += s1[i][threadIdx.x]; dq += s2[i][threadIdx.x]; }
        atomicAdd(&stats[c], da);
        atomicAdd(&stats[C + c], dq);
    }
}

// mean/rstd -> fused scale/shift; running stats update (momentum 0.1, unbiased running_var)
__global__ void bn_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ rstd_out, double n, int C, int training) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (training) {
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
    } else {
        float rstd = 1.0f / sqrtf(running_var[c] + BN_EPS);
        float sc = gamma[c] * rstd;
        scale[c] = sc;
        shift[c] = beta[c] - running_mean[c] * sc;
    }
}

// ---------------------------------------------------------------------------------------------
// K2: BatchNorm apply + ReLU + MaxPool1d(10, 2) + Dropout in one pass (CNN_pre.py:41-51)
//   a[b,j,c] = drop( max_{i in [2j,2j+9]} relu(y[b,i,c]*scale[c] + shift[c]) )
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
bn_relu_pool_drop_fwd_kernel(const T* __restrict__ y, const float* __restrict__ scale,
                             const float* __restrict__ shift, T* __restrict__ a, int B, int Lc, int Lp, int C,
                             int ld, float drop_p, const float* __restrict__ drop_u, RngState const* rng,
                             uint32_t rng_stream, int64_t row_offset) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)B * Lp * C;
    if (i >= total) return;
    int c = i % C;
    size_t bj = i / C;
    int j = bj % Lp, b = bj / Lp;
    float sc = scale[c], sh = shift[c];
    const T* src = y + ((size_t)b * Lc + (size_t)j * POOL_S) * ld + c;
    float m = 0.f;  // relu folded into the max
#pragma unroll
    for (int t = 0; t < POOL_K; ++t) m = fmaxf(m, fmaf(to_f(src[(size_t)t * ld]), sc, sh));
    if (drop_p > 0.f) {
        size_t ref_idx = ((size_t)b * C + c) * Lp + j;            // reference layout [B,C,Lp]
        float u = drop_u ? drop_u[ref_idx]
                         : rng_cnn_uniform(*rng, rng_stream, (uint64_t)(row_offset + b), C, c, Lp, j);
        m = (u >= drop_p) ? m / (1.f - drop_p) : 0.f;
    }
    a[((size_t)b * Lp + j) * ld + c] = from_f<T>(m);
}

// K2 backward, stage 1: route d(a) through dropout, the max-pool arg-max (first maximum) and ReLU to
// dz, stored in the dy buffer; accumulate the BatchNorm reductions sum(dz) and sum(dz * xhat).
// One CTA = one sample x 32 channels; the sample's y column block and dz live in shared memory.
template <typename T>
__global__ void __launch_bounds__(256)
pool_bn_bwd_stage1_kernel(const T* __restrict__ y, const T* __restrict__ a, const T* __restrict__ ga,
                          const float* __restrict__ scale, const float* __restrict__ shift,
                          const float* __restrict__ mean, const float* __restrict__ rstd, T* __restrict__ dz_out,
                          double* __restrict__ bstats, int B, int Lc, int Lp, int C, int ld, float drop_p) {
    extern __shared__ float smem[];
    float* ys = smem;               // [Lc][32]
    float* dzs = smem + Lc * 32;    // [Lc][32]
    __shared__ float r1[8][33], r2[8][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c = blockIdx.x * 32 + tx;
    const int b = blockIdx.y;
    const bool ok = c < C;
    for (int l = ty; l < Lc; l += 8) {
        ys[l * 32 + tx] = ok ? to_f(y[((size_t)b * Lc + l) * ld + c]) : 0.f;
        dzs[l * 32 + tx] = 0.f;
    }
    __syncthreads();
    const float sc = ok ? scale[c] : 0.f, sh = ok ? shift[c] : 0.f;
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    if (ok)
        for (int j = ty; j < Lp; j += 8) {
            float av = to_f(a[((size_t)b * Lp + j) * ld + c]);
            if (!(av > 0.f)) continue;   // dropped, or the whole window was <= 0 (ReLU kills the gradient)
            float g = to_f(ga[((size_t)b * Lp + j) * ld + c]) * inv_keep;
            int best = 0;
            float bm = -INFINITY;
#pragma unroll
            for (int t = 0; t < POOL_K; ++t) {
                float z = fmaf(ys[(j * POOL_S + t) * 32 + tx], sc, sh);
                if (z > bm) { bm = z; best = t; }   // strict '>' keeps the FIRST maximum, as PyTorch does
            }
            atomicAdd(&dzs[(j * POOL_S + best) * 32 + tx], g);
        }
    __syncthreads();
    float s_dz = 0.f, s_dzx = 0.f;
    if (ok) {
        const float mu = mean[c], rs = rstd[c];
        for (int l = ty; l < Lc; l += 8) {
            float dz = dzs[l * 32 + tx];
            dz_out[((size_t)b * Lc + l) * ld + c] = from_f<T>(dz);
            s_dz += dz;
            s_dzx += dz * (ys[l * 32 + tx] - mu) * rs;
        }
    }
    r1[ty][tx] = s_dz;
    r2[ty][tx] = s_dzx;
    __syncthreads();
    if (ty == 0 && ok) {
        double d1 = 0, d2 = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { d1 += r1[i][tx]; d2 += r2[i][tx]; }
        atomicAdd(&bstats[c], d1);        // dbeta
        atomicAdd(&bstats[C + c], d2);    // dgamma
    }
}

// K2 backward, stage 2: dy = gamma*rstd * (dz - dbeta/n - xhat*dgamma/n), in place; conv dbias = sum(dy)
template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const T* __restrict__ y, T* __restrict__ dz, const double* __restrict__ bstats,
                    const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                    float* __restrict__ dbias, int64_t R, int C, int ld, double n) {
    __shared__ float s1[8][33];
    int c = blockIdx.x * 32 + threadIdx.x;
    float acc = 0.f;
    if (c < C) {
        float dbeta_n = (float)(bstats[c] / n), dgamma_n = (float)(bstats[C + c] / n);
        float mu = mean[c], rs = rstd[c], gsc = gamma[c] * rs;
        for (int64_t r = (int64_t)blockIdx.y * 8 + threadIdx.y; r < R; r += (int64_t)gridDim.y * 8) {
            float xhat = (to_f(y[r * ld + c]) - mu) * rs;
            float v = gsc * (to_f(dz[r * ld + c]) - dbeta_n - xhat * dgamma_n);
            dz[r * ld + c] = from_f<T>(v);
            acc += v;
        }
    }
    s1[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (dbias && threadIdx.y == 0 && c < C) {
        float t = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += s1[i][threadIdx.x];
        atomicAdd(&dbias[c], t);
    }
}

__global__ void bn_bwd_finalize_kernel(const double* __restrict__ bstats, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    dbeta[c] = (float)bstats[c];
    dgamma[c] = (float)bstats[C + c];
}

// ---------------------------------------------------------------------------------------------
// embracement prologue (EmbraceNetMultimodal.py:178-187, :63-76): modality dropout -> availabilities,
// fp32 probability normalisation, fp64 cumulative threshold cum0[b] used by idx = (u > cum0)
// ---------------------------------------------------------------------------------------------
__global__ void embrace_prologue_kernel(double p_ffnn, const float* __restrict__ avail, int training_dropout,
                                        int has_u0, float u0, const float* __restrict__ modal_rows,
                                        const RngState* rng, int64_t row_offset, double* __restrict__ cum0,
                                        int B) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float a0 = 1.f, a1 = 1.f;
    if (avail) { a0 = avail[2 * b]; a1 = avail[2 * b + 1]; }
    if (training_dropout) {
        float coin = has_u0 ? u0 : rng_uniform_f32(*rng, RNG_MODAL_COIN, 0);
        if (coin >= 0.5f) {
            float ur = modal_rows ? modal_rows[b] : rng_uniform_f32(*rng, RNG_MODAL_ROWS, (uint64_t)(row_offset + b));
            int target = (int)rintf(ur);   // torch.round: half to even
            a0 = target == 0 ? 1.f : 0.f;
            a1 = target == 1 ? 1.f : 0.f;
        }
    }
    float s0 = (float)p_ffnn, s1 = (float)(1.0 - p_ffnn);   // torch.tensor([p, 1.0-p]) -> fp32
    float p0 = s0 * a0, p1 = s1 * a1;
    float sum = p0 + p1;
    p0 = p0 / sum;
    p1 = p1 / sum;
    cum0[b] = (double)p0 / ((double)p0 + (double)p1);
}

// stand-alone select (used when the docking_1 GEMM is not the SIMT kernel with the fused epilogue)
template <typename T>
__global__ void embrace_select_kernel(const T* __restrict__ d0, const T* __restrict__ d1, const double* __restrict__ u,
                                      const double* __restrict__ cum0, const RngState* rng, int64_t row_offset,
                                      T* __restrict__ e, uint8_t* __restrict__ idx, int B, int C, int ld) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * C) return;
    int b = i / C, c = i - (size_t)b * C;
    double uu = u ? u[i] : rng_uniform_f64(*rng, RNG_EMBRACE, (uint64_t)(row_offset + b) * C + c);
    int id = uu > cum0[b];
    e[(size_t)b * ld + c] = id ? d1[(size_t)b * ld + c] : d0[(size_t)b * ld + c];
    idx[i] = (uint8_t)id;
}

// ---------------------------------------------------------------------------------------------
// K6 + K8: class-weighted 2-class cross entropy on fp32 logits, its gradient, and the confusion
// counts of the hard predictions (training_models_multimodal.py:140-162, utils.py:80-140).
//   w_pos = N_neg / N, w_neg = N_pos / N  (no positives: w_pos = 0, w_neg = 1; no negatives: mirrored)
// Single CTA (B <= a few 10^4 rows).  n_pos_global / n_global >= 0 override the local counts (data parallel).
// ---------------------------------------------------------------------------------------------
struct StepMetricsDev { float loss; int tp, fp, fn, tn; };

__global__ void __launch_bounds__(1024)
ce_loss_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels, int B,
               int64_t n_pos_global, int64_t n_global, float* __restrict__ dlogits, StepMetricsDev* __restrict__ rec_base,
               int* __restrict__ rec_count, int max_rec, const int64_t* __restrict__ n_pos_global_dev = nullptr) {
    __shared__ int s_cnt[5];
    __shared__ float s_red[32];
    __shared__ float s_w[3];
    const int t = threadIdx.x;
    if (t < 5) s_cnt[t] = 0;
    __syncthreads();
    int npos = 0;
    for (int i = t; i < B; i += blockDim.x) npos += labels[i] == 1;
    npos = __reduce_add_sync(0xffffffffu, npos);
    if ((t & 31) == 0 && npos) atomicAdd(&s_cnt[4], npos);
    __syncthreads();
    if (t == 0) {
        double pos = n_global >= 0 ? (double)(n_pos_global_dev ? *n_pos_global_dev : n_pos_global) : (double)s_cnt[4];
        double tot = n_global >= 0 ? (double)n_global : (double)B;
        double neg = tot - pos;
        double pos_inv = pos != 0 ? 1.0 / pos : 0.0, neg_inv = neg != 0 ? 1.0 / neg : 0.0;
        double wp = pos_inv / (neg_inv + pos_inv), wn = neg_inv / (neg_inv + pos_inv);
        s_w[0] = (float)wn;
        s_w[1] = (float)wp;
        // sum of per-row weights over the (global) batch
        s_w[2] = (float)wn * (float)neg + (float)wp * (float)pos;
    }
    __syncthreads();
    const float wn = s_w[0], wp = s_w[1], wsum = s_w[2];
    float lsum = 0.f;
    int tp = 0, fp = 0, fn = 0, tn = 0;
    for (int i = t; i < B; i += blockDim.x) {
        float z0 = logits[2 * i], z1 = logits[2 * i + 1];
        int yv = labels[i];
        float m = fmaxf(z0, z1);
        float e0 = expf(z0 - m), e1 = expf(z1 - m);
        float se = e0 + e1;
        float lse = logf(se) + m;
        float w = yv == 1 ? wp : wn;
        lsum += w * (lse - (yv == 1 ? z1 : z0));
        if (dlogits) {
            float p0 = e0 / se, p1 = e1 / se;
            dlogits[2 * i] = w * (p0 - (yv == 0 ? 1.f : 0.f)) / wsum;
            dlogits[2 * i + 1] = w * (p1 - (yv == 1 ? 1.f : 0.f)) / wsum;
        }
        int pred = z1 > z0;   // torch.argmax: ties -> class 0
        tp += pred & (yv == 1);
        fp += pred & (yv != 1);
        fn += (!pred) & (yv == 1);
        tn += (!pred) & (yv != 1);
    }
    lsum = warp_sum(lsum);
    tp = __reduce_add_sync(0xffffffffu, tp);
    fp = __reduce_add_sync(0xffffffffu, fp);
    fn = __reduce_add_sync(0xffffffffu, fn);
    tn = __reduce_add_sync(0xffffffffu, tn);
    if ((t & 31) == 0) {
        s_red[t >> 5] = lsum;
        atomicAdd(&s_cnt[0], tp);
        atomicAdd(&s_cnt[1], fp);
        atomicAdd(&s_cnt[2], fn);
        atomicAdd(&s_cnt[3], tn);
    }
    __syncthreads();
    if (t == 0) {
        float tot = 0.f;
        for (int i = 0; i < (blockDim.x >> 5); ++i) tot += s_red[i];
        int slot = *rec_count;
        if (slot < max_rec) {
            StepMetricsDev r;
            r.loss = tot / wsum;
            r.tp = s_cnt[0]; r.fp = s_cnt[1]; r.fn = s_cnt[2]; r.tn = s_cnt[3];
            rec_base[slot] = r;
        }
        *rec_count = slot + 1;
    }
}

// softmax(logits)[:,1] (EmbraceNetMultimodal_NoTrain.py:210-214)
__global__ void softmax_p1_kernel(const float* __restrict__ logits, float* __restrict__ probs, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    float z0 = logits[2 * i], z1 = logits[2 * i + 1];
    float m = fmaxf(z0, z1);
    float e0 = expf(z0 - m), e1 = expf(z1 - m);
    probs[i] = e1 / (e0 + e1);
}

// bias gradient of a Linear layer: out[n] += sum_m g[m,n]
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ g, float* __restrict__ out, int64_t R, int N, int ld) {
    __shared__ float s1[8][33];
    int c = blockIdx.x * 32 + threadIdx.x;
    float acc = 0.f;
    if (c < N)
        for (int64_t r = (int64_t)blockIdx.y * 8 + threadIdx.y; r < R; r += (int64_t)gridDim.y * 8)
            acc += to_f(g[r * ld + c]);
    s1[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < N) {
        float t = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += s1[i][threadIdx.x];
        atomicAdd(&out[c], t);
    }
}

template <typename T>
__global__ void cast_f32_kernel(const float* __restrict__ src, T* __restrict__ dst, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = from_f<T>(src[i]);
}

// ---------------------------------------------------------------------------------------------
// K7: fused multi-tensor optimizer over the flat parameter arena.
// All reference optimizers use coupled L2 (g += wd*p); ADAMW is the decoupled variant.
// Scalars that depend on the step count are computed on the host per call (graph replay updates them
// through OptScalars in device memory).
// ---------------------------------------------------------------------------------------------
struct OptScalars {
    int kind;
    float lr, wd, b1, b2, eps, alpha;
    float bc1, bc2_sqrt, bc2;         // 1-b1^t, sqrt(1-b2^t), 1-b2^t
    float nadam_c_g, nadam_c_m;       // lr(1-mu_t)/(1-prod), lr*mu_{t+1}/(1-prod*mu_{t+1})
};

__global__ void __launch_bounds__(256)
opt_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                const OptScalars* __restrict__ sp, size_t n) {
    const OptScalars s = *sp;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float pi = p[i], gi = g[i];
        if (s.kind == 1) pi *= (1.f - s.lr * s.wd);   // AdamW
        else gi = fmaf(s.wd, pi, gi);
        if (s.kind == 3) {                            // RMSprop
            float vi = s.alpha * v[i] + (1.f - s.alpha) * gi * gi;
            v[i] = vi;
            pi -= s.lr * gi / (sqrtf(vi) + s.eps);
        } else {
            float mi = s.b1 * m[i] + (1.f - s.b1) * gi;
            float vi = s.b2 * v[i] + (1.f - s.b2) * gi * gi;
            m[i] = mi;
            v[i] = vi;
            if (s.kind == 2) {                        // Nadam
                float denom = sqrtf(vi / s.bc2) + s.eps;
                pi -= s.nadam_c_g * gi / denom;
                pi -= s.nadam_c_m * mi / denom;
            } else {
                float denom = sqrtf(vi) / s.bc2_sqrt + s.eps;
                pi -= (s.lr / s.bc1) * mi / denom;
            }
        }
        p[i] = pi;
    }
}

__global__ void rng_advance_kernel(RngState* s) { s->step += 1; }

}  // namespace emb
