// Generic SIMT GEMM  C[m,n] = sum_k A(m,k) * B(n,k)  with fp32 accumulation.
//
// This is the shape-generic back end: it serves (a) the fp32 verification precision, (b) shapes the
// tcgen05 kernels do not take (K or N below the UMMA/TMA granularity: the 2-logit head, FFNN widths
// of 4, in_features that are not a multiple of 8, test-sized channels) and (c) the on-device cross
// check of the tensor-core kernels.  Operands are described by accessors so that nn.Linear, the
// Conv1d implicit GEMMs (fwd / dgrad / wgrad) and the permuted docking_1 weight all run through the
// same kernel without materialising im2col or transposed copies.
#pragma once
#include "common.cuh"

namespace emb {

enum OperandMode : int {
    OP_ROWMAJOR = 0,      // (r,k) -> p[r*ld + k]
    OP_TRANSPOSED = 1,    // (r,k) -> p[k*ld + r]
    OP_CONV_SHIFT = 2,    // r=(b,l), k=(tap,c): x[b, l + sign*(tap-pad), c] channels-last, 0 outside [0,L)
    OP_CONV_SHIFT_T = 3,  // r=(tap,c), k=(b,l): same element, transposed roles (wgrad B operand)
    OP_CONV_W_FWD = 4,    // r=o, k=(tap,c): W[o][c][tap]   (reference Conv1d weight layout [Cout,Cin,k])
    OP_CONV_W_DGRAD = 5,  // r=c, k=(tap,o): W[o][c][tap]
    OP_FLAT_ACT = 6,      // r=b, k=(l,c): act[b, l, c] (channels-last, padded ld) as the flattened row
    OP_FLAT_ACT_T = 7,    // r=(l,c), k=b
    OP_W_PERM = 8,        // r=n, k=(l,c): W[n][c*L + l]  (reference flatten order c*L+l, CNN_pre.py:73)
    OP_W_PERM_T = 9,      // r=(l,c), k=n: W[k][c*L + l]
};

struct Operand {
    const void* p;
    int dtype;       // 0 fp32, 1 bf16
    int mode;
    int ld;          // leading dimension in elements (meaning depends on mode)
    int rows, cols;  // logical extent: rows = M or N, cols = K
    int L, C;        // conv / flatten geometry: L positions, C channels of the indexed activation
    int taps, pad, sign;
    int wrows;       // OP_CONV_W_DGRAD: Cin; OP_W_PERM*: K of the weight row
    int perm_off;    // OP_W_PERM*: leading columns that are not permuted
    int round_bf16;  // round fp32 source values to bf16 on load (bf16 precision with fp32 master weights)
};

__device__ __forceinline__ float operand_fetch(const Operand& o, int r, int k) {
    if (r >= o.rows || k >= o.cols) return 0.f;
    size_t idx;
    switch (o.mode) {
        case OP_ROWMAJOR: idx = (size_t)r * o.ld + k; break;
        case OP_TRANSPOSED: idx = (size_t)k * o.ld + r; break;
        case OP_CONV_SHIFT: {
            int tap = k / o.C, c = k - tap * o.C;
            int b = r / o.L, l = r - b * o.L;
            int ls = l + o.sign * (tap - o.pad);
            if (ls < 0 || ls >= o.L) return 0.f;
            idx = ((size_t)b * o.L + ls) * o.ld + c;
        } break;
        case OP_CONV_SHIFT_T: {
            int tap = r / o.C, c = r - tap * o.C;
            int b = k / o.L, l = k - b * o.L;
            int ls = l + o.sign * (tap - o.pad);
            if (ls < 0 || ls >= o.L) return 0.f;
            idx = ((size_t)b * o.L + ls) * o.ld + c;
        } break;
        case OP_CONV_W_FWD: {
            int tap = k / o.C, c = k - tap * o.C;
            idx = ((size_t)r * o.C + c) * o.taps + tap;
        } break;
        case OP_CONV_W_DGRAD: {
            int tap = k / o.C, oc = k - tap * o.C;
            idx = ((size_t)oc * o.wrows + r) * o.taps + tap;
        } break;
        case OP_FLAT_ACT: {
            int l = k / o.C, c = k - l * o.C;
            idx = ((size_t)r * o.L + l) * o.ld + c;
        } break;
        case OP_FLAT_ACT_T: {
            int l = r / o.C, c = r - l * o.C;
            idx = ((size_t)k * o.L + l) * o.ld + c;
        } break;
        case OP_W_PERM: {
            if (k < o.perm_off) { idx = (size_t)r * o.wrows + k; break; }
            int k2 = k - o.perm_off, l = k2 / o.C, c = k2 - l * o.C;
            idx = (size_t)r * o.wrows + o.perm_off + (size_t)c * o.L + l;
        } break;
        default: {  // OP_W_PERM_T
            if (r < o.perm_off) { idx = (size_t)k * o.wrows + r; break; }
            int r2 = r - o.perm_off, l = r2 / o.C, c = r2 - l * o.C;
            idx = (size_t)k * o.wrows + o.perm_off + (size_t)c * o.L + l;
        } break;
    }
    float v = o.dtype ? __bfloat162float(((const bf16*)o.p)[idx]) : ((const float*)o.p)[idx];
    if (o.round_bf16) v = __bfloat162float(__float2bfloat16_rn(v));
    return v;
}

enum EpiMode : int {
    EPI_LINEAR = 0,       // out = [dropout]([relu](acc + bias))          nn.Linear -> ReLU -> Dropout
    EPI_EMBRACE = 1,      // d1 = relu(acc+bias); idx = u > cum0; out = idx ? d1 : d0; writes idx
    EPI_MASKGRAD = 2,     // out = ref > 0 ? acc * scale : 0              (dgrad through ReLU+Dropout)
    EPI_EMBRACE_BWD = 3,  // dd0/dd1 = acc masked by idx and e > 0
    EPI_ATOMIC = 4,       // atomicAdd(out_f32[map(m,n)], acc)            (wgrad, split-K)
    EPI_POOL = 5,         // tensor-core conv forward, inference: eval BatchNorm + ReLU + MaxPool1d(10,2) on the tile, pooled rows out
};
enum OutMap : int { MAP_ROWMAJOR = 0, MAP_CONV_W = 1, MAP_W_PERM = 2 };

struct Epilogue {
    int mode;
    void* out;
    int out_dtype, ldo;
    const float* bias;
    int relu;
    // dropout (EPI_LINEAR)
    float drop_p;
    const float* drop_u;     // replay uniforms, reference layout [M_global?, N]: index m*N + n
    uint32_t rng_stream;
    const RngState* rng;
    int64_t row_offset;      // global row of local row 0 (Philox counters)
    int flatC, flat_ldc;     // EPI_LINEAR store remap n=(l,c) -> m*ldo + l*flat_ldc + c (0: plain m*ldo + n)
    // EPI_EMBRACE
    const void* d0;
    int ld_d0;
    const double* emb_u;     // replay [M, N] fp64 or NULL
    const double* cum0;      // [M]
    uint8_t* idx_out;        // [M, N]
    // EPI_MASKGRAD
    const void* ref;
    int ld_ref;
    float scale;
    // EPI_EMBRACE_BWD
    const uint8_t* idx;
    const void* e;
    int ld_e;
    void* out2;
    // EPI_ATOMIC
    int map, mapC, mapL, map_taps, map_wrows, map_off;
    // tensor-core conv forward only: per-column sum / sum of squares of the STORED (bf16-rounded) outputs, i.e. the
    // BatchNorm batch statistics of the layer, accumulated in the epilogue ([N] sums then [N] sums of squares)
    double* bn_stats;
    // EPI_POOL (tensor-core conv forward, eval mode): a[b, j, n] = max_{i<10} relu(bf16(acc + bias) * pool_scale[n] + pool_shift[n]) at l = 2j + i
    const float* pool_scale;
    const float* pool_shift;
    void* pool_out;          // bf16 [B, pool_Lp, pool_ld]
    int pool_Lp, pool_ld;
};

__device__ __forceinline__ void epilogue_apply(const Epilogue& ep, int m, int n, int M, int N, float acc) {
    switch (ep.mode) {
        case EPI_LINEAR: {
            float v = acc + (ep.bias ? ep.bias[n] : 0.f);
            if (ep.relu) v = fmaxf(v, 0.f);
            if (ep.drop_p > 0.f) {
                float u = ep.drop_u ? ep.drop_u[(size_t)m * N + n]
                                    : rng_uniform_f32(*ep.rng, ep.rng_stream, (uint64_t)(ep.row_offset + m) * N + n);
                v = (u >= ep.drop_p) ? v / (1.f - ep.drop_p) : 0.f;
            }
            size_t oi = (size_t)m * ep.ldo + n;
            if (ep.flatC) { int l = n / ep.flatC; oi = (size_t)m * ep.ldo + (size_t)l * ep.flat_ldc + (n - l * ep.flatC); }
            store_act(ep.out, ep.out_dtype, oi, v);
        } break;
        case EPI_EMBRACE: {
            float d1 = fmaxf(acc + ep.bias[n], 0.f);
            double u = ep.emb_u ? ep.emb_u[(size_t)m * N + n]
                                : rng_uniform_f64(*ep.rng, RNG_EMBRACE, (uint64_t)(ep.row_offset + m) * N + n);
            int id = u > ep.cum0[m];
            float d0 = load_act(ep.d0, ep.out_dtype, (size_t)m * ep.ld_d0 + n);
            store_act(ep.out, ep.out_dtype, (size_t)m * ep.ldo + n, id ? d1 : d0);
            ep.idx_out[(size_t)m * N + n] = (uint8_t)id;
        } break;
        case EPI_MASKGRAD: {
            float r = load_act(ep.ref, ep.out_dtype, (size_t)m * ep.ld_ref + n);
            store_act(ep.out, ep.out_dtype, (size_t)m * ep.ldo + n, r > 0.f ? acc * ep.scale : 0.f);
        } break;
        case EPI_EMBRACE_BWD: {
            float ev = load_act(ep.e, ep.out_dtype, (size_t)m * ep.ld_e + n);
            int id = ep.idx[(size_t)m * N + n];
            float g = ev > 0.f ? acc : 0.f;
            store_act(ep.out, ep.out_dtype, (size_t)m * ep.ldo + n, id == 0 ? g : 0.f);
            store_act(ep.out2, ep.out_dtype, (size_t)m * ep.ldo + n, id == 1 ? g : 0.f);
        } break;
        default: {  // EPI_ATOMIC
            size_t idx;
            if (ep.map == MAP_ROWMAJOR) idx = (size_t)m * ep.ldo + n;
            else if (ep.map == MAP_CONV_W) {  // n = (tap, c) -> dW[m][c][tap]
                int tap = n / ep.mapC, c = n - tap * ep.mapC;
                idx = ((size_t)m * ep.mapC + c) * ep.map_taps + tap;
            } else if (n < ep.map_off) {  // MAP_W_PERM, un-permuted prefix
                idx = (size_t)m * ep.map_wrows + n;
            } else {  // MAP_W_PERM: n = off + (l, c) -> dW[m][off + c*L + l]
                int n2 = n - ep.map_off, l = n2 / ep.mapC, c = n2 - l * ep.mapC;
                idx = (size_t)m * ep.map_wrows + ep.map_off + (size_t)c * ep.mapL + l;
            }
            atomicAdd((float*)ep.out + idx, acc);
        } break;
    }
}

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16, SG_THREADS = 256;

// a_kfast / b_kfast: which index consecutive threads walk when filling the smem tile (the one that is
// contiguous in memory for that operand, so the global loads coalesce).
__global__ void __launch_bounds__(SG_THREADS)
gemm_simt_kernel(Operand A, Operand B, Epilogue ep, int M, int N, int K, int k_chunk, int a_kfast, int b_kfast) {
    __shared__ float As[SG_BK][SG_BM + 4];
    __shared__ float Bs[SG_BK][SG_BN + 4];
    const int t = threadIdx.x;
    const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
    const int k_begin = blockIdx.z * k_chunk;
    const int k_end = min(K, k_begin + k_chunk);
    const int ty = t / 16, tx = t % 16;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = k_begin; k0 < k_end; k0 += SG_BK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int kk, mm;
            if (a_kfast) { kk = t % SG_BK; mm = t / SG_BK + 16 * i; }
            else { mm = t % SG_BM; kk = t / SG_BM + 4 * i; }
            int kg = k0 + kk;
            As[kk][mm] = (kg < k_end) ? operand_fetch(A, m0 + mm, kg) : 0.f;
            int kb, nn;
            if (b_kfast) { kb = t % SG_BK; nn = t / SG_BK + 16 * i; }
            else { nn = t % SG_BN; kb = t / SG_BN + 4 * i; }
            int kgb = k0 + kb;
            Bs[kb][nn] = (kgb < k_end) ? operand_fetch(B, n0 + nn, kgb) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SG_BK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n < N) epilogue_apply(ep, m, n, M, N, acc[i][j]);
        }
    }
}

inline Operand make_operand(const void* p, int dtype, int mode, int ld, int rows, int cols) {
    Operand o = {};
    o.p = p; o.dtype = dtype; o.mode = mode; o.ld = ld; o.rows = rows; o.cols = cols;
    o.sign = 1;
    return o;
}

inline bool operand_kfast(const Operand& o) {
    switch (o.mode) {
        case OP_ROWMAJOR: case OP_CONV_SHIFT: case OP_FLAT_ACT: case OP_CONV_W_DGRAD: return true;
        default: return false;  // transposed-like: rows are contiguous
    }
}

// split_k > 1 is only legal with EPI_ATOMIC
inline cudaError_t launch_gemm_simt(const Operand& A, const Operand& B, const Epilogue& ep, int M, int N, int K,
                                    int split_k, cudaStream_t st) {
    if (M <= 0 || N <= 0) return cudaSuccess;
    if (split_k < 1) split_k = 1;
    int k_chunk = round_up(cdiv(K, split_k), SG_BK);
    split_k = cdiv(K, k_chunk);
    if (split_k < 1) split_k = 1;
    dim3 grid(cdiv(N, SG_BN), cdiv(M, SG_BM), split_k);
    gemm_simt_kernel<<<grid, SG_THREADS, 0, st>>>(A, B, ep, M, N, K, k_chunk, operand_kfast(A), operand_kfast(B));
    return cudaGetLastError();
}

}  // namespace emb
