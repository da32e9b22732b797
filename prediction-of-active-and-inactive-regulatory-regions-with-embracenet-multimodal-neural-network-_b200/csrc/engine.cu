// libembrace_sm100.so -- planning, orchestration and the C ABI (include/embrace_b200.h).
//
// Path rebuilt here (reference file:line):
//   EmbraceNetMultimodal.forward   BIOINF_tesi/models/EmbraceNetMultimodal.py:159-193
//   EmbraceNet.forward             BIOINF_tesi/models/EmbraceNetMultimodal.py:34-90
//   FFNN_pre / CNN_pre             BIOINF_tesi/models/FFNN_pre.py:10-49, CNN_pre.py:12-76
//   FFNN / CNN (single modality)   BIOINF_tesi/models/FF_net.py:8-50, CNN_net.py:10-83
//   loss / metrics / optimizer     BIOINF_tesi/models/utils/training_models_multimodal.py:132-162, utils.py:80-140
#include <cstdarg>
#include <cstring>
#include <cmath>
#include <vector>
#include <type_traits>
#include <string>

#include "../../include/embrace_b200.h"
#include "common.cuh"
#include "kernels.cuh"
#include "kernels_fast.cuh"
#include "kernels_tma.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "conv_tc.cuh"
#include "onehot_wgrad_tc.cuh"
#include "probe.cuh"
#include "dp_peer.cuh"
#include "conv_pool_tc.cuh"
#include "onehot_pool_tc.cuh"

namespace emb {

thread_local std::string g_last_error;

int set_error(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

struct Tensor {       // one state_dict entry
    std::string name;
    int64_t offset, numel;
    int ndim, shape[3];
    bool is_buffer;
};

struct LinearLayer {
    int in, out;
    bf16* wc;         // bf16 operand copy [out][round_up(in,8)] (permuted to channels-last order when perm_in)
    float* wg_tmp;    // perm_in layers: weight gradient in engine (channels-last) column order before un-permuting
    int64_t w, b;     // offsets into the params arena
    float drop;
    bool relu;
    bool perm_in;     // weight columns follow the reference's flatten order c*L + l while the engine's activation is l*C + c
    bool flat_in;     // the input operand IS the in-place channels-last CNN activation (docking_1, last_layer1)
    int perm_off;     // leading weight columns that are NOT permuted (ConcatNet: the FFNN part of the concatenation)
};

struct ConvLayer {
    int cin, cout, k, pad, Lc, Lp, ld;
    int64_t w, b, gamma, beta;   // params arena
    int64_t rm, rv;              // buffers arena
    float drop;
    bf16* wc;                    // bf16 operand copy [k][cout][round_up(cin,8)]
    float* wg_tmp;               // weight gradient accumulated as [k][cout][cin] (coalesced atomics) before the permuted write-back
    // workspace
    void *y, *a, *dy, *ga;
    uint8_t* amax;               // [B, Lp, cout] where each pooled element's gradient goes (written by the forward pooling kernel)
    float *scale, *shift, *mean, *rstd;
    double *stats, *bstats;
};

struct Act {        // [rows, width] activation with leading dimension ld
    void* p = nullptr;
    int width = 0, ld = 0;
};

}  // namespace emb

using namespace emb;

struct EmbEngine {
    EmbArchSpec spec;
    int max_batch, prec, esize;      // esize: bytes per activation element
    int device = -1;                 // CUDA ordinal of the bound memory; every entry makes it current (emb_bind records it)
    bool use_tc = false;
    std::vector<Tensor> tensors;
    int64_t n_params = 0, n_buffers = 0;
    std::vector<LinearLayer> ffnn, post, head;   // head: FFNN final Linear / CNN last_layer1..last_output
    LinearLayer dock0{}, dock1{};
    std::vector<ConvLayer> cnn;
    int ffnn_out = 0, cnn_out = 0, cnn_Lp_last = 0, cnn_C_last = 0, cnn_ld_last = 0;

    // bound memory
    float *params = nullptr, *grads = nullptr, *buffers = nullptr, *opt_m = nullptr, *opt_v = nullptr;
    char* ws = nullptr;
    int64_t ws_bytes = 0;
    bool owns_memory = false;

    // workspace carve-out
    Act x0, d0, e, dd0, dd1;
    Act xcat, gcat;                  // ConcatNet: cat(FFNN out, flattened CNN out) in engine column order, and its gradient
    std::vector<Act> ffnn_h, ffnn_g, post_h, post_g, head_h, head_g;
    Act gflat;                       // gradient w.r.t. the flattened CNN output (== cnn.back().ga)
    uint8_t* idx = nullptr;
    double* cum0 = nullptr;
    // zero-before-use buffers, carved contiguously so that one memset per phase clears them
    char *zero_fwd = nullptr, *zero_bwd = nullptr;        // [BatchNorm stats of all layers] / [backward stats + weight-gradient scratch]
    int64_t zero_fwd_bytes = 0, zero_bwd_bytes = 0;
    void* wc_table = nullptr;                              // device table (WcEntry[]) of the fused weight-cache refresh
    int wc_entries = 0;
    int64_t wc_total = 0;
    float *logits = nullptr, *dlogits = nullptr, *probs = nullptr;
    float* in_x = nullptr;           // staging for the *_host entries
    uint8_t* in_bases = nullptr;
    int32_t* in_labels = nullptr;
    float* in_avail = nullptr;
    RngState* rng = nullptr;
    StepMetricsDev* rec = nullptr;
    OptScalars* opt_scalars = nullptr;
    int* rec_count = nullptr;
    static const int MAX_REC = 1 << 16;

    // per-forward state needed by backward
    int last_B = 0;
    bool last_training = false;
    const uint8_t* last_bases = nullptr;

    // optimizer / shard state
    int64_t opt_t = 0;
    double nadam_mu_product = 1.0;
    int64_t row_offset = 0, global_batch = -1, global_pos = -1;
    uint64_t seed = 0x5EEDull;
    EmbAllreduceFn allreduce = nullptr;
    void* allreduce_user = nullptr;
    EmbPhaseFn phase_hook = nullptr;
    void* phase_user = nullptr;
    int64_t launches = 0;
    // Independent branches of the step run on side streams (fork / join through events; inside a captured step they become parallel
    // branches of the CUDA graph): the FFNN chain next to the CNN chain, every weight gradient next to the data-gradient chain.
    // At small batch the step is a chain of latency-bound launches, and this halves its critical path.
    cudaStream_t side[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> fork_ev;
    size_t fork_next = 0;
    bool fork_on = true;
    // peer-memory data parallelism (emb_dp_attach, csrc/dp_peer.cuh): SyncBN / loss-weight / gradient exchanges as kernels
    bool dp_on = false;
    DpCtx dp{};
    DpArenas dp_ar{};
    unsigned int* dp_epoch = nullptr;      // device counter of training steps (workspace)
    int64_t* dp_gpos = nullptr;            // device slot: global positive count of the current step (workspace)
    int64_t dp_lo = 0, dp_hi = 0;          // this rank's slice of the parameter arena (optimizer state and work are sharded)
    // CUDA-graph replay of the whole train step (emb_set_graph): one instantiated graph per batch size
    struct StepGraph { int B; int has_opt; int opt_kind; int64_t kernels; cudaGraphExec_t exec; };
    bool graph_on = false;
    bool graph_coll = false;         // emb_set_graph(2): the data-parallel hooks (SyncBN / gradient all-reduce callbacks) are captured too
    int64_t* gpos_dev = nullptr;     // global positive count of the batch for the captured loss kernel (stream-ordered copy per step)
    std::vector<StepGraph> graphs;
    std::vector<int> graph_seen;     // batch sizes that already ran once eagerly (lazy initialisation happens there)
    bool capturing = false;
    // pipelined host entry (emb_train_step_host_pipelined): two device staging slots filled by a copy stream one step ahead
    float* stg_x[2] = {nullptr, nullptr};
    uint8_t* stg_bases[2] = {nullptr, nullptr};
    int32_t* stg_labels[2] = {nullptr, nullptr};
    float* stg_avail[2] = {nullptr, nullptr};
    int64_t infer_calls = 0;                   // emb_predict_host_pipelined: calls so far / a call whose scores are not yet confirmed
    bool infer_pending = false;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t infer_copy = nullptr;         // emb_predict_host_pipelined: uploads of the next batch
    cudaEvent_t infer_ev[4] = {};              // [slot] upload done, [2 + slot] batch complete
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr}, ev_step[2] = {nullptr, nullptr}, ev_d2h = nullptr;
    EmbStepMetrics* pin_metrics = nullptr;      // pinned host landing slot of the previous step's record
    int64_t pipe_steps = 0;                     // steps enqueued so far
    bool pipe_pending = false;                  // a step whose metrics have not been delivered yet
    cudaStream_t gstream = nullptr;  // graphs are captured and launched on an engine-owned stream (the caller's may be the legacy
    cudaEvent_t gev_in = nullptr, gev_out = nullptr;   // default stream, which cannot be captured); events order it with the caller's
    // per-kernel timing of the GEMM class (bench.py roofline): CUDA event pairs around each launch
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;
    size_t prof_used = 0;
    double prof_flops = 0;
    std::vector<double> prof_flops_each;
    int64_t prof_launches = 0;
};

namespace {

int dtype_of(const EmbEngine* e) { return e->prec == EMB_PREC_BF16 ? 1 : 0; }

void add_tensor(EmbEngine* e, const std::string& name, std::initializer_list<int> shape, bool is_buffer, int64_t* off) {
    Tensor t;
    t.name = name;
    t.ndim = (int)shape.size();
    t.numel = 1;
    int i = 0;
    t.shape[0] = t.shape[1] = t.shape[2] = 1;
    for (int s : shape) { t.shape[i++] = s; t.numel *= s; }
    t.is_buffer = is_buffer;
    int64_t& cursor = is_buffer ? e->n_buffers : e->n_params;
    t.offset = cursor;
    cursor += round_up64(t.numel, 4);   // keep every tensor 16-byte aligned in the arena
    *off = t.offset;
    e->tensors.push_back(t);
}

LinearLayer add_linear(EmbEngine* e, const std::string& prefix, int in, int out, float drop, bool relu, bool perm_in) {
    LinearLayer l{};
    l.in = in; l.out = out; l.drop = drop; l.relu = relu; l.perm_in = perm_in; l.flat_in = perm_in; l.perm_off = 0;
    add_tensor(e, prefix + ".weight", {out, in}, false, &l.w);
    add_tensor(e, prefix + ".bias", {out}, false, &l.b);
    return l;
}

int plan(EmbEngine* e) {
    const EmbArchSpec& s = e->spec;
    const bool cat_kind = s.kind == EMB_KIND_CONCATNET;
    const bool emb_kind = s.kind == EMB_KIND_EMBRACENET || cat_kind;     // both carry the FFNN. / CNN. prefixes
    const std::string pf = emb_kind ? "FFNN.model." : "model.";
    const std::string pc = emb_kind ? "CNN.CNN_model." : "CNN_model.";
    if (s.kind != EMB_KIND_CNN) {
        if (s.n_ffnn < 1 || s.n_ffnn > EMB_MAX_FFNN || s.in_features < 1) return set_error(EMB_E_ARG, "bad FFNN spec");
        int in = s.in_features;
        for (int i = 0; i < s.n_ffnn; ++i) {
            if (s.ffnn_units[i] < 1 || s.ffnn_dropout[i] < 0 || s.ffnn_dropout[i] >= 1) return set_error(EMB_E_ARG, "bad FFNN layer %d", i);
            e->ffnn.push_back(add_linear(e, pf + std::to_string(3 * i), in, s.ffnn_units[i], s.ffnn_dropout[i], true, false));
            in = s.ffnn_units[i];
        }
        e->ffnn_out = in;
        if (s.kind == EMB_KIND_FFNN) e->head.push_back(add_linear(e, pf + std::to_string(3 * s.n_ffnn), in, 2, 0.f, false, false));
    }
    if (s.kind != EMB_KIND_FFNN) {
        if (s.n_cnn < 1 || s.n_cnn > EMB_MAX_CNN) return set_error(EMB_E_ARG, "bad CNN spec");
        int cin = 4, L = SEQ_LEN;
        for (int i = 0; i < s.n_cnn; ++i) {
            ConvLayer c{};
            c.cin = cin; c.cout = s.cnn_channels[i]; c.k = s.cnn_kernels[i]; c.drop = s.cnn_dropout[i];
            if (c.cout < 1 || c.k < 1 || (c.k & 1) == 0 || c.drop < 0 || c.drop >= 1) return set_error(EMB_E_ARG, "bad CNN layer %d (odd kernel sizes only)", i);
            if (i == 0 && c.cout * c.k > OHB_MAXP * 256) return set_error(EMB_E_UNSUPPORTED, "first conv: out_channels*kernel_size > %d", OHB_MAXP * 256);
            c.pad = (c.k - 1) / 2;
            c.Lc = (L + 2 * c.pad - c.k) + 1;                 // size_out_convolution, stride 1
            c.Lp = (c.Lc - POOL_K) / POOL_S + 1;              // size_out_convolution(.., 10, 0, 2)
            if (c.Lc < POOL_K) return set_error(EMB_E_ARG, "sequence too short at CNN layer %d", i);
            c.ld = round_up(c.cout, 8);
            add_tensor(e, pc + std::to_string(5 * i) + ".weight", {c.cout, c.cin, c.k}, false, &c.w);
            add_tensor(e, pc + std::to_string(5 * i) + ".bias", {c.cout}, false, &c.b);
            add_tensor(e, pc + std::to_string(5 * i + 1) + ".weight", {c.cout}, false, &c.gamma);
            add_tensor(e, pc + std::to_string(5 * i + 1) + ".bias", {c.cout}, false, &c.beta);
            add_tensor(e, pc + std::to_string(5 * i + 1) + ".running_mean", {c.cout}, true, &c.rm);
            add_tensor(e, pc + std::to_string(5 * i + 1) + ".running_var", {c.cout}, true, &c.rv);
            e->cnn.push_back(c);
            cin = c.cout;
            L = c.Lp;
        }
        e->cnn_Lp_last = L;
        e->cnn_C_last = cin;
        e->cnn_ld_last = e->cnn.back().ld;
        e->cnn_out = cin * L;
        if (s.kind == EMB_KIND_CNN) {   // CNN_net.py:71-81: three Linear layers, no activation in between
            e->head.push_back(add_linear(e, "last_layer1", e->cnn_out, 1000, 0.f, false, true));
            e->head.push_back(add_linear(e, "last_layer2", 1000, 64, 0.f, false, false));
            e->head.push_back(add_linear(e, "last_output", 64, 2, 0.f, false, false));
        }
    }
    if (cat_kind) {
        // ConcatNetMultimodal.py:36-62: post = (Linear -> ReLU -> Dropout) x n (1..3) -> Linear(-> 2) over cat(FFNN, CNN)
        if (s.n_post < 1 || s.n_post > EMB_MAX_POST) return set_error(EMB_E_ARG, "ConcatNet needs 1..%d post layers", EMB_MAX_POST);
        int in = e->ffnn_out + e->cnn_out;
        for (int i = 0; i < s.n_post; ++i) {
            e->post.push_back(add_linear(e, "post." + std::to_string(3 * i), in, s.post_units[i], s.post_dropout[i], true, false));
            if (i == 0) { e->post[0].perm_in = true; e->post[0].perm_off = e->ffnn_out; }     // the CNN columns are in c*L + l order
            in = s.post_units[i];
        }
        e->head.push_back(add_linear(e, "post." + std::to_string(3 * s.n_post), in, 2, 0.f, false, false));
    } else if (emb_kind) {
        const int C = s.embracement_size;
        if (C < 1 || s.n_post < 0 || s.n_post > EMB_MAX_POST || s.p_ffnn < 0 || s.p_ffnn > 1) return set_error(EMB_E_ARG, "bad EmbraceNet spec");
        e->dock0 = add_linear(e, "embracenet.docking_0", e->ffnn_out, C, 0.f, true, false);
        e->dock1 = add_linear(e, "embracenet.docking_1", e->cnn_out, C, 0.f, true, true);
        int in = C;
        for (int i = 0; i < s.n_post; ++i) {
            e->post.push_back(add_linear(e, "post." + std::to_string(3 * i), in, s.post_units[i], s.post_dropout[i], true, false));
            in = s.post_units[i];
        }
        e->head.push_back(add_linear(e, "post." + std::to_string(3 * s.n_post), in, 2, 0.f, false, false));
    }
    return EMB_OK;
}

struct Bump {
    char* base;
    int64_t off = 0;
    template <typename T> T* take(int64_t count) {
        off = round_up64(off, 256);
        T* p = base ? (T*)(base + off) : nullptr;
        off += count * (int64_t)sizeof(T);
        return p;
    }
    void* take_bytes(int64_t bytes) { return (void*)take<char>(bytes); }
};

Act take_act(Bump& bp, int64_t rows, int width, int esize) {
    Act a;
    a.width = width;
    a.ld = round_up(width, 8);
    a.p = bp.take_bytes(rows * a.ld * esize);
    return a;
}

// bf16 GEMM-operand copies of the fp32 master weights, refreshed at the start of every forward
__global__ void wcache_linear_kernel(const float* __restrict__ w, bf16* __restrict__ out, int N, int K, int ldk, int perm, int L, int C) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)N * ldk) return;
    int n = i / ldk, kk = i - (size_t)n * ldk;
    float v = 0.f;
    if (kk < K) {
        int src = kk;
        if (perm) { int l = kk / C, c = kk - l * C; src = c * L + l; }
        v = w[(size_t)n * K + src];
    }
    out[i] = __float2bfloat16_rn(v);
}
__global__ void wcache_conv_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int Cin, int taps, int ldc) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)taps * Cout * ldc) return;
    int c = i % ldc;
    size_t r = i / ldc;
    int o = r % Cout, tap = r / Cout;
    out[i] = __float2bfloat16_rn(c < Cin ? w[((size_t)o * Cin + c) * taps + tap] : 0.f);
}

// All bf16 operand copies in ONE launch: a small device table lists the tensors; a thread finds its entry by a linear scan
// over the (<= 32) start offsets.  kind 0: Linear [N][ldk] (optionally permuted to channels-last input order), 1: Conv [taps][Cout][ldc].
struct WcEntry {
    const float* src;
    bf16* dst;
    unsigned long long start, count;
    int kind, N, K, ldk, perm, L, C, taps;
};
__global__ void wcache_fused_kernel(const WcEntry* __restrict__ tab, int n_entries, unsigned long long total) {
    __shared__ WcEntry s_tab[32];
    for (int i = threadIdx.x; i < n_entries; i += blockDim.x) s_tab[i] = tab[i];
    __syncthreads();
    for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (unsigned long long)gridDim.x * blockDim.x) {
        int ei = 0;
        while (ei + 1 < n_entries && g >= s_tab[ei + 1].start) ++ei;
        const WcEntry& t = s_tab[ei];
        const size_t i = (size_t)(g - t.start);
        float v = 0.f;
        if (t.kind == 0) {
            const int n = (int)(i / t.ldk), kk = (int)(i - (size_t)n * t.ldk);
            if (kk < t.K) {
                int srck = kk;
                if (t.perm && kk >= t.taps) { const int k2 = kk - t.taps, l = k2 / t.C, c = k2 - l * t.C; srck = t.taps + c * t.L + l; }
                v = t.src[(size_t)n * t.K + srck];
            }
        } else {      // conv: N = Cout, K = Cin, ldk = ldc
            const int c = (int)(i % t.ldk);
            const size_t r = i / t.ldk;
            const int o = (int)(r % t.N), tap = (int)(r / t.N);
            if (c < t.K) v = t.src[((size_t)o * t.K + c) * t.taps + tap];
        }
        t.dst[i] = __float2bfloat16_rn(v);
    }
}

// dW[n][c*L + l] = tmp[n][l*C + c]: the permuted-input Linear's weight gradient back to the reference's flatten order
__global__ void unpermute_wgrad_kernel(const float* __restrict__ tmp, float* __restrict__ dw, int N, int L, int C, int off) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t K = (size_t)off + (size_t)L * C;       // `off` leading columns are not permuted (ConcatNet)
    if (i >= (size_t)N * K) return;
    size_t n = i / K;
    int r = (int)(i - n * K);
    if (r < off) { dw[i] = tmp[i]; return; }
    r -= off;
    int c = r / L, l = r - c * L;                        // destination index walks the reference layout (coalesced writes)
    dw[i] = tmp[n * K + off + (size_t)l * C + c];
}

// dW[o][c][tap] = tmp[tap][o][c]: the conv weight gradient back to the reference's [Cout, Cin, k] layout
__global__ void unpermute_conv_wgrad_kernel(const float* __restrict__ tmp, float* __restrict__ dw, int Cout, int Cin, int taps) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t oc = (size_t)Cout * Cin;
    if (i >= oc * taps) return;
    const size_t r = i / taps;                           // (o, c); destination index walks the reference layout
    const int tap = (int)(i - r * taps);
    dw[i] = tmp[(size_t)tap * oc + r];
}

// ---------------------------------------------------------------------------------------------
// The 2-logit head (post.N / FFNN final Linear / CNN last_output: nn.Linear(K -> 2)).  A [B, K] x [K, 2] product is a
// pair of dot products per row -- far below any GEMM tile -- so it gets three small bandwidth-shaped kernels instead of
// the generic SIMT GEMM (which took 25-55 us per call at batch 8192).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
head_fwd_kernel(const T* __restrict__ x, int ld, const float* __restrict__ w, const float* __restrict__ bias, int round_w,
                float* __restrict__ logits, int B, int K) {
    extern __shared__ float hw[];                       // [2][K]
    for (int i = threadIdx.x; i < 2 * K; i += blockDim.x) hw[i] = round_w ? __bfloat162float(__float2bfloat16_rn(w[i])) : w[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int b = blockIdx.x * 8 + warp; b < B; b += gridDim.x * 8) {
        float a0 = 0.f, a1 = 0.f;
        for (int k = lane; k < K; k += 32) {
            const float v = to_f(x[(size_t)b * ld + k]);
            a0 = fmaf(v, hw[k], a0);
            a1 = fmaf(v, hw[K + k], a1);
        }
        a0 = warp_sum(a0);
        a1 = warp_sum(a1);
        if (lane == 0) { logits[2 * b] = a0 + bias[0]; logits[2 * b + 1] = a1 + bias[1]; }
    }
}

// dW[j][k] += sum_b dl[b][j] * x[b][k],  db[j] += sum_b dl[b][j]        (dl fp32 [B, 2])
template <typename T>
__global__ void __launch_bounds__(256)
head_wgrad_kernel(const float* __restrict__ dl, const T* __restrict__ x, int ld, float* __restrict__ dw, float* __restrict__ db, int B, int K,
                  int rows_per_block) {
    const int b0 = blockIdx.x * rows_per_block, b1 = min(B, b0 + rows_per_block);
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        float a0 = 0.f, a1 = 0.f;
        for (int b = b0; b < b1; ++b) {
            const float v = to_f(x[(size_t)b * ld + k]);
            const float2 d = *reinterpret_cast<const float2*>(dl + 2 * b);
            a0 = fmaf(d.x, v, a0);
            a1 = fmaf(d.y, v, a1);
        }
        atomicAdd(&dw[k], a0);
        atomicAdd(&dw[K + k], a1);
    }
    if (threadIdx.x < 2) {
        float a = 0.f;
        for (int b = b0; b < b1; ++b) a += dl[2 * b + threadIdx.x];
        atomicAdd(&db[threadIdx.x], a);
    }
}

// acc[b][k] = dl[b][0] * W[0][k] + dl[b][1] * W[1][k], finished by the usual epilogue functor
__global__ void __launch_bounds__(256)
head_dgrad_kernel(const float* __restrict__ dl, const float* __restrict__ w, int round_w, Epilogue ep, int B, int K) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * K) return;
    const int b = (int)(i / K), k = (int)(i - (size_t)b * K);
    float w0 = w[k], w1 = w[K + k];
    if (round_w) { w0 = __bfloat162float(__float2bfloat16_rn(w0)); w1 = __bfloat162float(__float2bfloat16_rn(w1)); }
    epilogue_apply(ep, b, k, B, K, fmaf(dl[2 * b], w0, dl[2 * b + 1] * w1));
}

// ConcatNet (ConcatNetMultimodal.py:74): cat(FFNN out [B, fo], flattened CNN out) in engine column order (l*C + c)
template <typename T>
__global__ void concat_kernel(const T* __restrict__ f, int ldf, int fo, const T* __restrict__ a, int Lp, int C, int lda,
                              T* __restrict__ out, int ldo, int B) {
    const size_t K = (size_t)fo + (size_t)Lp * C;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * K) return;
    const size_t b = i / K;
    const int k = (int)(i - b * K);
    if (k < fo) { out[b * ldo + k] = f[b * ldf + k]; return; }
    const int k2 = k - fo, l = k2 / C, c = k2 - l * C;
    out[b * ldo + k] = a[(b * Lp + l) * lda + c];
}
// ... and its backward: the FFNN part goes through the ReLU / Dropout mask of the last FFNN layer, the CNN part is the
// gradient of the flattened pooled output
template <typename T>
__global__ void concat_split_kernel(const T* __restrict__ g, int ldg, int fo, const T* __restrict__ href, int ldh, float scale,
                                    T* __restrict__ gf, int ldgf, T* __restrict__ ga, int Lp, int C, int lda, int B) {
    const size_t K = (size_t)fo + (size_t)Lp * C;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * K) return;
    const size_t b = i / K;
    const int k = (int)(i - b * K);
    const float v = to_f(g[b * ldg + k]);
    if (k < fo) { gf[b * ldgf + k] = from_f<T>(to_f(href[b * ldh + k]) > 0.f ? v * scale : 0.f); return; }
    const int k2 = k - fo, l = k2 / C, c = k2 - l * C;
    ga[(b * Lp + l) * lda + c] = from_f<T>(v);
}

// carve the workspace; with base == nullptr only the size is computed
int64_t carve(EmbEngine* e, char* base) {
    Bump bp{base};
    const int64_t Bm = e->max_batch;
    const int es = e->esize;
    const EmbArchSpec& s = e->spec;
    e->ffnn_h.clear(); e->ffnn_g.clear(); e->post_h.clear(); e->post_g.clear(); e->head_h.clear(); e->head_g.clear();
    e->rng = bp.take<RngState>(1);
    e->rec_count = bp.take<int>(1);
    e->rec = bp.take<StepMetricsDev>(EmbEngine::MAX_REC);
    e->opt_scalars = bp.take<OptScalars>(1);
    e->dp_epoch = bp.take<unsigned int>(1);
    e->dp_gpos = bp.take<int64_t>(1);
    e->logits = bp.take<float>(Bm * 2);
    e->dlogits = bp.take<float>(Bm * 2);
    e->probs = bp.take<float>(Bm);
    e->in_x = bp.take<float>(Bm * std::max(1, s.in_features));
    e->in_bases = bp.take<uint8_t>(Bm * SEQ_LEN);
    e->in_labels = bp.take<int32_t>(Bm);
    e->in_avail = bp.take<float>(Bm * 2);
    for (int k = 0; k < 2; ++k) {
        e->stg_x[k] = bp.take<float>(Bm * std::max(1, s.in_features));
        e->stg_bases[k] = bp.take<uint8_t>(Bm * SEQ_LEN);
        e->stg_labels[k] = bp.take<int32_t>(Bm);
        e->stg_avail[k] = bp.take<float>(Bm * 2);
    }
    if (e->prec == EMB_PREC_BF16) {
        auto wl = [&](LinearLayer& l) {
            l.wc = bp.take<bf16>((int64_t)l.out * round_up(l.in, 8));
        };
        for (auto& l : e->ffnn) wl(l);
        for (auto& l : e->post) wl(l);
        for (auto& l : e->head) wl(l);
        if (s.kind == EMB_KIND_EMBRACENET) { wl(e->dock0); wl(e->dock1); }
        for (size_t i = 1; i < e->cnn.size(); ++i) {
            e->cnn[i].wc = bp.take<bf16>((int64_t)e->cnn[i].k * e->cnn[i].cout * round_up(e->cnn[i].cin, 8));
        }
    }
    if (s.kind != EMB_KIND_CNN) {
        e->x0 = take_act(bp, Bm, s.in_features, es);
        for (auto& l : e->ffnn) {
            e->ffnn_h.push_back(take_act(bp, Bm, l.out, es));
            e->ffnn_g.push_back(take_act(bp, Bm, l.out, es));
        }
    }
    for (auto& c : e->cnn) {
        c.y = bp.take_bytes(Bm * c.Lc * c.ld * es);
        c.dy = bp.take_bytes(Bm * c.Lc * c.ld * es);
        c.a = bp.take_bytes(Bm * c.Lp * c.ld * es);
        c.ga = bp.take_bytes(Bm * c.Lp * c.ld * es);
        c.amax = bp.take<uint8_t>(Bm * c.Lp * c.ld + 64);
        c.scale = bp.take<float>(c.cout);
        c.shift = bp.take<float>(c.cout);
        c.mean = bp.take<float>(c.cout);
        c.rstd = bp.take<float>(c.cout);
    }
    {   // the forward's zero region: BatchNorm sum / sum-of-squares of every layer
        bp.off = round_up64(bp.off, 256);
        const int64_t z0 = bp.off;
        e->zero_fwd = base ? base + z0 : nullptr;
        for (auto& c : e->cnn) c.stats = bp.take<double>(2 * c.cout);
        e->zero_fwd_bytes = round_up64(bp.off, 256) - z0;
        // the backward's zero region: backward BatchNorm sums and the weight-gradient scratch in operand layout
        bp.off = round_up64(bp.off, 256);
        const int64_t z1 = bp.off;
        e->zero_bwd = base ? base + z1 : nullptr;
        for (auto& c : e->cnn) c.bstats = bp.take<double>(2 * c.cout);
        if (e->prec == EMB_PREC_BF16) {
            for (size_t i = 1; i < e->cnn.size(); ++i) e->cnn[i].wg_tmp = bp.take<float>((int64_t)e->cnn[i].k * e->cnn[i].cout * e->cnn[i].cin);
            auto wt = [&](LinearLayer& l) { if (l.perm_in) l.wg_tmp = bp.take<float>((int64_t)l.out * l.in); };
            for (auto& l : e->head) wt(l);
            for (auto& l : e->post) wt(l);
            if (s.kind == EMB_KIND_EMBRACENET) wt(e->dock1);
        }
        e->zero_bwd_bytes = round_up64(bp.off, 256) - z1;
        e->wc_table = bp.take_bytes(32 * 128);
    }
    if (s.kind == EMB_KIND_CONCATNET) {
        e->xcat = take_act(bp, Bm, e->ffnn_out + e->cnn_out, es);
        e->gcat = take_act(bp, Bm, e->ffnn_out + e->cnn_out, es);
        for (auto& l : e->post) {
            e->post_h.push_back(take_act(bp, Bm, l.out, es));
            e->post_g.push_back(take_act(bp, Bm, l.out, es));
        }
    }
    if (s.kind == EMB_KIND_EMBRACENET) {
        const int C = s.embracement_size;
        e->d0 = take_act(bp, Bm, C, es);
        e->e = take_act(bp, Bm, C, es);
        e->dd0 = take_act(bp, Bm, C, es);
        e->dd1 = take_act(bp, Bm, C, es);
        e->idx = bp.take<uint8_t>(Bm * C);
        e->cum0 = bp.take<double>(Bm);
        for (auto& l : e->post) {
            e->post_h.push_back(take_act(bp, Bm, l.out, es));
            e->post_g.push_back(take_act(bp, Bm, l.out, es));
        }
    }
    if (s.kind == EMB_KIND_CNN)
        for (size_t i = 0; i + 1 < e->head.size(); ++i) {
            e->head_h.push_back(take_act(bp, Bm, e->head[i].out, es));
            e->head_g.push_back(take_act(bp, Bm, e->head[i].out, es));
        }
    return round_up64(bp.off, 256);
}

// ---------------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------------
#define LAUNCHED(e) ((e)->launches++)

Epilogue base_epi(EmbEngine* e, int mode, void* out, int ldo) {
    Epilogue ep = {};
    ep.mode = mode;
    ep.out = out;
    ep.out_dtype = dtype_of(e);
    ep.ldo = ldo;
    ep.rng = e->rng;
    ep.row_offset = e->row_offset;
    ep.scale = 1.f;
    return ep;
}

Operand weight_operand(EmbEngine* e, const LinearLayer& l, bool transposed) {
    // forward: rows = out units, k = in;  transposed (dgrad): rows = in, k = out
    Operand o;
    if (!l.perm_in) {
        o = transposed ? make_operand(e->params + l.w, 0, OP_TRANSPOSED, l.in, l.in, l.out)
                       : make_operand(e->params + l.w, 0, OP_ROWMAJOR, l.in, l.out, l.in);
    } else {
        o = transposed ? make_operand(e->params + l.w, 0, OP_W_PERM_T, 0, l.in, l.out)
                       : make_operand(e->params + l.w, 0, OP_W_PERM, 0, l.out, l.in);
        o.L = e->cnn_Lp_last; o.C = e->cnn_C_last; o.wrows = l.in; o.perm_off = l.perm_off;
    }
    o.round_bf16 = e->prec == EMB_PREC_BF16;
    return o;
}

// the input activation of a Linear layer as the A operand (rows = batch)
Operand input_operand(EmbEngine* e, const LinearLayer& l, const Act& in, int B, bool transposed) {
    Operand o;
    const int dt = dtype_of(e);
    if (!l.flat_in) {
        o = transposed ? make_operand(in.p, dt, OP_TRANSPOSED, in.ld, l.in, B) : make_operand(in.p, dt, OP_ROWMAJOR, in.ld, B, l.in);
    } else {
        o = transposed ? make_operand(in.p, dt, OP_FLAT_ACT_T, e->cnn_ld_last, l.in, B)
                       : make_operand(in.p, dt, OP_FLAT_ACT, e->cnn_ld_last, B, l.in);
        o.L = e->cnn_Lp_last; o.C = e->cnn_C_last;
    }
    return o;
}

void prof_begin(EmbEngine* e, double flops, cudaStream_t st) {
    if (!e->prof_on) return;
    if (e->prof_used + 2 > e->prof_ev.size()) {
        if (e->prof_ev.size() >= (1u << 16)) return;
        for (int i = 0; i < 512; ++i) { cudaEvent_t ev; cudaEventCreate(&ev); e->prof_ev.push_back(ev); }
    }
    cudaEventRecord(e->prof_ev[e->prof_used], st);
    e->prof_flops += flops;
    e->prof_flops_each.push_back(flops);
    e->prof_launches += 1;
}
void prof_end(EmbEngine* e, cudaStream_t st) {
    if (!e->prof_on || e->prof_used + 2 > e->prof_ev.size()) return;
    cudaEventRecord(e->prof_ev[e->prof_used + 1], st);
    e->prof_used += 2;
}

int run_gemm(EmbEngine* e, const Operand& A, const Operand& B, const Epilogue& ep, int M, int N, int K, int split_k, cudaStream_t st) {
    prof_begin(e, 2.0 * M * N * K, st);
    cudaError_t err = launch_gemm_simt(A, B, ep, M, N, K, split_k, st);
    prof_end(e, st);
    if (err != cudaSuccess) return set_error(EMB_E_CUDA, "gemm launch: %s", cudaGetErrorString(err));
    LAUNCHED(e);
    return EMB_OK;
}

int pick_split_k(int M, int N, int K) {
    if (tuning().deterministic) return 1;          // one contributor per output element: the atomic accumulation has a fixed order
    int tiles = cdiv(M, SG_BM) * cdiv(N, SG_BN);
    int want = std::max(1, (148 * 4) / std::max(1, tiles));
    int maxk = std::max(1, K / 64);
    return std::min(want, maxk);
}

int run_tc(EmbEngine* e, const TcProblem& pr, const Epilogue& ep, double flops, cudaStream_t st) {
    prof_begin(e, flops, st);
    int rc = tc_dispatch(pr, ep, st);
    prof_end(e, st);
    if (rc) return rc;
    LAUNCHED(e);
    return EMB_OK;
}

bool tc_on(const EmbEngine* e) { return e->use_tc && e->prec == EMB_PREC_BF16; }

// ---- fork / join of independent branches -------------------------------------------------------------------------------
// forking is off while per-launch GEMM timing runs (concurrent kernels would share the GPU and inflate each other's time) and in
// the host-callback data-parallel mode (its hooks assume one stream)
bool forking(const EmbEngine* e) { return e->fork_on && e->side[0] && !e->prof_on && !e->allreduce && !e->phase_hook && tuning().fork; }
// `to` waits for everything enqueued on `from` so far
int stream_after(EmbEngine* e, cudaStream_t from, cudaStream_t to) {
    if (from == to) return EMB_OK;
    cudaEvent_t ev = e->fork_ev[e->fork_next++ % e->fork_ev.size()];
    EMB_CUDA_OK(cudaEventRecord(ev, from));
    EMB_CUDA_OK(cudaStreamWaitEvent(to, ev, 0));
    return EMB_OK;
}

// can this Linear layer's GEMMs run on the tensor-core kernel?  (the flattened CNN input must be dense)
bool tc_linear_ok(const EmbEngine* e, const LinearLayer& l, int B = 1 << 30) {
    if (!tc_on(e) || l.out < 16 || l.in < 16) return false;
    // a persistent tcgen05 launch costs ~10 us before its first MMA retires (TMEM allocation, barrier set-up, pipeline fill): below
    // `tc_min_mflop` the SIMT kernel finishes first (small-batch steps, the narrow FFNN layers)
    if (2.0 * (double)B * l.out * l.in < 1e6 * (double)tuning().tc_min_mflop) return false;
    if (l.flat_in && e->cnn_ld_last != e->cnn_C_last) return false;
    return true;
}

int build_wcache_table(EmbEngine* e) {
    std::vector<WcEntry> tab;
    unsigned long long cursor = 0;
    auto wl = [&](const LinearLayer& l) {
        if (!l.wc) return;
        WcEntry t = {};
        t.src = e->params + l.w; t.dst = l.wc; t.kind = 0; t.N = l.out; t.K = l.in; t.ldk = round_up(l.in, 8);
        t.perm = l.perm_in ? 1 : 0; t.L = e->cnn_Lp_last; t.C = e->cnn_C_last; t.taps = l.perm_off;     // taps doubles as the un-permuted prefix
        t.start = cursor; t.count = (unsigned long long)l.out * t.ldk;
        cursor += t.count;
        tab.push_back(t);
    };
    for (auto& l : e->ffnn) wl(l);
    for (auto& l : e->post) wl(l);
    for (auto& l : e->head) wl(l);
    if (e->spec.kind == EMB_KIND_EMBRACENET) { wl(e->dock0); wl(e->dock1); }
    for (size_t i = 1; i < e->cnn.size(); ++i) {
        ConvLayer& c = e->cnn[i];
        if (!c.wc) continue;
        WcEntry t = {};
        t.src = e->params + c.w; t.dst = c.wc; t.kind = 1; t.N = c.cout; t.K = c.cin; t.ldk = round_up(c.cin, 8); t.taps = c.k;
        t.start = cursor; t.count = (unsigned long long)c.k * c.cout * t.ldk;
        cursor += t.count;
        tab.push_back(t);
    }
    if (tab.size() > 32) return set_error(EMB_E_UNSUPPORTED, "too many weight tensors for the fused weight-cache refresh");
    e->wc_entries = (int)tab.size();
    e->wc_total = (int64_t)cursor;
    if (!tab.empty()) EMB_CUDA_OK(cudaMemcpy(e->wc_table, tab.data(), tab.size() * sizeof(WcEntry), cudaMemcpyHostToDevice));
    return EMB_OK;
}

int refresh_wcache(EmbEngine* e, cudaStream_t st) {
    if (!tc_on(e) || e->wc_entries == 0) return EMB_OK;
    const int grid = (int)std::min<int64_t>(148 * 16, cdiv(e->wc_total, 256));
    wcache_fused_kernel<<<grid, 256, 0, st>>>((const WcEntry*)e->wc_table, e->wc_entries, (unsigned long long)e->wc_total);
    EMB_CHECK_LAUNCH();
    LAUNCHED(e);
    return EMB_OK;
}

bool tc_conv_ok(const EmbEngine* e, const ConvLayer& c) {
    return tc_on(e) && c.Lc <= 128 && c.cin >= 16 && c.cout >= 16 && (c.cin % 8) == 0 && (c.cout % 8) == 0;
}

// y = [dropout]([relu](x W^T + b))
int linear_forward(EmbEngine* e, const LinearLayer& l, const Act& in, const Act& out, int B, bool training,
                   const float* drop_u, uint32_t stream_id, cudaStream_t st) {
    Operand A = input_operand(e, l, in, B, false);
    Operand W = weight_operand(e, l, false);
    Epilogue ep = base_epi(e, EPI_LINEAR, out.p, out.ld);
    ep.bias = e->params + l.b;
    ep.relu = l.relu;
    if (training && l.drop > 0.f) { ep.drop_p = l.drop; ep.drop_u = drop_u; ep.rng_stream = stream_id; }
    if (tc_linear_ok(e, l, B)) {
        TcProblem pr = {};
        pr.kind = TC_LINEAR_FWD; pr.a = (const bf16*)in.p; pr.lda = in.ld; pr.b = l.wc; pr.ldb = round_up(l.in, 8);
        pr.M = B; pr.N = l.out; pr.K = l.in;
        return run_tc(e, pr, ep, 2.0 * B * l.out * l.in, st);
    }
    return run_gemm(e, A, W, ep, B, l.out, l.in, 1, st);
}

// dW += g^T x ; db += colsum(g)          (g: gradient w.r.t. the layer's pre-activation, [B, out])
int linear_wgrad(EmbEngine* e, const LinearLayer& l, const void* g, int g_dtype, int g_ld, const Act& in, int B, cudaStream_t st) {
    Operand A = make_operand(g, g_dtype, OP_TRANSPOSED, g_ld, l.out, B);
    Operand X = input_operand(e, l, in, B, true);
    Epilogue ep = base_epi(e, EPI_ATOMIC, e->grads + l.w, l.in);
    if (l.perm_in) { ep.map = MAP_W_PERM; ep.mapC = e->cnn_C_last; ep.mapL = e->cnn_Lp_last; ep.map_wrows = l.in; ep.map_off = l.perm_off; }
    int rc;
    if (l.out == 2 && g_dtype == 0 && g_ld == 2 && !l.perm_in) {
        // the 2-logit head: weight and bias gradient in one small kernel
        const int rpb = tuning().deterministic ? B : 64;       // deterministic: one block owns the whole sum
        if (dtype_of(e)) head_wgrad_kernel<bf16><<<cdiv(B, rpb), 256, 0, st>>>((const float*)g, (const bf16*)in.p, in.ld, e->grads + l.w, e->grads + l.b, B, l.in, rpb);
        else head_wgrad_kernel<float><<<cdiv(B, rpb), 256, 0, st>>>((const float*)g, (const float*)in.p, in.ld, e->grads + l.w, e->grads + l.b, B, l.in, rpb);
        EMB_CHECK_LAUNCH();
        LAUNCHED(e);
        return EMB_OK;
    }
    if (tc_linear_ok(e, l, B) && g_dtype == 1 && (g_ld % 8) == 0) {
        TcProblem pr = {};
        pr.kind = TC_LINEAR_WGRAD; pr.a = (const bf16*)g; pr.lda = g_ld; pr.b = (const bf16*)in.p; pr.ldb = in.ld;
        pr.M = l.out; pr.N = l.in; pr.K = B;
        if (l.perm_in && l.wg_tmp) {
            // accumulate in the engine's column order (coalesced), then un-permute once into the gradient arena
            Epilogue et = base_epi(e, EPI_ATOMIC, l.wg_tmp, l.in);
            rc = run_tc(e, pr, et, 2.0 * B * l.out * l.in, st);
            if (rc) return rc;
            size_t tot = (size_t)l.out * l.in;
            unpermute_wgrad_kernel<<<cdiv(tot, 256), 256, 0, st>>>(l.wg_tmp, e->grads + l.w, l.out, e->cnn_Lp_last, e->cnn_C_last, l.perm_off);
            EMB_CHECK_LAUNCH();
            LAUNCHED(e);
        } else {
            rc = run_tc(e, pr, ep, 2.0 * B * l.out * l.in, st);
        }
    } else {
        rc = run_gemm(e, A, X, ep, l.out, l.in, B, pick_split_k(l.out, l.in, B), st);
    }
    if (rc) return rc;
    dim3 grid(cdiv(l.out, 32), tuning().deterministic ? 1 : std::min(64, cdiv(B, 8)));
    if (g_dtype) colsum_kernel<bf16><<<grid, dim3(32, 8), 0, st>>>((const bf16*)g, e->grads + l.b, B, l.out, g_ld);
    else colsum_kernel<float><<<grid, dim3(32, 8), 0, st>>>((const float*)g, e->grads + l.b, B, l.out, g_ld);
    EMB_CHECK_LAUNCH();
    LAUNCHED(e);
    return EMB_OK;
}

// gradient w.r.t. the layer input: acc = g W, finished by `ep`
int linear_dgrad(EmbEngine* e, const LinearLayer& l, const void* g, int g_dtype, int g_ld, int B, Epilogue ep, cudaStream_t st) {
    if (l.out == 2 && g_dtype == 0 && g_ld == 2 && !l.perm_in) {
        const size_t tot = (size_t)B * l.in;
        head_dgrad_kernel<<<cdiv(tot, 256), 256, 0, st>>>((const float*)g, e->params + l.w, e->prec == EMB_PREC_BF16 ? 1 : 0, ep, B, l.in);
        EMB_CHECK_LAUNCH();
        LAUNCHED(e);
        return EMB_OK;
    }
    if (tc_linear_ok(e, l, B) && g_dtype == 1 && (g_ld % 8) == 0) {
        TcProblem pr = {};
        pr.kind = TC_LINEAR_DGRAD; pr.a = (const bf16*)g; pr.lda = g_ld; pr.b = l.wc; pr.ldb = round_up(l.in, 8);
        pr.M = B; pr.N = l.in; pr.K = l.out;
        return run_tc(e, pr, ep, 2.0 * B * l.out * l.in, st);
    }
    Operand A = make_operand(g, g_dtype, OP_ROWMAJOR, g_ld, B, l.out);
    Operand W = weight_operand(e, l, true);
    return run_gemm(e, A, W, ep, B, l.in, l.out, 1, st);
}

template <typename T>
int cnn_forward_t(EmbEngine* e, const uint8_t* bases, int B, bool training, const EmbDraws* dr, cudaStream_t st) {
    const int dt = dtype_of(e);
    for (size_t i = 0; i < e->cnn.size(); ++i) {
        ConvLayer& c = e->cnn[i];
        bool stats_done = false;
        const bool even = (c.cout % 2) == 0;
        if (i == 0) {
            const int groups = c.cout / 8;
            if (std::is_same<T, bf16>::value && tc_on(e) && onehot_fwd_tc_ok(c.y, bases, c.cout, c.k, c.ld)) {
                // one-hot rows expanded in shared memory, all taps through a Toeplitz descriptor, fp32 weights as an exact hi/mid/lo bf16 split
                if (!training && tuning().infer_fuse == 2 && tuning().infer_fuse_k1 && onehot_pool_tc_ok(bases, c.cout, c.k, c.ld, c.Lp)) {
                    // inference, transposed form (onehot_pool_tc.cuh): lane = (position block, channel), pooling in registers; y0 is never written
                    bn_finalize_kernel<<<cdiv(c.cout, 128), 128, 0, st>>>(c.stats, e->params + c.gamma, e->params + c.beta, e->buffers + c.rm,
                                                                          e->buffers + c.rv, c.scale, c.shift, c.mean, c.rstd, 1.0, c.cout, 0);
                    EMB_CHECK_LAUNCH();
                    LAUNCHED(e);
                    int rcp = onehot_conv_pool_tc(bases, e->params + c.w, e->params + c.b, c.scale, c.shift, (bf16*)c.a, B, c.cout, c.k, c.Lp, st);
                    if (rcp) return rcp;
                    LAUNCHED(e);
                    continue;
                }
                if (!training && tuning().infer_fuse == 1) {
                    // inference: eval BatchNorm + ReLU + MaxPool on the staged sample inside the same kernel; y0 is never written
                    bn_finalize_kernel<<<cdiv(c.cout, 128), 128, 0, st>>>(c.stats, e->params + c.gamma, e->params + c.beta, e->buffers + c.rm,
                                                                          e->buffers + c.rv, c.scale, c.shift, c.mean, c.rstd, 1.0, c.cout, 0);
                    EMB_CHECK_LAUNCH();
                    LAUNCHED(e);
                    int rcp = onehot_conv_fwd_tc(bases, e->params + c.w, e->params + c.b, (bf16*)c.y, nullptr, B, c.cout, c.k, c.ld, st, c.scale, c.shift,
                                                 (bf16*)c.a, c.Lp);
                    if (rcp) return rcp;
                    LAUNCHED(e);
                    continue;
                }
                int rcf = onehot_conv_fwd_tc(bases, e->params + c.w, e->params + c.b, (bf16*)c.y, training ? c.stats : nullptr, B, c.cout, c.k, c.ld, st);
                if (rcf) return rcf;
                stats_done = true;
            } else if ((c.cout % 8) == 0 && (256 % groups) == 0) {
                // vectorised gather-sum; BatchNorm statistics of layer 0 are accumulated by the same kernel
                const int n_tp = (c.k + 1) / 2, n_tr = (c.k + 2) / 3;
                const size_t smem3 = (size_t)(n_tr * 125 * c.cout + c.cout + 32 * groups * 16) * sizeof(float) + SEQ_LEN + 2 * c.pad + 32;
                if (smem3 <= (size_t)tc_max_smem() && B >= 32 && (1024 % groups) == 0 && !tuning().k1_pairs) {
                    // tap-triple tables (160 KB for 64 channels, k = 15): one persistent 1024-thread CTA per SM
                    onehot_conv_fwd_triple_kernel<T><<<std::min(B, tc_num_sms()), 1024, smem3, st>>>(bases, e->params + c.w, e->params + c.b, (T*)c.y,
                                                                                                   training ? c.stats : nullptr, B, c.cout, c.k, c.ld);
                } else {
                size_t smem = (size_t)(n_tp * 25 * c.cout + c.cout + 256 * 16) * sizeof(float) + SEQ_LEN + 2 * c.pad + 32;
                int grid = std::min(B, 148 * 3);
                onehot_conv_fwd_pair_kernel<T><<<grid, 256, smem, st>>>(bases, e->params + c.w, e->params + c.b, (T*)c.y,
                                                                       training ? c.stats : nullptr, B, c.cout, c.k, c.ld);
                }
                stats_done = true;
            } else {
                size_t smem = (size_t)(c.k * 4 * c.cout + c.cout) * sizeof(float) + SEQ_LEN + 2 * c.pad + 16;
                int grid = std::min(B, 148 * 8);
                onehot_conv_fwd_kernel<T><<<grid, 256, smem, st>>>(bases, e->params + c.w, e->params + c.b, (T*)c.y, B, c.cout, c.k, c.ld, 0);
            }
            EMB_CHECK_LAUNCH();
            LAUNCHED(e);
        } else {
            ConvLayer& pr_ = e->cnn[i - 1];
            Operand A = make_operand(pr_.a, dt, OP_CONV_SHIFT, pr_.ld, B * c.Lc, c.k * c.cin);
            A.L = c.Lc; A.C = c.cin; A.taps = c.k; A.pad = c.pad; A.sign = 1;
            Operand W = make_operand(e->params + c.w, 0, OP_CONV_W_FWD, 0, c.cout, c.k * c.cin);
            W.C = c.cin; W.taps = c.k; W.round_bf16 = e->prec == EMB_PREC_BF16;
            Epilogue ep = base_epi(e, EPI_LINEAR, c.y, c.ld);
            ep.bias = e->params + c.b;
            int rc;
            if (!training && tuning().infer_fuse == 2 && tc_on(e) && std::is_same<T, bf16>::value &&
                tc_conv_pool_ok(c.Lc, c.pad, c.k, c.cin, c.cout, c.Lp, pr_.ld, c.ld)) {
                // Inference, transposed form (conv_pool_tc.cuh): TMEM lane = output channel, columns = positions; eval BatchNorm + ReLU +
                // MaxPool slide along each epilogue thread's own registers.  The pre-pooling conv output is never written.
                bn_finalize_kernel<<<cdiv(c.cout, 128), 128, 0, st>>>(c.stats, e->params + c.gamma, e->params + c.beta, e->buffers + c.rm,
                                                                      e->buffers + c.rv, c.scale, c.shift, c.mean, c.rstd, 1.0, c.cout, 0);
                EMB_CHECK_LAUNCH();
                LAUNCHED(e);
                prof_begin(e, 2.0 * B * c.Lc * c.cout * c.k * c.cin, st);
                int rcp = tc_conv_pool((const bf16*)pr_.a, c.wc, round_up(c.cin, 8), e->params + c.b, c.scale, c.shift, (bf16*)c.a, (bf16*)c.y, B, c.Lc, c.Lp,
                                       c.cin, c.cout, c.k, c.pad, st);
                prof_end(e, st);
                if (rcp) return rcp;
                LAUNCHED(e);
                continue;
            }
            if (!training && tc_conv_ok(e, c) && c.ld == c.cout && tuning().infer_fuse == 1) {
                // Inference: eval-mode BatchNorm (one scale / shift per channel from the running statistics) + ReLU + MaxPool run in
                // the conv GEMM's epilogue on the tile, which holds whole samples: the pre-pooling conv output is never written.
                bn_finalize_kernel<<<cdiv(c.cout, 128), 128, 0, st>>>(c.stats, e->params + c.gamma, e->params + c.beta, e->buffers + c.rm,
                                                                      e->buffers + c.rv, c.scale, c.shift, c.mean, c.rstd, 1.0, c.cout, 0);
                EMB_CHECK_LAUNCH();
                LAUNCHED(e);
                ep.mode = EPI_POOL;
                ep.pool_scale = c.scale; ep.pool_shift = c.shift; ep.pool_out = c.a; ep.pool_Lp = c.Lp; ep.pool_ld = c.ld;
                TcProblem pr = {};
                pr.kind = TC_CONV_FWD; pr.a = (const bf16*)pr_.a; pr.lda = pr_.ld; pr.b = c.wc; pr.ldb = round_up(c.cin, 8);
                pr.M = B * c.Lc; pr.N = c.cout; pr.B = B; pr.L = c.Lc; pr.Cin = c.cin; pr.Cout = c.cout; pr.taps = c.k; pr.pad = c.pad;
                rc = run_tc(e, pr, ep, 2.0 * B * c.Lc * c.cout * c.k * c.cin, st);
                if (rc) return rc;
                continue;
            }
            if (tc_conv_ok(e, c)) {
                TcProblem pr = {};
                pr.kind = TC_CONV_FWD; pr.a = (const bf16*)pr_.a; pr.lda = pr_.ld; pr.b = c.wc; pr.ldb = round_up(c.cin, 8);
                pr.M = B * c.Lc; pr.N = c.cout; pr.B = B; pr.L = c.Lc; pr.Cin = c.cin; pr.Cout = c.cout; pr.taps = c.k; pr.pad = c.pad;
                if (training && c.cout <= 512 && tuning().epi_stats) {
                    ep.bn_stats = c.stats;         // BatchNorm batch statistics accumulated by the GEMM epilogue: no separate pass over y
                    stats_done = true;
                }
                rc = run_tc(e, pr, ep, 2.0 * B * c.Lc * c.cout * c.k * c.cin, st);
            } else {
                rc = run_gemm(e, A, W, ep, B * c.Lc, c.cout, c.k * c.cin, 1, st);
            }
            if (rc) return rc;
        }
        const int64_t R = (int64_t)B * c.Lc;
        if (training) {
            if (!stats_done) {
                if (even) {
                    dim3 grid(cdiv(c.cout / 2, 32), (unsigned)std::min<int64_t>(148 * 8 / std::max(1, cdiv(c.cout / 2, 32)), cdiv(R, 8)));
                    bn_stats_v2_kernel<T><<<grid, dim3(32, 8), 0, st>>>((const T*)c.y, c.stats, R, c.cout, c.ld);
                } else {
                    dim3 grid(cdiv(c.cout, 32), (unsigned)std::min<int64_t>(296, cdiv(R, 8)));
                    bn_stats_kernel<T><<<grid, dim3(32, 8), 0, st>>>((const T*)c.y, c.stats, R, c.cout, c.ld);
                }
                EMB_CHECK_LAUNCH();
                LAUNCHED(e);
            }
            if (e->allreduce && !e->dp_on) {
                int rc = e->allreduce(e->allreduce_user, c.stats, 2 * c.cout, st);
                if (rc) return set_error(EMB_E_STATE, "allreduce callback failed (%d)", rc);
            }
        }
        double n = (double)(e->global_batch > 0 ? e->global_batch : B) * c.Lc;
        if (training && e->dp_on)      // SyncBN: the all-reduce of the partial sums happens inside the finalize kernel (peer memory)
            bn_finalize_dp_kernel<<<1, 256, 0, st>>>(e->dp, (int)i, c.stats, e->params + c.gamma, e->params + c.beta, e->buffers + c.rm,
                                                     e->buffers + c.rv, c.scale, c.shift, c.mean, c.rstd, n, c.cout);
        else
            bn_finalize_kernel<<<cdiv(c.cout, 128), 128, 0, st>>>(c.stats, e->params + c.gamma, e->params + c.beta, e->buffers + c.rm,
                                                                  e->buffers + c.rv, c.scale, c.shift, c.mean, c.rstd, n, c.cout, training ? 1 : 0);
        EMB_CHECK_LAUNCH();
        LAUNCHED(e);
        size_t total = (size_t)B * c.Lp * c.cout;
        float p = training ? c.drop : 0.f;
        const float* du = (dr && p > 0.f) ? dr->cnn_drop[i] : nullptr;
        const bool kt_bwd = std::is_same<T, bf16>::value && kt_ok(c.cout, c.ld) && kt_bwd_smem(c.Lc, c.Lp, c.cout) <= (size_t)kt_max_smem() &&
                            !tuning().no_tma_k2;    // the backward will consume the arg-max codes
        if (training && std::is_same<T, bf16>::value && tuning().k2_wide && (c.cout % 8) == 0 && c.ld == c.cout) {
            // (eval keeps the streaming kernel: without codes and Dropout it already runs at 5.9 TB/s)
            // packed kernel: 8 channels per thread, pooling on the stored bf16 values, BatchNorm on the pooled maximum only
            const int blocks = cdiv((size_t)B * (c.cout / 8), 256);
            uint8_t* am = (training && kt_bwd) ? c.amax : nullptr;
#define EMB_K2_PACKED(MODE)                                                                                                         \
            k2_fwd_packed_kernel<MODE><<<blocks, 256, 0, st>>>((const bf16*)c.y, c.scale, c.shift, (bf16*)c.a, B, c.Lc, c.Lp, c.cout, p, du, \
                                                               e->rng, RNG_CNN_DROP + (uint32_t)i, e->row_offset, am)
            if (p <= 0.f) EMB_K2_PACKED(0);
            else if (du) EMB_K2_PACKED(1);
            else EMB_K2_PACKED(2);
#undef EMB_K2_PACKED
        } else if (even) {
            const int blocks = cdiv((size_t)B * (c.cout / 2), 256);
#define EMB_K2_FWD(MODE)                                                                                                          \
            bn_relu_pool_drop_fwd_stream_kernel<T, MODE><<<blocks, 256, 0, st>>>((const T*)c.y, c.scale, c.shift, (T*)c.a, B, c.Lc, c.Lp, \
                                                                                c.cout, c.ld, p, du, e->rng, RNG_CNN_DROP + (uint32_t)i, e->row_offset, (training && kt_bwd) ? c.amax : nullptr)
            if (p <= 0.f) EMB_K2_FWD(0);
            else if (du) EMB_K2_FWD(1);
            else EMB_K2_FWD(2);
#undef EMB_K2_FWD
        }
        else
            bn_relu_pool_drop_fwd_kernel<T><<<cdiv(total, 256), 256, 0, st>>>((const T*)c.y, c.scale, c.shift, (T*)c.a, B, c.Lc, c.Lp, c.cout, c.ld,
                                                                               p, du, e->rng, RNG_CNN_DROP + (uint32_t)i, e->row_offset);
        EMB_CHECK_LAUNCH();
        LAUNCHED(e);
    }
    return EMB_OK;
}

int cnn_forward(EmbEngine* e, const uint8_t* bases, int B, bool training, const EmbDraws* dr, cudaStream_t st) {
    return e->prec == EMB_PREC_BF16 ? cnn_forward_t<bf16>(e, bases, B, training, dr, st) : cnn_forward_t<float>(e, bases, B, training, dr, st);
}

// backward through the CNN stack; on entry cnn.back().ga holds the gradient w.r.t. the flattened output
template <typename T>
int cnn_backward_t(EmbEngine* e, int B, cudaStream_t st, cudaStream_t sw) {
    const int dt = dtype_of(e);
    for (int i = (int)e->cnn.size() - 1; i >= 0; --i) {
        ConvLayer& c = e->cnn[i];
        const int64_t R = (int64_t)B * c.Lc;
        const bool even = (c.cout % 2) == 0;
        const bool kt = std::is_same<T, bf16>::value && kt_ok(c.cout, c.ld) && kt_bwd_smem(c.Lc, c.Lp, c.cout) <= (size_t)kt_max_smem() &&
                        !tuning().no_tma_k2;
        // Conv bias gradient: sum(dy) is EXACTLY zero behind a training-mode BatchNorm (its backward subtracts the mean of dz and the
        // xhat-weighted mean, which is what makes sum_n dy = 0).  The reference's fp64 autograd leaves ~1e-18 of rounding noise there,
        // far below Adam's eps; summing it in fp32 leaves ~1e-9, which Adam's normalisation would turn into a random walk of the
        // bias that the reference does not have.  The exact value is written instead (the arena was zeroed): nothing to accumulate.
        float* const conv_dbias = nullptr;
        const bool det = tuning().deterministic != 0;
        PoolBwdArgs ka = {};
        if (kt) {
            // pass 1 of 2: the BatchNorm reductions only; dz is recomputed (not stored) by pass 2 below
            ka.y = (const bf16*)c.y; ka.amax = c.amax; ka.ga = (const bf16*)c.ga; ka.scale = c.scale; ka.shift = c.shift;
            ka.mean = c.mean; ka.rstd = c.rstd; ka.gamma = e->params + c.gamma; ka.bstats_in = c.bstats; ka.bstats_out = c.bstats;
            ka.dy = (bf16*)c.dy; ka.dbias = conv_dbias; ka.B = B; ka.Lc = c.Lc; ka.Lp = c.Lp; ka.C = c.cout; ka.drop_p = c.drop;
            ka.n = (double)(e->global_batch > 0 ? e->global_batch : B) * c.Lc;
            kt_segments(c.cout, c.Lc, 2, &ka.nseg, &ka.P);
            const size_t ksm = kt_bwd_smem(c.Lc, c.Lp, c.cout);
            pool_bn_bwd_tma_kernel<0><<<std::min(B, tc_num_sms() * kt_ctas_per_sm(ksm, 3)), KT_THREADS, ksm, st>>>(ka);
            EMB_CHECK_LAUNCH();
            LAUNCHED(e);
        } else if (even) {
            dim3 grid(cdiv(c.cout / 2, 32), cdiv(B, 8));
            pool_bn_bwd_stream_kernel<T><<<grid, dim3(32, 8), 0, st>>>((const T*)c.y, (const T*)c.a, (const T*)c.ga, c.scale, c.shift, c.mean,
                                                                       c.rstd, (T*)c.dy, c.bstats, B, c.Lc, c.Lp, c.cout, c.ld, c.drop);
            EMB_CHECK_LAUNCH();
            LAUNCHED(e);
        } else {
            dim3 grid(cdiv(c.cout, 32), B);
            size_t smem = (size_t)2 * c.Lc * 32 * sizeof(float);
            pool_bn_bwd_stage1_kernel<T><<<grid, dim3(32, 8), smem, st>>>((const T*)c.y, (const T*)c.a, (const T*)c.ga, c.scale, c.shift, c.mean,
                                                                          c.rstd, (T*)c.dy, c.bstats, B, c.Lc, c.Lp, c.cout, c.ld, c.drop);
            EMB_CHECK_LAUNCH();
            LAUNCHED(e);
        }
        if (e->dp_on) {
            bn_bwd_finalize_dp_kernel<<<1, 256, 0, st>>>(e->dp, EMB_MAX_CNN + i, c.bstats, e->grads + c.gamma, e->grads + c.beta, c.cout);
            EMB_CHECK_LAUNCH();
            LAUNCHED(e);
        } else if (e->allreduce) {
            int rc = e->allreduce(e->allreduce_user, c.bstats, 2 * c.cout, st);
            if (rc) return set_error(EMB_E_STATE, "allreduce callback failed (%d)", rc);
        }
        // under data parallelism the all-reduced statistics are already global: only the first shard contributes
        // them to the gradient arena (which is then sum-reduced across ranks)
        if (!e->dp_on && (!e->allreduce || e->row_offset == 0)) {
            bn_bwd_finalize_kernel<<<cdiv(c.cout, 128), 128, 0, st>>>(c.bstats, e->grads + c.gamma, e->grads + c.beta, c.cout);
            EMB_CHECK_LAUNCH();
            LAUNCHED(e);
        }
        double n = (double)(e->global_batch > 0 ? e->global_batch : B) * c.Lc;
        if (kt) {
            const size_t ksm = kt_bwd_smem(c.Lc, c.Lp, c.cout);
            pool_bn_bwd_tma_kernel<1><<<std::min(B, tc_num_sms() * kt_ctas_per_sm(ksm, 3)), KT_THREADS, ksm, st>>>(ka);
        } else if ((c.cout % 8) == 0 && c.cout <= 2048) {
            const int G = c.cout / 8, rpb = std::max(1, 256 / G);
            dim3 block(G, rpb);
            const int grid = (int)std::min<int64_t>(148 * 8, cdiv(R, rpb));
            bn_bwd_apply_v8_kernel<T><<<grid, block, (size_t)rpb * c.cout * sizeof(float), st>>>((const T*)c.y, (T*)c.dy, c.bstats, e->params + c.gamma,
                                                                                                c.mean, c.rstd, conv_dbias, R, c.cout, c.ld, n);
        } else if (even) {
            const int gx = cdiv(c.cout / 2, 32);
            dim3 grid(gx, (unsigned)std::min<int64_t>(std::max(1, 148 * 8 / gx), cdiv(R, 8)));
            bn_bwd_apply_v2_kernel<T><<<grid, dim3(32, 8), 0, st>>>((const T*)c.y, (T*)c.dy, c.bstats, e->params + c.gamma, c.mean, c.rstd,
                                                                    conv_dbias, R, c.cout, c.ld, n);
        } else {
            dim3 grid(cdiv(c.cout, 32), (unsigned)std::min<int64_t>(296, cdiv(R, 8)));
            bn_bwd_apply_kernel<T><<<grid, dim3(32, 8), 0, st>>>((const T*)c.y, (T*)c.dy, c.bstats, e->params + c.gamma, c.mean, c.rstd,
                                                                 conv_dbias, R, c.cout, c.ld, n);
        }
        EMB_CHECK_LAUNCH();
        LAUNCHED(e);
        if (i == 0) {
            // conv-0 dbias was already accumulated by the bn_bwd_apply kernel
            if (std::is_same<T, bf16>::value && tc_on(e) && onehot_wgrad_tc_ok(c.dy, e->last_bases, c.cout, c.k, c.ld)) {
                // the weight gradient of the one-hot layer as ONE tensor-core contraction per 16 positions (onehot_wgrad_tc.cuh)
                int rcw = onehot_conv_wgrad_tc(e->last_bases, (const bf16*)c.dy, e->grads + c.w, B, c.cout, c.k, c.ld, st, det ? 1 : 0);
                if (rcw) return rcw;
            } else if (even && (c.cout / 2) * c.k <= 1024) {
                const int threads = round_up((c.cout / 2) * c.k, 32);
                size_t smem = (size_t)(SEQ_LEN + 2 * c.pad) * c.cout * sizeof(float) + SEQ_LEN + 16;
                int grid = det ? 1 : std::min(B, 148 * 2);
                if (std::is_same<T, bf16>::value && (c.cout % 8) == 0) {
                    const size_t smem16 = (size_t)(SEQ_LEN + 2 * c.pad) * c.cout * sizeof(bf16) + SEQ_LEN + 16;
                    onehot_conv_bwd_lists16_kernel<<<grid, threads, smem16, st>>>(e->last_bases, (const bf16*)c.dy, e->grads + c.w, B, c.cout, c.k, c.ld);
                } else
                onehot_conv_bwd_lists_kernel<T><<<grid, threads, smem, st>>>(e->last_bases, (const T*)c.dy, e->grads + c.w, B, c.cout, c.k, c.ld);
            } else {
                int grid = det ? 1 : std::min(B, 148 * 4);
                onehot_conv_bwd_kernel<T><<<grid, 256, 0, st>>>(e->last_bases, (const T*)c.dy, e->grads + c.w, nullptr, B, c.cout, c.k, c.ld);
            }
            EMB_CHECK_LAUNCH();
            LAUNCHED(e);
        } else {
            ConvLayer& pr = e->cnn[i - 1];
            const bool tc = tc_conv_ok(e, c);
            const double flops = 2.0 * R * c.cout * c.k * c.cin;
            // wgrad: dW[o][c][tap] = sum_{b,l} dy[b,l,o] * a_prev[b, l+tap-pad, c]
            Operand A = make_operand(c.dy, dt, OP_TRANSPOSED, c.ld, c.cout, (int)R);
            Operand X = make_operand(pr.a, dt, OP_CONV_SHIFT_T, pr.ld, c.k * c.cin, (int)R);
            X.L = c.Lc; X.C = c.cin; X.taps = c.k; X.pad = c.pad; X.sign = 1;
            Epilogue ep = base_epi(e, EPI_ATOMIC, e->grads + c.w, 0);
            ep.map = MAP_CONV_W; ep.mapC = c.cin; ep.map_taps = c.k;
            int rc;
            if ((rc = stream_after(e, st, sw))) return rc;        // the weight gradient runs next to the data gradient
            if (tc) {
                TcProblem tp = {};
                tp.kind = TC_CONV_WGRAD; tp.a = (const bf16*)c.dy; tp.lda = c.ld; tp.b = (const bf16*)pr.a; tp.ldb = pr.ld;
                tp.M = c.cout; tp.N = c.cin; tp.B = B; tp.L = c.Lc; tp.Cin = c.cin; tp.Cout = c.cout; tp.taps = c.k; tp.pad = c.pad;
                // accumulate as [tap][Cout][Cin] (coalesced atomics), then one permuted write-back into the gradient arena
                const size_t wn = (size_t)c.k * c.cout * c.cin;
                tp.wgrad_tap_stride = c.cout * c.cin;
                Epilogue et = base_epi(e, EPI_ATOMIC, c.wg_tmp, c.cin);
                rc = run_tc(e, tp, et, flops, sw);
                if (rc) return rc;
                unpermute_conv_wgrad_kernel<<<cdiv(wn, 256), 256, 0, sw>>>(c.wg_tmp, e->grads + c.w, c.cout, c.cin, c.k);
                EMB_CHECK_LAUNCH();
                LAUNCHED(e);
            } else {
                rc = run_gemm(e, A, X, ep, c.cout, c.k * c.cin, (int)R, pick_split_k(c.cout, c.k * c.cin, (int)R), sw);
            }
            if (rc) return rc;
            // dgrad: ga_prev[b,l',c] = sum_{tap,o} dy[b, l'-tap+pad, o] * W[o][c][tap]
            Operand G = make_operand(c.dy, dt, OP_CONV_SHIFT, c.ld, (int)R, c.k * c.cout);
            G.L = c.Lc; G.C = c.cout; G.taps = c.k; G.pad = c.pad; G.sign = -1;
            Operand W = make_operand(e->params + c.w, 0, OP_CONV_W_DGRAD, 0, c.cin, c.k * c.cout);
            W.C = c.cout; W.taps = c.k; W.wrows = c.cin; W.round_bf16 = e->prec == EMB_PREC_BF16;
            Epilogue ed = base_epi(e, EPI_LINEAR, pr.ga, pr.ld);
            if (tc) {
                TcProblem tp = {};
                tp.kind = TC_CONV_DGRAD; tp.a = (const bf16*)c.dy; tp.lda = c.ld; tp.b = c.wc; tp.ldb = round_up(c.cin, 8);
                tp.M = (int)R; tp.N = c.cin; tp.B = B; tp.L = c.Lc; tp.Cin = c.cin; tp.Cout = c.cout; tp.taps = c.k; tp.pad = c.pad;
                rc = run_tc(e, tp, ed, flops, st);
            } else {
                rc = run_gemm(e, G, W, ed, (int)R, c.cin, c.k * c.cout, 1, st);
            }
            if (rc) return rc;
        }
    }
    return EMB_OK;
}

int cnn_backward(EmbEngine* e, int B, cudaStream_t st, cudaStream_t sw) {
    return e->prec == EMB_PREC_BF16 ? cnn_backward_t<bf16>(e, B, st, sw) : cnn_backward_t<float>(e, B, st, sw);
}

int check_ready(EmbEngine* e, int B, bool need_grads) {
    if (!e) return set_error(EMB_E_ARG, "null engine");
    if (!e->params || !e->ws) return set_error(EMB_E_STATE, "emb_bind() has not been called");
    int cur = -1;
    if (e->device >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != e->device) EMB_CUDA_OK(cudaSetDevice(e->device));
    if (need_grads && (!e->grads)) return set_error(EMB_E_STATE, "no gradient arena bound");
    if (B < 1 || B > e->max_batch) return set_error(EMB_E_ARG, "batch %d outside [1, %d]", B, e->max_batch);
    return EMB_OK;
}

int forward_impl(EmbEngine* e, const float* x_ffnn, const uint8_t* bases, const float* avail, int B, bool training,
                 const EmbDraws* dr, float* logits_out, cudaStream_t st) {
    const EmbArchSpec& s = e->spec;
    const int dt = dtype_of(e);
    int rc;
    e->last_B = B;
    e->last_training = training;
    e->last_bases = bases;
    if (training && e->dp_on) {
        dp_epoch_advance_kernel<<<1, 1, 0, st>>>(e->dp_epoch);        // one epoch per training step: the exchange kernels' flag value
        EMB_CHECK_LAUNCH();
        LAUNCHED(e);
    }
    if ((rc = refresh_wcache(e, st))) return rc;
    if (training && e->zero_fwd_bytes) EMB_CUDA_OK(cudaMemsetAsync(e->zero_fwd, 0, e->zero_fwd_bytes, st));
    const Act* ffnn_last = nullptr;
    // the FFNN chain (and docking_0) is independent of the CNN chain until the embracement / concatenation: side stream
    const bool two_chains = s.kind == EMB_KIND_EMBRACENET || s.kind == EMB_KIND_CONCATNET;
    cudaStream_t sf = (two_chains && forking(e)) ? e->side[0] : st;
    if (s.kind != EMB_KIND_CNN) {
        if (!x_ffnn) return set_error(EMB_E_ARG, "x_ffnn is NULL");
        if ((rc = stream_after(e, st, sf))) return rc;
        size_t tot = (size_t)B * e->x0.ld;
        if (dt) cast_rows_kernel<bf16><<<cdiv(tot, 256), 256, 0, sf>>>(x_ffnn, (bf16*)e->x0.p, B, s.in_features, e->x0.ld);
        else cast_rows_kernel<float><<<cdiv(tot, 256), 256, 0, sf>>>(x_ffnn, (float*)e->x0.p, B, s.in_features, e->x0.ld);
        EMB_CHECK_LAUNCH();
        LAUNCHED(e);
        const Act* in = &e->x0;
        for (size_t i = 0; i < e->ffnn.size(); ++i) {
            const float* du = (dr && training && e->ffnn[i].drop > 0) ? dr->ffnn_drop[i] : nullptr;
            rc = linear_forward(e, e->ffnn[i], *in, e->ffnn_h[i], B, training, du, RNG_FFNN_DROP + (uint32_t)i, sf);
            if (rc) return rc;
            in = &e->ffnn_h[i];
        }
        ffnn_last = in;
        if (s.kind == EMB_KIND_EMBRACENET) {
            rc = linear_forward(e, e->dock0, *ffnn_last, e->d0, B, false, nullptr, 0, sf);
            if (rc) return rc;
        }
    }
    Act flat;
    if (s.kind != EMB_KIND_FFNN) {
        if (!bases) return set_error(EMB_E_ARG, "bases is NULL");
        if (training && bases != e->in_bases) {   // backward re-reads the bases: keep a private copy (256 B/row)
            EMB_CUDA_OK(cudaMemcpyAsync(e->in_bases, bases, (size_t)B * SEQ_LEN, cudaMemcpyDeviceToDevice, st));
            bases = e->in_bases;
            e->last_bases = bases;
        }
        rc = cnn_forward(e, bases, B, training, dr, st);
        if (rc) return rc;
        flat.p = e->cnn.back().a;
        flat.width = e->cnn_out;
        flat.ld = e->cnn_Lp_last * e->cnn_ld_last;
    }
    float* logits = logits_out ? logits_out : e->logits;
    const Act* head_in = nullptr;
    if (s.kind == EMB_KIND_EMBRACENET) {
        const int C = s.embracement_size;
        const bool mdrop = training && s.embracenet_dropout;
        embrace_prologue_kernel<<<cdiv(B, 128), 128, 0, st>>>(s.p_ffnn, avail, mdrop ? 1 : 0, dr ? dr->has_modal_u0 : 0, dr ? dr->modal_u0 : 0.f,
                                                              dr ? dr->modal_rows : nullptr, e->rng, e->row_offset, e->cum0, B);
        EMB_CHECK_LAUNCH();
        LAUNCHED(e);
        if ((rc = stream_after(e, sf, st))) return rc;      // join: docking_1's epilogue reads docking_0's output
        {   // docking_1 GEMM with the embracement select fused into its epilogue
            Operand A = input_operand(e, e->dock1, flat, B, false);
            Operand W = weight_operand(e, e->dock1, false);
            Epilogue ep = base_epi(e, EPI_EMBRACE, e->e.p, e->e.ld);
            ep.bias = e->params + e->dock1.b;
            ep.d0 = e->d0.p; ep.ld_d0 = e->d0.ld;
            ep.emb_u = dr ? dr->embrace_u : nullptr;
            ep.cum0 = e->cum0;
            ep.idx_out = e->idx;
            if (tc_linear_ok(e, e->dock1, B)) {
                TcProblem pr = {};
                pr.kind = TC_LINEAR_FWD; pr.a = (const bf16*)flat.p; pr.lda = flat.ld; pr.b = e->dock1.wc; pr.ldb = round_up(e->dock1.in, 8);
                pr.M = B; pr.N = C; pr.K = e->dock1.in;
                rc = run_tc(e, pr, ep, 2.0 * B * C * e->dock1.in, st);
            } else {
                rc = run_gemm(e, A, W, ep, B, C, e->dock1.in, 1, st);
            }
            if (rc) return rc;
        }
        const Act* in = &e->e;
        for (size_t i = 0; i < e->post.size(); ++i) {
            const float* du = (dr && training && e->post[i].drop > 0) ? dr->post_drop[i] : nullptr;
            rc = linear_forward(e, e->post[i], *in, e->post_h[i], B, training, du, RNG_POST_DROP + (uint32_t)i, st);
            if (rc) return rc;
            in = &e->post_h[i];
        }
        head_in = in;
    } else if (s.kind == EMB_KIND_CONCATNET) {
        if ((rc = stream_after(e, sf, st))) return rc;      // join: the concatenation reads both chains
        const size_t tot = (size_t)B * (e->ffnn_out + e->cnn_out);
        if (dt) concat_kernel<bf16><<<cdiv(tot, 256), 256, 0, st>>>((const bf16*)ffnn_last->p, ffnn_last->ld, e->ffnn_out, (const bf16*)flat.p, e->cnn_Lp_last,
                                                                 e->cnn_C_last, e->cnn_ld_last, (bf16*)e->xcat.p, e->xcat.ld, B);
        else concat_kernel<float><<<cdiv(tot, 256), 256, 0, st>>>((const float*)ffnn_last->p, ffnn_last->ld, e->ffnn_out, (const float*)flat.p, e->cnn_Lp_last,
                                                                 e->cnn_C_last, e->cnn_ld_last, (float*)e->xcat.p, e->xcat.ld, B);
        EMB_CHECK_LAUNCH();
        LAUNCHED(e);
        const Act* in = &e->xcat;
        for (size_t i = 0; i < e->post.size(); ++i) {
            const float* du = (dr && training && e->post[i].drop > 0) ? dr->post_drop[i] : nullptr;
            rc = linear_forward(e, e->post[i], *in, e->post_h[i], B, training, du, RNG_POST_DROP + (uint32_t)i, st);
            if (rc) return rc;
            in = &e->post_h[i];
        }
        head_in = in;
    } else if (s.kind == EMB_KIND_FFNN) {
        head_in = ffnn_last;
    } else {
        const Act* in = &flat;
        for (size_t i = 0; i + 1 < e->head.size(); ++i) {
            rc = linear_forward(e, e->head[i], *in, e->head_h[i], B, false, nullptr, 0, st);
            if (rc) return rc;
            in = &e->head_h[i];
        }
        head_in = in;
    }
    {   // final Linear -> fp32 logits
        const LinearLayer& l = e->head.back();
        Operand A = input_operand(e, l, *head_in, B, false);
        Operand W = weight_operand(e, l, false);
        Epilogue ep = base_epi(e, EPI_LINEAR, logits, 2);
        ep.out_dtype = 0;
        ep.bias = e->params + l.b;
        if (l.out == 2 && !l.perm_in && l.in <= 4096) {
            const int grid = std::min(cdiv(B, 8), 148 * 8);
            const size_t smem = (size_t)2 * l.in * sizeof(float);
            if (dt) head_fwd_kernel<bf16><<<grid, 256, smem, st>>>((const bf16*)head_in->p, head_in->ld, e->params + l.w, e->params + l.b, 1, logits, B, l.in);
            else head_fwd_kernel<float><<<grid, 256, smem, st>>>((const float*)head_in->p, head_in->ld, e->params + l.w, e->params + l.b, 0, logits, B, l.in);
            EMB_CHECK_LAUNCH();
            LAUNCHED(e);
        } else {
            rc = run_gemm(e, A, W, ep, B, 2, l.in, 1, st);
            if (rc) return rc;
        }
    }
    rng_advance_kernel<<<1, 1, 0, st>>>(e->rng);   // next forward (train or eval: multinomial is sampled in both) draws afresh
    EMB_CHECK_LAUNCH();
    LAUNCHED(e);
    return EMB_OK;
}

int backward_impl(EmbEngine* e, const float* dlogits, cudaStream_t st) {
    const EmbArchSpec& s = e->spec;
    const int B = e->last_B;
    const int dt = dtype_of(e);
    if (!e->last_training || B < 1) return set_error(EMB_E_STATE, "emb_backward needs a preceding emb_forward_train");
    int rc;
    EMB_CUDA_OK(cudaMemsetAsync(e->grads, 0, e->n_params * sizeof(float), st));
    if (e->zero_bwd_bytes) EMB_CUDA_OK(cudaMemsetAsync(e->zero_bwd, 0, e->zero_bwd_bytes, st));
    const void* g = dlogits;     // gradient w.r.t. the current layer's pre-activation
    int g_dt = 0, g_ld = 2;
    // Branches: the data-gradient chain stays on `st`; every weight gradient goes to `sw` as soon as its input gradient exists;
    // the FFNN backward chain runs on `sf` next to the docking_1 / CNN backward.  All three are joined before returning.
    const bool fk = forking(e);
    cudaStream_t sw = fk ? e->side[1] : st, sf = fk ? e->side[0] : st;

    auto mask_epi = [&](const Act& ref, float drop, const Act& out) {
        Epilogue ep = base_epi(e, EPI_MASKGRAD, out.p, out.ld);
        ep.ref = ref.p; ep.ld_ref = ref.ld;
        ep.scale = drop > 0.f ? 1.f / (1.f - drop) : 1.f;
        return ep;
    };
    // weight gradient of `l` on the weight-gradient stream, ordered after everything `from` has produced so far
    auto wgrad_on = [&](cudaStream_t from, const LinearLayer& l, const void* gg, int gdt, int gld, const Act& in) -> int {
        int r = stream_after(e, from, sw);
        if (r) return r;
        return linear_wgrad(e, l, gg, gdt, gld, in, B, sw);
    };
    // the FFNN stack's backward, entirely on `sf` (its GEMMs are tiny: the chain is what costs)
    auto ffnn_chain = [&]() -> int {
        int r = stream_after(e, st, sf);
        if (r) return r;
        for (int i = (int)e->ffnn.size() - 1; i >= 0; --i) {
            const Act& in = i ? e->ffnn_h[i - 1] : e->x0;
            if ((r = wgrad_on(sf, e->ffnn[i], e->ffnn_g[i].p, dt, e->ffnn_g[i].ld, in))) return r;
            if (i) {
                r = linear_dgrad(e, e->ffnn[i], e->ffnn_g[i].p, dt, e->ffnn_g[i].ld, B, mask_epi(e->ffnn_h[i - 1], e->ffnn[i - 1].drop, e->ffnn_g[i - 1]), sf);
                if (r) return r;
            }
        }
        return EMB_OK;
    };

    if (s.kind == EMB_KIND_EMBRACENET) {
        const LinearLayer& hl = e->head.back();
        const int np = (int)e->post.size();
        const Act& head_in = np ? e->post_h[np - 1] : e->e;
        rc = wgrad_on(st, hl, g, g_dt, g_ld, head_in);
        if (rc) return rc;
        auto embrace_bwd_epi = [&]() {
            Epilogue ep = base_epi(e, EPI_EMBRACE_BWD, e->dd0.p, e->dd0.ld);
            ep.out2 = e->dd1.p;
            ep.idx = e->idx;
            ep.e = e->e.p; ep.ld_e = e->e.ld;
            return ep;
        };
        if (np == 0) rc = linear_dgrad(e, hl, g, g_dt, g_ld, B, embrace_bwd_epi(), st);
        else rc = linear_dgrad(e, hl, g, g_dt, g_ld, B, mask_epi(e->post_h[np - 1], e->post[np - 1].drop, e->post_g[np - 1]), st);
        if (rc) return rc;
        for (int i = np - 1; i >= 0; --i) {
            const Act& in = i ? e->post_h[i - 1] : e->e;
            rc = wgrad_on(st, e->post[i], e->post_g[i].p, dt, e->post_g[i].ld, in);
            if (rc) return rc;
            if (i == 0) rc = linear_dgrad(e, e->post[i], e->post_g[i].p, dt, e->post_g[i].ld, B, embrace_bwd_epi(), st);
            else rc = linear_dgrad(e, e->post[i], e->post_g[i].p, dt, e->post_g[i].ld, B, mask_epi(e->post_h[i - 1], e->post[i - 1].drop, e->post_g[i - 1]), st);
            if (rc) return rc;
        }
        // docking layers
        const int nf = (int)e->ffnn.size();
        rc = wgrad_on(st, e->dock0, e->dd0.p, dt, e->dd0.ld, e->ffnn_h[nf - 1]);
        if (rc) return rc;
        rc = linear_dgrad(e, e->dock0, e->dd0.p, dt, e->dd0.ld, B, mask_epi(e->ffnn_h[nf - 1], e->ffnn[nf - 1].drop, e->ffnn_g[nf - 1]), st);
        if (rc) return rc;
        if (fk && (rc = ffnn_chain())) return rc;                 // forked here; otherwise it runs below, in the round-1 order
        Act flat;
        flat.p = e->cnn.back().a; flat.width = e->cnn_out; flat.ld = e->cnn_Lp_last * e->cnn_ld_last;
        rc = wgrad_on(st, e->dock1, e->dd1.p, dt, e->dd1.ld, flat);
        if (rc) return rc;
        {
            Epilogue ep = base_epi(e, EPI_LINEAR, e->cnn.back().ga, e->cnn_Lp_last * e->cnn_ld_last);
            if (e->cnn_ld_last != e->cnn_C_last) { ep.flatC = e->cnn_C_last; ep.flat_ldc = e->cnn_ld_last; }   // dense rows need no remap
            rc = linear_dgrad(e, e->dock1, e->dd1.p, dt, e->dd1.ld, B, ep, st);
            if (rc) return rc;
        }
    } else if (s.kind == EMB_KIND_CONCATNET) {
        const LinearLayer& hl = e->head.back();
        const int np = (int)e->post.size(), nf = (int)e->ffnn.size();
        rc = wgrad_on(st, hl, g, g_dt, g_ld, e->post_h[np - 1]);
        if (rc) return rc;
        rc = linear_dgrad(e, hl, g, g_dt, g_ld, B, mask_epi(e->post_h[np - 1], e->post[np - 1].drop, e->post_g[np - 1]), st);
        if (rc) return rc;
        for (int i = np - 1; i >= 0; --i) {
            const Act& in = i ? e->post_h[i - 1] : e->xcat;
            rc = wgrad_on(st, e->post[i], e->post_g[i].p, dt, e->post_g[i].ld, in);
            if (rc) return rc;
            if (i) rc = linear_dgrad(e, e->post[i], e->post_g[i].p, dt, e->post_g[i].ld, B, mask_epi(e->post_h[i - 1], e->post[i - 1].drop, e->post_g[i - 1]), st);
            else rc = linear_dgrad(e, e->post[i], e->post_g[i].p, dt, e->post_g[i].ld, B, base_epi(e, EPI_LINEAR, e->gcat.p, e->gcat.ld), st);
            if (rc) return rc;
        }
        const size_t tot = (size_t)B * (e->ffnn_out + e->cnn_out);
        const float scale = e->ffnn[nf - 1].drop > 0.f ? 1.f / (1.f - e->ffnn[nf - 1].drop) : 1.f;
        if (dt) concat_split_kernel<bf16><<<cdiv(tot, 256), 256, 0, st>>>((const bf16*)e->gcat.p, e->gcat.ld, e->ffnn_out, (const bf16*)e->ffnn_h[nf - 1].p,
                e->ffnn_h[nf - 1].ld, scale, (bf16*)e->ffnn_g[nf - 1].p, e->ffnn_g[nf - 1].ld, (bf16*)e->cnn.back().ga, e->cnn_Lp_last, e->cnn_C_last, e->cnn_ld_last, B);
        else concat_split_kernel<float><<<cdiv(tot, 256), 256, 0, st>>>((const float*)e->gcat.p, e->gcat.ld, e->ffnn_out, (const float*)e->ffnn_h[nf - 1].p,
                e->ffnn_h[nf - 1].ld, scale, (float*)e->ffnn_g[nf - 1].p, e->ffnn_g[nf - 1].ld, (float*)e->cnn.back().ga, e->cnn_Lp_last, e->cnn_C_last, e->cnn_ld_last, B);
        EMB_CHECK_LAUNCH();
        LAUNCHED(e);
        if (fk && (rc = ffnn_chain())) return rc;
    } else if (s.kind == EMB_KIND_FFNN) {
        const LinearLayer& hl = e->head.back();
        const int nf = (int)e->ffnn.size();
        rc = wgrad_on(st, hl, g, g_dt, g_ld, e->ffnn_h[nf - 1]);
        if (rc) return rc;
        rc = linear_dgrad(e, hl, g, g_dt, g_ld, B, mask_epi(e->ffnn_h[nf - 1], e->ffnn[nf - 1].drop, e->ffnn_g[nf - 1]), st);
        if (rc) return rc;
    } else {   // CNN: three plain Linear layers
        Act flat;
        flat.p = e->cnn.back().a; flat.width = e->cnn_out; flat.ld = e->cnn_Lp_last * e->cnn_ld_last;
        for (int i = (int)e->head.size() - 1; i >= 0; --i) {
            const Act& in = i ? e->head_h[i - 1] : flat;
            rc = wgrad_on(st, e->head[i], g, g_dt, g_ld, in);
            if (rc) return rc;
            Epilogue ep;
            if (i) ep = base_epi(e, EPI_LINEAR, e->head_g[i - 1].p, e->head_g[i - 1].ld);
            else {
                ep = base_epi(e, EPI_LINEAR, e->cnn.back().ga, e->cnn_Lp_last * e->cnn_ld_last);
                if (e->cnn_ld_last != e->cnn_C_last) { ep.flatC = e->cnn_C_last; ep.flat_ldc = e->cnn_ld_last; }   // dense rows need no remap
            }
            rc = linear_dgrad(e, e->head[i], g, g_dt, g_ld, B, ep, st);
            if (rc) return rc;
            if (i) { g = e->head_g[i - 1].p; g_dt = dt; g_ld = e->head_g[i - 1].ld; }
        }
    }
    if (s.kind != EMB_KIND_CNN && !(fk && (s.kind == EMB_KIND_EMBRACENET || s.kind == EMB_KIND_CONCATNET))) {
        // the FFNN stack on the main chain (single-modality FFNN, or forking off): sf == st unless kind FFNN forks its wgrads
        cudaStream_t keep = sf;
        sf = st;
        rc = ffnn_chain();
        sf = keep;
        if (rc) return rc;
    }
    if (e->phase_hook) {
        // every gradient outside the CNN stack is final here: a data-parallel host can start reducing those arena slices
        // while the CNN backward (the bulk of the step) still runs
        rc = e->phase_hook(e->phase_user, 1, (void*)st);
        if (rc) return set_error(EMB_E_STATE, "phase hook failed (%d)", rc);
    }
    if (s.kind != EMB_KIND_FFNN) {
        rc = cnn_backward(e, B, st, sw);
        if (rc) return rc;
    }
    if (fk) {       // join the branches: the caller (optimizer, gradient reduction, the host) sees complete gradients on `st`
        if ((rc = stream_after(e, sw, st))) return rc;
        if ((rc = stream_after(e, sf, st))) return rc;
    }
    return EMB_OK;
}

void fill_opt_scalars(EmbEngine* e, const EmbOptConfig& c, OptScalars* o) {
    e->opt_t += 1;
    const double t = (double)e->opt_t;
    o->kind = c.kind;
    o->lr = c.lr; o->wd = c.weight_decay; o->b1 = c.beta1; o->b2 = c.beta2; o->eps = c.eps; o->alpha = c.alpha;
    o->bc1 = (float)(1.0 - std::pow((double)c.beta1, t));
    o->bc2 = (float)(1.0 - std::pow((double)c.beta2, t));
    o->bc2_sqrt = (float)std::sqrt(1.0 - std::pow((double)c.beta2, t));
    o->nadam_c_g = o->nadam_c_m = 0.f;
    if (c.kind == EMB_OPT_NADAM) {
        double psi = c.momentum_decay;
        double mu = c.beta1 * (1.0 - 0.5 * std::pow(0.96, t * psi));
        double mu_next = c.beta1 * (1.0 - 0.5 * std::pow(0.96, (t + 1) * psi));
        e->nadam_mu_product *= mu;
        double mp = e->nadam_mu_product;
        o->nadam_c_g = (float)(c.lr * (1.0 - mu) / (1.0 - mp));
        o->nadam_c_m = (float)(c.lr * mu_next / (1.0 - mp * mu_next));
    }
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

const char* emb_last_error(void) { return g_last_error.c_str(); }
int emb_abi_version(void) { return EMB_ABI_VERSION; }

int emb_set_option(const char* name, int32_t value) {
    if (!name) return set_error(EMB_E_ARG, "null option name");
    int n = 0;
    const TuningName* names = tuning_names(&n);
    for (int i = 0; i < n; ++i)
        if (!strcmp(name, names[i].name) || !strcmp(name, names[i].env)) { tuning().*(names[i].field) = value; return EMB_OK; }
    return set_error(EMB_E_ARG, "unknown option '%s'", name);
}

int emb_get_option(const char* name, int32_t* value_out) {
    if (!name || !value_out) return set_error(EMB_E_ARG, "null argument");
    int n = 0;
    const TuningName* names = tuning_names(&n);
    for (int i = 0; i < n; ++i)
        if (!strcmp(name, names[i].name) || !strcmp(name, names[i].env)) { *value_out = tuning().*(names[i].field); return EMB_OK; }
    return set_error(EMB_E_ARG, "unknown option '%s'", name);
}

int emb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
    }
    return ok;
}

int emb_create(const EmbArchSpec* spec, int32_t max_batch, int32_t precision, EmbEngine** out) {
    if (!spec || !out) return set_error(EMB_E_ARG, "null argument");
    if (max_batch < 1) return set_error(EMB_E_ARG, "max_batch must be >= 1");
    if (precision != EMB_PREC_FP32 && precision != EMB_PREC_BF16) return set_error(EMB_E_ARG, "unknown precision %d", precision);
    if (spec->kind < 0 || spec->kind > 3) return set_error(EMB_E_ARG, "unknown kind %d", spec->kind);
    EmbEngine* e = new EmbEngine();
    e->spec = *spec;
    e->max_batch = max_batch;
    e->prec = precision;
    e->esize = precision == EMB_PREC_BF16 ? 2 : 4;
    int rc = plan(e);
    if (rc) { delete e; return rc; }
    e->ws_bytes = carve(e, nullptr);
    *out = e;
    return EMB_OK;
}

void emb_destroy(EmbEngine* e) {
    if (!e) return;
    for (auto ev : e->prof_ev) cudaEventDestroy(ev);
    for (auto& g : e->graphs) cudaGraphExecDestroy(g.exec);
    for (auto ev : e->fork_ev) cudaEventDestroy(ev);
    for (int k = 0; k < 2; ++k) if (e->side[k]) cudaStreamDestroy(e->side[k]);
    if (e->gpos_dev) cudaFree(e->gpos_dev);
    if (e->gstream) { cudaStreamDestroy(e->gstream); cudaEventDestroy(e->gev_in); cudaEventDestroy(e->gev_out); }
    if (e->infer_copy) {
        cudaStreamDestroy(e->infer_copy);
        for (auto ev : e->infer_ev) cudaEventDestroy(ev);
    }
    if (e->copy_stream) {
        cudaStreamDestroy(e->copy_stream);
        for (int k = 0; k < 2; ++k) { cudaEventDestroy(e->ev_h2d[k]); cudaEventDestroy(e->ev_consumed[k]); cudaEventDestroy(e->ev_step[k]); }
        cudaEventDestroy(e->ev_d2h);
        cudaFreeHost(e->pin_metrics);
    }
    if (e->owns_memory) {
        cudaFree(e->params); cudaFree(e->grads); cudaFree(e->buffers); cudaFree(e->opt_m); cudaFree(e->opt_v); cudaFree(e->ws);
    }
    delete e;
}

int64_t emb_param_count(const EmbEngine* e) { return e ? e->n_params : 0; }
int64_t emb_buffer_count(const EmbEngine* e) { return e ? e->n_buffers : 0; }
int64_t emb_workspace_bytes(const EmbEngine* e) { return e ? e->ws_bytes : 0; }
int32_t emb_num_tensors(const EmbEngine* e) { return e ? (int32_t)e->tensors.size() : 0; }

int emb_param_info(const EmbEngine* e, int32_t i, EmbParamInfo* out) {
    if (!e || !out || i < 0 || i >= (int)e->tensors.size()) return set_error(EMB_E_ARG, "bad tensor index");
    const Tensor& t = e->tensors[i];
    memset(out, 0, sizeof *out);
    strncpy(out->name, t.name.c_str(), sizeof(out->name) - 1);
    out->offset = t.offset; out->numel = t.numel; out->ndim = t.ndim;
    for (int k = 0; k < 3; ++k) out->shape[k] = t.shape[k];
    out->is_buffer = t.is_buffer;
    return EMB_OK;
}

int32_t emb_output_size(const EmbEngine* e, int32_t which) { return !e ? 0 : (which == 0 ? e->ffnn_out : e->cnn_out); }

int emb_bind(EmbEngine* e, float* params, float* grads, float* buffers, float* opt_m, float* opt_v, void* workspace, int64_t workspace_bytes) {
    if (!e) return set_error(EMB_E_ARG, "null engine");
    if (emb_device_count() < 1) return set_error(EMB_E_NO_DEVICE, "no sm_100 (B200) device visible: this library has no CPU path");
    if (!params && !workspace) {
        const int64_t np = std::max<int64_t>(e->n_params, 4), nb = std::max<int64_t>(e->n_buffers, 4);
        EMB_CUDA_OK(cudaMalloc(&e->params, np * sizeof(float)));
        EMB_CUDA_OK(cudaMalloc(&e->grads, np * sizeof(float)));
        EMB_CUDA_OK(cudaMalloc(&e->buffers, nb * sizeof(float)));
        EMB_CUDA_OK(cudaMalloc(&e->opt_m, np * sizeof(float)));
        EMB_CUDA_OK(cudaMalloc(&e->opt_v, np * sizeof(float)));
        EMB_CUDA_OK(cudaMalloc(&e->ws, e->ws_bytes));
        EMB_CUDA_OK(cudaMemset(e->params, 0, np * sizeof(float)));
        EMB_CUDA_OK(cudaMemset(e->grads, 0, np * sizeof(float)));
        EMB_CUDA_OK(cudaMemset(e->buffers, 0, nb * sizeof(float)));
        EMB_CUDA_OK(cudaMemset(e->opt_m, 0, np * sizeof(float)));
        EMB_CUDA_OK(cudaMemset(e->opt_v, 0, np * sizeof(float)));
        e->owns_memory = true;
    } else {
        if (!params || !workspace) return set_error(EMB_E_ARG, "params and workspace must both be given");
        if (workspace_bytes < e->ws_bytes) return set_error(EMB_E_ARG, "workspace too small: %lld < %lld", (long long)workspace_bytes, (long long)e->ws_bytes);
        if (((uintptr_t)workspace & 255) || ((uintptr_t)params & 15)) return set_error(EMB_E_ARG, "workspace must be 256-byte and params 16-byte aligned");
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, workspace) == cudaSuccess && pa.type == cudaMemoryTypeDevice) EMB_CUDA_OK(cudaSetDevice(pa.device));
        e->params = params; e->grads = grads; e->buffers = buffers; e->opt_m = opt_m; e->opt_v = opt_v; e->ws = (char*)workspace;
    }
    EMB_CUDA_OK(cudaGetDevice(&e->device));
    if (!e->side[0]) {
        for (int k = 0; k < 2; ++k) EMB_CUDA_OK(cudaStreamCreateWithFlags(&e->side[k], cudaStreamNonBlocking));
        e->fork_ev.resize(48);
        for (auto& ev : e->fork_ev) EMB_CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    }
    carve(e, e->ws);
    if (e->prec == EMB_PREC_BF16) { int rcw = build_wcache_table(e); if (rcw) return rcw; }
    RngState rs{e->seed, 0};
    EMB_CUDA_OK(cudaMemcpy(e->rng, &rs, sizeof rs, cudaMemcpyHostToDevice));
    EMB_CUDA_OK(cudaMemset(e->rec_count, 0, sizeof(int)));
    EMB_CUDA_OK(cudaMemset(e->dp_epoch, 0, sizeof(unsigned int)));
    e->dp.epoch = e->dp_epoch;
    cudaFuncSetAttribute(pool_bn_bwd_stage1_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * SEQ_LEN * 32 * 4);
    cudaFuncSetAttribute(pool_bn_bwd_stage1_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * SEQ_LEN * 32 * 4);
    cudaFuncSetAttribute(onehot_conv_bwd_lists_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (SEQ_LEN + 16) * 64 * 4 + 512);
    cudaFuncSetAttribute(onehot_conv_bwd_lists_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (SEQ_LEN + 16) * 64 * 4 + 512);
    cudaFuncSetAttribute(onehot_conv_fwd_pair_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (8 * 25 * 64 + 64 + 4096) * 4 + 512);
    cudaFuncSetAttribute(onehot_conv_fwd_pair_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (8 * 25 * 64 + 64 + 4096) * 4 + 512);
    int rc = tc_init();
    if (rc) return rc;
    EMB_CUDA_OK(cudaFuncSetAttribute(onehot_conv_fwd_triple_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_max_smem()));
    EMB_CUDA_OK(cudaFuncSetAttribute(onehot_conv_fwd_triple_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_max_smem()));
    EMB_CUDA_OK(cudaFuncSetAttribute(pool_bn_bwd_tma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kt_max_smem()));
    EMB_CUDA_OK(cudaFuncSetAttribute(pool_bn_bwd_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kt_max_smem()));
    return EMB_OK;
}

int emb_set_seed(EmbEngine* e, uint64_t seed) {
    if (!e) return set_error(EMB_E_ARG, "null engine");
    e->seed = seed;
    if (e->rng) {
        RngState rs{seed, 0};
        EMB_CUDA_OK(cudaMemcpy(e->rng, &rs, sizeof rs, cudaMemcpyHostToDevice));
    }
    return EMB_OK;
}

int emb_set_shard(EmbEngine* e, int64_t row_offset, int64_t global_batch) {
    if (!e) return set_error(EMB_E_ARG, "null engine");
    e->row_offset = row_offset;
    e->global_batch = global_batch;
    return EMB_OK;
}

int emb_set_global_positives(EmbEngine* e, int64_t n_pos_global) {
    if (!e) return set_error(EMB_E_ARG, "null engine");
    e->global_pos = n_pos_global;
    return EMB_OK;
}

int emb_set_allreduce(EmbEngine* e, EmbAllreduceFn fn, void* user) {
    if (!e) return set_error(EMB_E_ARG, "null engine");
    e->allreduce = fn;
    e->allreduce_user = user;
    return EMB_OK;
}

int emb_set_phase_hook(EmbEngine* e, EmbPhaseFn fn, void* user) {
    if (!e) return set_error(EMB_E_ARG, "null engine");
    e->phase_hook = fn;
    e->phase_user = user;
    return EMB_OK;
}

// ---- data parallelism over peer memory (csrc/dp_peer.cuh) ------------------------------------------------------------------
static int dp_reduce_launch(EmbEngine* e, cudaStream_t st, int do_opt, int write_grads);
int64_t emb_dp_comm_bytes(void) { return (int64_t)sizeof(DpComm); }

int emb_dp_attach(EmbEngine* e, int32_t rank, int32_t world, void* const* comm, float* const* params, float* const* grads) {
    int rc = check_ready(e, 1, true);
    if (rc) return rc;
    if (world < 1 || world > DP_MAX_WORLD || rank < 0 || rank >= world) return set_error(EMB_E_ARG, "bad rank/world %d/%d (world <= %d)", rank, world, DP_MAX_WORLD);
    if (!comm || !params || !grads) return set_error(EMB_E_ARG, "null pointer table");
    if (params[rank] != e->params || grads[rank] != e->grads) return set_error(EMB_E_ARG, "entry [rank] must be this engine's own arenas");
    for (auto& c : e->cnn)
        if (2 * c.cout > DP_SLOT_DOUBLES) return set_error(EMB_E_UNSUPPORTED, "data parallel SyncBN supports up to %d channels", DP_SLOT_DOUBLES / 2);
    if (e->cnn.size() > (size_t)EMB_MAX_CNN) return set_error(EMB_E_UNSUPPORTED, "too many conv layers");
    {   // CUDA loads kernels lazily, and loading one may synchronise the context: a kernel that is first launched while an
        // exchange kernel of this device is waiting for a peer ON THE SAME DEVICE (the single-GPU test) would deadlock.  Load now.
        cudaFuncAttributes fa;
        EMB_CUDA_OK(cudaFuncGetAttributes(&fa, dp_epoch_advance_kernel));
        EMB_CUDA_OK(cudaFuncGetAttributes(&fa, dp_barrier_kernel));
        EMB_CUDA_OK(cudaFuncGetAttributes(&fa, bn_finalize_dp_kernel));
        EMB_CUDA_OK(cudaFuncGetAttributes(&fa, bn_bwd_finalize_dp_kernel));
        EMB_CUDA_OK(cudaFuncGetAttributes(&fa, dp_count_positives_kernel));
        EMB_CUDA_OK(cudaFuncGetAttributes(&fa, dp_reduce_opt_kernel));
        EMB_CUDA_OK(cudaFuncGetAttributes(&fa, ce_loss_kernel));
    }
    e->dp.rank = rank; e->dp.world = world; e->dp.epoch = e->dp_epoch;
    for (int q = 0; q < world; ++q) {
        if (!comm[q] || !params[q] || !grads[q]) return set_error(EMB_E_ARG, "null pointer for rank %d", q);
        if (((uintptr_t)comm[q] & 127) || ((uintptr_t)params[q] & 15) || ((uintptr_t)grads[q] & 15)) return set_error(EMB_E_ARG, "misaligned peer pointer (rank %d)", q);
        e->dp.comm[q] = (DpComm*)comm[q];
        e->dp_ar.params[q] = params[q];
        e->dp_ar.grads[q] = grads[q];
    }
    // contiguous, 16-byte aligned slices of the arena: rank r owns [r * per, (r + 1) * per)
    const int64_t n = round_up64(e->n_params, 4);
    const int64_t per = round_up64((n + world - 1) / world, 4);
    e->dp_lo = std::min<int64_t>(n, (int64_t)rank * per);
    e->dp_hi = std::min<int64_t>(n, e->dp_lo + per);
    // a fresh group starts at epoch 0 with clean flags on every member (the caller zeroes the comm blocks before attaching)
    EMB_CUDA_OK(cudaMemset(e->dp_epoch, 0, sizeof(unsigned int)));
    for (auto& g : e->graphs) cudaGraphExecDestroy(g.exec);       // graphs captured before attaching do not contain the exchanges
    e->graphs.clear();
    e->graph_seen.clear();
    e->dp_on = world > 1;
    return EMB_OK;
}

int emb_dp_detach(EmbEngine* e) {
    if (!e) return set_error(EMB_E_ARG, "null engine");
    e->dp_on = false;
    for (auto& g : e->graphs) cudaGraphExecDestroy(g.exec);
    e->graphs.clear();
    e->graph_seen.clear();
    return EMB_OK;
}

int emb_dp_allreduce_grads(EmbEngine* e, void* stream) {
    int rc = check_ready(e, 1, true);
    if (rc) return rc;
    if (!e->dp_on) return EMB_OK;
    return dp_reduce_launch(e, (cudaStream_t)stream, 0, 1);
}

// CUDA IPC plumbing for the peer pointers (one process per GPU).  The handle names the ALLOCATION that contains dev_ptr
// (PyTorch's caching allocator packs tensors into larger cudaMalloc blocks), *offset_out the position inside it.
int emb_ipc_export(const void* dev_ptr, void* handle_out, int64_t* offset_out) {
    if (!dev_ptr || !handle_out || !offset_out) return set_error(EMB_E_ARG, "null argument");
    CUdeviceptr base = 0;
    size_t size = 0;
    typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
    static RangeFn range_fn = nullptr;
    if (!range_fn) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
            return set_error(EMB_E_CUDA, "cuMemGetAddressRange not available");
        range_fn = (RangeFn)fn;
    }
    if (range_fn(&base, &size, (CUdeviceptr)dev_ptr) != CUDA_SUCCESS) return set_error(EMB_E_CUDA, "cuMemGetAddressRange failed");
    cudaIpcMemHandle_t h;
    cudaError_t err = cudaIpcGetMemHandle(&h, (void*)base);
    if (err != cudaSuccess) {
        cudaGetLastError();
        return set_error(EMB_E_CUDA, "cudaIpcGetMemHandle: %s (memory must come from cudaMalloc: no expandable segments / cudaMallocAsync)", cudaGetErrorString(err));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    memcpy(handle_out, &h, 64);
    *offset_out = (int64_t)((CUdeviceptr)dev_ptr - base);
    return EMB_OK;
}

int emb_ipc_open(const void* handle, int64_t offset, void** dev_ptr_out) {
    if (!handle || !dev_ptr_out) return set_error(EMB_E_ARG, "null argument");
    struct Opened { char h[64]; int dev; void* base; };
    static std::vector<Opened> opened;                         // a handle can be opened once per process and device: cache the mapping
    int dev = 0;
    cudaGetDevice(&dev);
    for (auto& o : opened)
        if (o.dev == dev && !memcmp(o.h, handle, 64)) { *dev_ptr_out = (char*)o.base + offset; return EMB_OK; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* base = nullptr;
    cudaError_t err = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
    if (err != cudaSuccess) { cudaGetLastError(); return set_error(EMB_E_CUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(err)); }
    Opened o;
    memcpy(o.h, handle, 64);
    o.dev = dev; o.base = base;
    opened.push_back(o);
    *dev_ptr_out = (char*)base + offset;
    return EMB_OK;
}

int emb_set_tensor_core(EmbEngine* e, int32_t on) {
    if (!e) return set_error(EMB_E_ARG, "null engine");
    if (on && e->prec != EMB_PREC_BF16) return set_error(EMB_E_UNSUPPORTED, "tensor-core GEMMs need EMB_PREC_BF16");
    e->use_tc = on != 0;
    return EMB_OK;
}

int64_t emb_launch_count(const EmbEngine* e) { return e ? e->launches : 0; }

int emb_opt_state_get(const EmbEngine* e, int64_t* step_out, double* nadam_mu_product_out) {
    if (!e) return set_error(EMB_E_ARG, "null engine");
    if (step_out) *step_out = e->opt_t;
    if (nadam_mu_product_out) *nadam_mu_product_out = e->nadam_mu_product;
    return EMB_OK;
}

int emb_opt_state_set(EmbEngine* e, int64_t step, double nadam_mu_product) {
    if (!e || step < 0) return set_error(EMB_E_ARG, "bad optimizer state");
    e->opt_t = step;
    e->nadam_mu_product = nadam_mu_product;
    return EMB_OK;
}

int emb_profile_gemm(EmbEngine* e, int32_t enable) {
    if (!e) return set_error(EMB_E_ARG, "null engine");
    e->prof_on = enable != 0;
    e->prof_used = 0;
    e->prof_flops = 0;
    e->prof_flops_each.clear();
    e->prof_launches = 0;
    return EMB_OK;
}

int emb_profile_read(EmbEngine* e, double* ms_out, double* flops_out, int64_t* launches_out) {
    if (!e) return set_error(EMB_E_ARG, "null engine");
    EMB_CUDA_OK(cudaDeviceSynchronize());
    double ms = 0;
    for (size_t i = 0; i + 1 < e->prof_used; i += 2) {
        float t = 0;
        EMB_CUDA_OK(cudaEventElapsedTime(&t, e->prof_ev[i], e->prof_ev[i + 1]));
        ms += t;
        if (tuning().prof_dump && i / 2 < e->prof_flops_each.size())
            fprintf(stderr, "gemm %3zu  %9.1f us  %8.2f GFLOP  %7.1f TFLOP/s\n", i / 2, t * 1e3, e->prof_flops_each[i / 2] / 1e9,
                    e->prof_flops_each[i / 2] / (t * 1e-3) / 1e12);
    }
    if (ms_out) *ms_out = ms;
    if (flops_out) *flops_out = e->prof_flops;
    if (launches_out) *launches_out = e->prof_launches;
    return EMB_OK;
}

int emb_forward_train(EmbEngine* e, const float* x_ffnn, const uint8_t* bases, const float* availabilities, int32_t B,
                      const EmbDraws* draws, float* logits_out, void* stream) {
    int rc = check_ready(e, B, false);
    if (rc) return rc;
    return forward_impl(e, x_ffnn, bases, availabilities, B, true, draws, logits_out, (cudaStream_t)stream);
}

int emb_forward_infer(EmbEngine* e, const float* x_ffnn, const uint8_t* bases, const float* availabilities, int32_t B,
                      const EmbDraws* draws, float* logits_out, float* probs_out, void* stream) {
    int rc = check_ready(e, B, false);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    float* lg = logits_out ? logits_out : e->logits;
    rc = forward_impl(e, x_ffnn, bases, availabilities, B, false, draws, lg, st);
    if (rc) return rc;
    if (probs_out) {
        softmax_p1_kernel<<<cdiv(B, 256), 256, 0, st>>>(lg, probs_out, B);
        EMB_CHECK_LAUNCH();
        LAUNCHED(e);
    }
    return EMB_OK;
}

int emb_loss_ce_weighted(EmbEngine* e, const float* logits, const int32_t* labels, int32_t B, float* dlogits_out, void* stream) {
    int rc = check_ready(e, B, false);
    if (rc) return rc;
    if (!logits || !labels) return set_error(EMB_E_ARG, "null logits/labels");
    if (e->dp_on && e->last_training && e->global_batch > 0) {
        // data parallel: the class weights come from the GLOBAL positive count, exchanged over peer memory by a one-CTA kernel
        dp_count_positives_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(e->dp, labels, B, e->dp_gpos);
        EMB_CHECK_LAUNCH();
        LAUNCHED(e);
        ce_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(logits, labels, B, 0, e->global_batch, dlogits_out, e->rec, e->rec_count,
                                                             EmbEngine::MAX_REC, e->dp_gpos);
        EMB_CHECK_LAUNCH();
        LAUNCHED(e);
        return EMB_OK;
    }
    const bool sharded = e->global_batch > 0 && e->global_pos >= 0;
    ce_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(logits, labels, B, sharded ? e->global_pos : -1, sharded ? e->global_batch : -1,
                                                         dlogits_out, e->rec, e->rec_count, EmbEngine::MAX_REC,
                                                         sharded && e->graph_coll ? e->gpos_dev : nullptr);
    EMB_CHECK_LAUNCH();
    LAUNCHED(e);
    return EMB_OK;
}

int emb_backward(EmbEngine* e, const float* dlogits, void* stream) {
    int rc = check_ready(e, e ? std::max(1, e->last_B) : 1, true);
    if (rc) return rc;
    if (!dlogits) return set_error(EMB_E_ARG, "null dlogits");
    return backward_impl(e, dlogits, (cudaStream_t)stream);
}

// host half of an optimizer step: the step-dependent scalars travel through a small device slot written by a
// stream-ordered copy (outside any captured graph, so a replayed graph sees fresh values)
static int opt_step_prepare(EmbEngine* e, const EmbOptConfig* cfg, cudaStream_t st) {
    if (!cfg || cfg->kind < 0 || cfg->kind > 3) return set_error(EMB_E_ARG, "bad optimizer config");
    if (!e->opt_m || !e->opt_v) return set_error(EMB_E_STATE, "no optimizer state bound");
    OptScalars h;
    fill_opt_scalars(e, *cfg, &h);
    EMB_CUDA_OK(cudaMemcpyAsync(e->opt_scalars, &h, sizeof h, cudaMemcpyHostToDevice, st));
    return EMB_OK;
}
// data parallel: reduce-scatter of the gradient arenas + optimizer on this rank's slice + all-gather of the parameters
static int dp_reduce_launch(EmbEngine* e, cudaStream_t st, int do_opt, int write_grads) {
    dp_barrier_kernel<<<1, 32, 0, st>>>(e->dp, DP_SYNC_GRADS_READY);          // every rank's backward pass has finished
    EMB_CHECK_LAUNCH();
    const int64_t n4 = (e->dp_hi - e->dp_lo) / 4;
    if (n4 > 0) {
        const int grid = (int)std::min<int64_t>(tc_num_sms() * 4, cdiv(n4, 256));
        dp_reduce_opt_kernel<<<grid, 256, 0, st>>>(e->dp.rank, e->dp.world, e->dp_ar, e->opt_m, e->opt_v, e->opt_scalars, (long long)e->dp_lo,
                                                   (long long)e->dp_hi, do_opt, write_grads);
        EMB_CHECK_LAUNCH();
    }
    dp_barrier_kernel<<<1, 32, 0, st>>>(e->dp, DP_SYNC_PARAMS_DONE);          // everybody's stores have landed; my gradients are free again
    EMB_CHECK_LAUNCH();
    e->launches += 3;
    return EMB_OK;
}

static int opt_step_launch(EmbEngine* e, cudaStream_t st) {
    if (e->dp_on) return dp_reduce_launch(e, st, 1, 0);
    int grid = std::min<int64_t>(148 * 8, cdiv(e->n_params, 256));
    opt_step_kernel<<<grid, 256, 0, st>>>(e->params, e->grads, e->opt_m, e->opt_v, e->opt_scalars, (size_t)e->n_params);
    EMB_CHECK_LAUNCH();
    LAUNCHED(e);
    return EMB_OK;
}

int emb_opt_step(EmbEngine* e, const EmbOptConfig* cfg, void* stream) {
    int rc = check_ready(e, 1, true);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = opt_step_prepare(e, cfg, st))) return rc;
    return opt_step_launch(e, st);
}

int emb_set_graph(EmbEngine* e, int32_t on) {
    if (!e) return set_error(EMB_E_ARG, "null engine");
    e->graph_on = on != 0;
    e->graph_coll = on == 2;
    if (e->graph_coll && !e->gpos_dev) EMB_CUDA_OK(cudaMalloc(&e->gpos_dev, sizeof(int64_t)));
    if (!on) {
        for (auto& g : e->graphs) cudaGraphExecDestroy(g.exec);
        e->graphs.clear();
        e->graph_seen.clear();
    }
    return EMB_OK;
}

// The train step as ONE graph launch.  Inputs are staged into the engine's own buffers first (so the captured kernel
// arguments never change); everything that varies from step to step lives in device memory: the Philox step counter,
// the metrics record cursor and the optimizer scalars (written by opt_step_prepare before the launch).
static int train_step_graph(EmbEngine* e, const float* x_ffnn, const uint8_t* bases, const int32_t* labels, int32_t B,
                            const EmbOptConfig* cfg, cudaStream_t caller) {
    int rc;
    if (!e->gstream) {
        EMB_CUDA_OK(cudaStreamCreateWithFlags(&e->gstream, cudaStreamNonBlocking));
        EMB_CUDA_OK(cudaEventCreateWithFlags(&e->gev_in, cudaEventDisableTiming));
        EMB_CUDA_OK(cudaEventCreateWithFlags(&e->gev_out, cudaEventDisableTiming));
    }
    cudaStream_t st = e->gstream;
    EMB_CUDA_OK(cudaEventRecord(e->gev_in, caller));
    EMB_CUDA_OK(cudaStreamWaitEvent(st, e->gev_in, 0));
    if (e->spec.kind != EMB_KIND_CNN && x_ffnn != e->in_x)
        EMB_CUDA_OK(cudaMemcpyAsync(e->in_x, x_ffnn, (size_t)B * e->spec.in_features * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (e->spec.kind != EMB_KIND_FFNN && bases != e->in_bases)
        EMB_CUDA_OK(cudaMemcpyAsync(e->in_bases, bases, (size_t)B * SEQ_LEN, cudaMemcpyDeviceToDevice, st));
    if (labels != e->in_labels) EMB_CUDA_OK(cudaMemcpyAsync(e->in_labels, labels, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    if (cfg && (rc = opt_step_prepare(e, cfg, st))) return rc;
    if (e->graph_coll && e->global_pos >= 0) EMB_CUDA_OK(cudaMemcpyAsync(e->gpos_dev, &e->global_pos, sizeof(int64_t), cudaMemcpyHostToDevice, st));
    const int has_opt = cfg ? 1 : 0, kind = cfg ? cfg->kind : -1;
    EmbEngine::StepGraph* g = nullptr;
    for (auto& c : e->graphs) if (c.B == B && c.has_opt == has_opt && c.opt_kind == kind) g = &c;
    if (!g) {
        const int64_t l0 = e->launches;
        EMB_CUDA_OK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        e->capturing = true;
        rc = forward_impl(e, e->in_x, e->in_bases, nullptr, B, true, nullptr, e->logits, st);
        if (!rc) rc = emb_loss_ce_weighted(e, e->logits, e->in_labels, B, e->dlogits, (void*)st);
        if (!rc) rc = backward_impl(e, e->dlogits, st);
        if (!rc && e->graph_coll && e->phase_hook && e->phase_hook(e->phase_user, 2, (void*)st)) rc = set_error(EMB_E_STATE, "phase hook (2) failed");
        if (!rc && cfg) rc = opt_step_launch(e, st);
        e->capturing = false;
        cudaGraph_t graph = nullptr;
        cudaError_t err = cudaStreamEndCapture(st, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (err != cudaSuccess) return set_error(EMB_E_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(err));
        EmbEngine::StepGraph ng{B, has_opt, kind, e->launches - l0, nullptr};
        e->launches = l0;
        err = cudaGraphInstantiate(&ng.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (err != cudaSuccess) return set_error(EMB_E_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(err));
        e->graphs.push_back(ng);
        g = &e->graphs.back();
    }
    EMB_CUDA_OK(cudaGraphLaunch(g->exec, st));
    EMB_CUDA_OK(cudaEventRecord(e->gev_out, st));
    EMB_CUDA_OK(cudaStreamWaitEvent(caller, e->gev_out, 0));
    e->launches += g->kernels;
    e->last_B = B; e->last_training = true; e->last_bases = e->in_bases;
    return EMB_OK;
}

int emb_train_step(EmbEngine* e, const float* x_ffnn, const uint8_t* bases, const int32_t* labels, int32_t B,
                   const EmbDraws* draws, const EmbOptConfig* cfg, void* stream) {
    int rc = check_ready(e, B, true);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (e->graph_on && !draws && !e->prof_on && (e->graph_coll || e->dp_on || (!e->allreduce && !e->phase_hook && e->global_batch < 0))) {
        // the first step of a batch size runs eagerly (one-time initialisation such as function attributes happens there)
        bool seen = false;
        for (int b : e->graph_seen) seen |= b == B;
        if (seen) return train_step_graph(e, x_ffnn, bases, labels, B, cfg, st);
        e->graph_seen.push_back(B);
    }
    if (e->graph_coll && e->global_pos >= 0) EMB_CUDA_OK(cudaMemcpyAsync(e->gpos_dev, &e->global_pos, sizeof(int64_t), cudaMemcpyHostToDevice, st));
    rc = forward_impl(e, x_ffnn, bases, nullptr, B, true, draws, e->logits, st);
    if (rc) return rc;
    rc = emb_loss_ce_weighted(e, e->logits, labels, B, e->dlogits, stream);
    if (rc) return rc;
    rc = backward_impl(e, e->dlogits, st);
    if (rc) return rc;
    // graph-with-collectives mode: the host finishes its gradient all-reduce inside the step (phase 2), before the optimizer
    if (e->graph_on && e->graph_coll && e->phase_hook && e->phase_hook(e->phase_user, 2, stream)) return set_error(EMB_E_STATE, "phase hook (2) failed");
    if (cfg) rc = emb_opt_step(e, cfg, stream);
    return rc;
}

int emb_train_step_host(EmbEngine* e, const float* x_ffnn_host, const uint8_t* bases_host, const int32_t* labels_host, int32_t B,
                        const EmbOptConfig* cfg, EmbStepMetrics* metrics_host, void* stream) {
    int rc = check_ready(e, B, true);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (e->spec.kind != EMB_KIND_CNN) EMB_CUDA_OK(cudaMemcpyAsync(e->in_x, x_ffnn_host, (size_t)B * e->spec.in_features * sizeof(float), cudaMemcpyHostToDevice, st));
    if (e->spec.kind != EMB_KIND_FFNN) EMB_CUDA_OK(cudaMemcpyAsync(e->in_bases, bases_host, (size_t)B * SEQ_LEN, cudaMemcpyHostToDevice, st));
    EMB_CUDA_OK(cudaMemcpyAsync(e->in_labels, labels_host, (size_t)B * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    EMB_CUDA_OK(cudaMemsetAsync(e->rec_count, 0, sizeof(int), st));
    rc = emb_train_step(e, e->in_x, e->in_bases, e->in_labels, B, nullptr, cfg, stream);
    if (rc) return rc;
    if (metrics_host) {
        EMB_CUDA_OK(cudaMemcpyAsync(metrics_host, e->rec, sizeof(EmbStepMetrics), cudaMemcpyDeviceToHost, st));
        EMB_CUDA_OK(cudaStreamSynchronize(st));
    }
    return EMB_OK;
}

__global__ void set_int_kernel(int* p, int v) { *p = v; }

// Software-pipelined host entry.  Call i copies batch i host->device on a COPY stream into staging slot i & 1 (it overlaps
// the compute of step i-1, which is still running), enqueues step i on `stream`, and returns the metrics of step i-1
// (*have_metrics = 0 on the first call).  emb_train_step_host_flush() returns the last step's metrics.  Every step's
// inputs cross PCIe/NVLink-C2C and every step's record is read back, but the device never waits for the host.
int emb_train_step_host_pipelined(EmbEngine* e, const float* x_ffnn_host, const uint8_t* bases_host, const int32_t* labels_host, int32_t B,
                                  const EmbOptConfig* cfg, EmbStepMetrics* metrics_prev_host, int32_t* have_metrics, void* stream) {
    int rc = check_ready(e, B, true);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (!e->copy_stream) {
        EMB_CUDA_OK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) {
            EMB_CUDA_OK(cudaEventCreateWithFlags(&e->ev_h2d[k], cudaEventDisableTiming));
            EMB_CUDA_OK(cudaEventCreateWithFlags(&e->ev_consumed[k], cudaEventDisableTiming));
            EMB_CUDA_OK(cudaEventCreateWithFlags(&e->ev_step[k], cudaEventDisableTiming));
        }
        EMB_CUDA_OK(cudaEventCreateWithFlags(&e->ev_d2h, cudaEventDisableTiming));
        EMB_CUDA_OK(cudaMallocHost(&e->pin_metrics, 2 * sizeof(EmbStepMetrics)));
    }
    const int64_t i = e->pipe_steps;
    const int k = (int)(i & 1);
    cudaStream_t cs = e->copy_stream;
    // staging slot k was last read by the device-to-device copy of step i-2
    if (i >= 2) EMB_CUDA_OK(cudaStreamWaitEvent(cs, e->ev_consumed[k], 0));
    if (e->spec.kind != EMB_KIND_CNN) EMB_CUDA_OK(cudaMemcpyAsync(e->stg_x[k], x_ffnn_host, (size_t)B * e->spec.in_features * sizeof(float), cudaMemcpyHostToDevice, cs));
    if (e->spec.kind != EMB_KIND_FFNN) EMB_CUDA_OK(cudaMemcpyAsync(e->stg_bases[k], bases_host, (size_t)B * SEQ_LEN, cudaMemcpyHostToDevice, cs));
    EMB_CUDA_OK(cudaMemcpyAsync(e->stg_labels[k], labels_host, (size_t)B * sizeof(int32_t), cudaMemcpyHostToDevice, cs));
    EMB_CUDA_OK(cudaEventRecord(e->ev_h2d[k], cs));
    // compute stream: wait for the copy, move the batch into the engine's fixed input buffers (what a captured graph reads)
    EMB_CUDA_OK(cudaStreamWaitEvent(st, e->ev_h2d[k], 0));
    if (e->spec.kind != EMB_KIND_CNN) EMB_CUDA_OK(cudaMemcpyAsync(e->in_x, e->stg_x[k], (size_t)B * e->spec.in_features * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (e->spec.kind != EMB_KIND_FFNN) EMB_CUDA_OK(cudaMemcpyAsync(e->in_bases, e->stg_bases[k], (size_t)B * SEQ_LEN, cudaMemcpyDeviceToDevice, st));
    EMB_CUDA_OK(cudaMemcpyAsync(e->in_labels, e->stg_labels[k], (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    EMB_CUDA_OK(cudaEventRecord(e->ev_consumed[k], st));
    set_int_kernel<<<1, 1, 0, st>>>(e->rec_count, k);          // this step's record lands in rec[k]
    EMB_CHECK_LAUNCH();
    rc = emb_train_step(e, e->in_x, e->in_bases, e->in_labels, B, nullptr, cfg, stream);
    if (rc) return rc;
    EMB_CUDA_OK(cudaEventRecord(e->ev_step[k], st));
    e->pipe_steps = i + 1;
    // deliver the record of step i-1: its device-to-host copy rides the copy stream behind this step's host-to-device copy
    if (have_metrics) *have_metrics = 0;
    if (e->pipe_pending) {
        const int kp = k ^ 1;
        EMB_CUDA_OK(cudaStreamWaitEvent(cs, e->ev_step[kp], 0));
        EMB_CUDA_OK(cudaMemcpyAsync(&e->pin_metrics[kp], e->rec + kp, sizeof(EmbStepMetrics), cudaMemcpyDeviceToHost, cs));
        EMB_CUDA_OK(cudaEventRecord(e->ev_d2h, cs));
        EMB_CUDA_OK(cudaEventSynchronize(e->ev_d2h));
        if (metrics_prev_host) *metrics_prev_host = e->pin_metrics[kp];
        if (have_metrics) *have_metrics = 1;
    }
    e->pipe_pending = true;
    return EMB_OK;
}

int emb_train_step_host_flush(EmbEngine* e, EmbStepMetrics* metrics_host, int32_t* have_metrics, void* stream) {
    int rc = check_ready(e, 1, false);
    if (rc) return rc;
    if (have_metrics) *have_metrics = 0;
    if (!e->pipe_pending) return EMB_OK;
    const int kp = (int)((e->pipe_steps - 1) & 1);
    cudaStream_t cs = e->copy_stream;
    EMB_CUDA_OK(cudaStreamWaitEvent(cs, e->ev_step[kp], 0));
    EMB_CUDA_OK(cudaMemcpyAsync(&e->pin_metrics[kp], e->rec + kp, sizeof(EmbStepMetrics), cudaMemcpyDeviceToHost, cs));
    EMB_CUDA_OK(cudaStreamSynchronize(cs));
    if (metrics_host) *metrics_host = e->pin_metrics[kp];
    if (have_metrics) *have_metrics = 1;
    e->pipe_pending = false;
    (void)stream;
    return EMB_OK;
}

// Host-buffer scoring, synchronous: upload, forward, scores back, stream synchronised.  (Cutting one call into chunks whose uploads
// overlap the previous chunk's compute was measured and dropped: four quarter-size forwards cost what the overlap saved, 1.89 vs
// 1.92 ms for 65 536 regions.  A scoring LOOP overlaps whole batches instead: emb_predict_host_pipelined below.)
int emb_predict_host(EmbEngine* e, const float* x_ffnn_host, const uint8_t* bases_host, const float* availabilities_host, int32_t B,
                     float* probs_host, void* stream) {
    int rc = check_ready(e, B, false);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (e->spec.kind != EMB_KIND_CNN) EMB_CUDA_OK(cudaMemcpyAsync(e->in_x, x_ffnn_host, (size_t)B * e->spec.in_features * sizeof(float), cudaMemcpyHostToDevice, st));
    if (e->spec.kind != EMB_KIND_FFNN) EMB_CUDA_OK(cudaMemcpyAsync(e->in_bases, bases_host, (size_t)B * SEQ_LEN, cudaMemcpyHostToDevice, st));
    if (availabilities_host) EMB_CUDA_OK(cudaMemcpyAsync(e->in_avail, availabilities_host, (size_t)B * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
    rc = emb_forward_infer(e, e->in_x, e->in_bases, availabilities_host ? e->in_avail : nullptr, B, nullptr, e->logits, e->probs, stream);
    if (rc) return rc;
    EMB_CUDA_OK(cudaMemcpyAsync(probs_host, e->probs, (size_t)B * sizeof(float), cudaMemcpyDeviceToHost, st));
    EMB_CUDA_OK(cudaStreamSynchronize(st));
    return EMB_OK;
}

// Software-pipelined scoring loop (the inference twin of emb_train_step_host_pipelined).  Call i uploads batch i on the engine's
// copy stream into staging slot i & 1 -- while batch i-1 is still being computed -- enqueues its forward and the device-to-host
// copy of its scores into `probs_host` on `stream`, and returns once batch i-1 is COMPLETE (*prev_done = 1: the buffer given to
// the previous call now holds its scores).  emb_predict_host_flush waits for the last batch.  Host buffers should be pinned.
int emb_predict_host_pipelined(EmbEngine* e, const float* x_ffnn_host, const uint8_t* bases_host, const float* availabilities_host, int32_t B,
                               float* probs_host, int32_t* prev_done, void* stream) {
    int rc = check_ready(e, B, false);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (!e->infer_copy) {
        EMB_CUDA_OK(cudaStreamCreateWithFlags(&e->infer_copy, cudaStreamNonBlocking));
        for (auto& ev : e->infer_ev) EMB_CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    }
    cudaStream_t cs = e->infer_copy;
    const int k = (int)(e->infer_calls & 1);
    const int F = e->spec.in_features;
    // slot k was read by batch i-2, which the previous call already waited for: the upload can start at once
    if (e->spec.kind != EMB_KIND_CNN) EMB_CUDA_OK(cudaMemcpyAsync(e->stg_x[k], x_ffnn_host, (size_t)B * F * sizeof(float), cudaMemcpyHostToDevice, cs));
    if (e->spec.kind != EMB_KIND_FFNN) EMB_CUDA_OK(cudaMemcpyAsync(e->stg_bases[k], bases_host, (size_t)B * SEQ_LEN, cudaMemcpyHostToDevice, cs));
    if (availabilities_host) EMB_CUDA_OK(cudaMemcpyAsync(e->stg_avail[k], availabilities_host, (size_t)B * 2 * sizeof(float), cudaMemcpyHostToDevice, cs));
    EMB_CUDA_OK(cudaEventRecord(e->infer_ev[k], cs));
    EMB_CUDA_OK(cudaStreamWaitEvent(st, e->infer_ev[k], 0));
    rc = emb_forward_infer(e, e->stg_x[k], e->stg_bases[k], availabilities_host ? e->stg_avail[k] : nullptr, B, nullptr, e->logits, e->probs, stream);
    if (rc) return rc;
    EMB_CUDA_OK(cudaMemcpyAsync(probs_host, e->probs, (size_t)B * sizeof(float), cudaMemcpyDeviceToHost, st));
    EMB_CUDA_OK(cudaEventRecord(e->infer_ev[2 + k], st));
    if (prev_done) *prev_done = 0;
    if (e->infer_pending) {
        EMB_CUDA_OK(cudaEventSynchronize(e->infer_ev[2 + (k ^ 1)]));
        if (prev_done) *prev_done = 1;
    }
    e->infer_pending = true;
    e->infer_calls += 1;
    return EMB_OK;
}

int emb_predict_host_flush(EmbEngine* e, void* stream) {
    int rc = check_ready(e, 1, false);
    if (rc) return rc;
    (void)stream;
    if (e->infer_pending) EMB_CUDA_OK(cudaEventSynchronize(e->infer_ev[2 + (int)((e->infer_calls - 1) & 1)]));
    e->infer_pending = false;
    return EMB_OK;
}

int emb_metrics_reset(EmbEngine* e, void* stream) {
    int rc = check_ready(e, 1, false);
    if (rc) return rc;
    EMB_CUDA_OK(cudaMemsetAsync(e->rec_count, 0, sizeof(int), (cudaStream_t)stream));
    return EMB_OK;
}

int emb_metrics_read(EmbEngine* e, EmbStepMetrics* out_host, int32_t max_records, void* stream) {
    int rc = check_ready(e, 1, false);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    int n = 0;
    EMB_CUDA_OK(cudaMemcpyAsync(&n, e->rec_count, sizeof(int), cudaMemcpyDeviceToHost, st));
    EMB_CUDA_OK(cudaStreamSynchronize(st));
    n = std::min(n, std::min(max_records, (int)EmbEngine::MAX_REC));
    if (n > 0 && out_host) {
        EMB_CUDA_OK(cudaMemcpyAsync(out_host, e->rec, (size_t)n * sizeof(EmbStepMetrics), cudaMemcpyDeviceToHost, st));
        EMB_CUDA_OK(cudaStreamSynchronize(st));
    }
    return n;
}

int emb_last_selection(EmbEngine* e, uint8_t* idx_out, int32_t B, void* stream) {
    int rc = check_ready(e, B, false);
    if (rc) return rc;
    if (e->spec.kind != EMB_KIND_EMBRACENET) return set_error(EMB_E_STATE, "not an EmbraceNet engine");
    EMB_CUDA_OK(cudaMemcpyAsync(idx_out, e->idx, (size_t)B * e->spec.embracement_size, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return EMB_OK;
}

// ---- single-kernel entry points ---------------------------------------------------------------
int emb_k_onehot_conv_fwd(const uint8_t* bases, const float* w, const float* bias, int32_t B, int32_t C1, int32_t k, int32_t precision,
                          void* y, void* stream) {
    if (emb_device_count() < 1) return set_error(EMB_E_NO_DEVICE, "no sm_100 device");
    if (B < 1 || C1 < 1 || k < 1 || !(k & 1)) return set_error(EMB_E_ARG, "bad shape");
    size_t smem = (size_t)(k * 4 * C1 + C1) * sizeof(float) + SEQ_LEN + k + 16;
    int grid = std::min(B, 148 * 8);
    int ld = round_up(C1, 8);
    if (precision) onehot_conv_fwd_kernel<bf16><<<grid, 256, smem, (cudaStream_t)stream>>>(bases, w, bias, (bf16*)y, B, C1, k, ld, 0);
    else onehot_conv_fwd_kernel<float><<<grid, 256, smem, (cudaStream_t)stream>>>(bases, w, bias, (float*)y, B, C1, k, ld, 0);
    EMB_CHECK_LAUNCH();
    return EMB_OK;
}

int emb_k_onehot_conv_bwd(const uint8_t* bases, const void* dy, int32_t B, int32_t C1, int32_t k, int32_t precision, float* dw, float* dbias,
                          void* stream) {
    if (emb_device_count() < 1) return set_error(EMB_E_NO_DEVICE, "no sm_100 device");
    if (B < 1 || C1 < 1 || k < 1 || !(k & 1) || C1 * k > OHB_MAXP * 256) return set_error(EMB_E_ARG, "bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    EMB_CUDA_OK(cudaMemsetAsync(dw, 0, (size_t)C1 * 4 * k * sizeof(float), st));
    EMB_CUDA_OK(cudaMemsetAsync(dbias, 0, (size_t)C1 * sizeof(float), st));
    int grid = std::min(B, 148 * 4);
    int ld = round_up(C1, 8);
    if (precision) onehot_conv_bwd_kernel<bf16><<<grid, 256, 0, st>>>(bases, (const bf16*)dy, dw, dbias, B, C1, k, ld);
    else onehot_conv_bwd_kernel<float><<<grid, 256, 0, st>>>(bases, (const float*)dy, dw, dbias, B, C1, k, ld);
    EMB_CHECK_LAUNCH();
    return EMB_OK;
}

// The tensor-core form of K1 (onehot_wgrad_tc.cuh): y bf16 [B, 256, C1]; stats (nullable) [2][C1] doubles are ACCUMULATED.
int emb_k_onehot_conv_fwd_tc(const uint8_t* bases, const float* w, const float* bias, int32_t B, int32_t C1, int32_t k, void* y_bf16, double* stats,
                             void* stream) {
    if (emb_device_count() < 1) return set_error(EMB_E_NO_DEVICE, "no sm_100 device");
    if (B < 1 || !onehot_fwd_tc_ok(y_bf16, bases, C1, k, C1)) return set_error(EMB_E_ARG, "bad shape or alignment");
    int rc = onehot_conv_fwd_tc(bases, w, bias, (bf16*)y_bf16, stats, B, C1, k, C1, (cudaStream_t)stream);
    return rc ? rc : EMB_OK;
}

// The tensor-core form of the same gradient (onehot_wgrad_tc.cuh): dy bf16 [B, 256, C1], C1 % 8 == 0, C1 <= 64, odd k <= 15.
int emb_k_onehot_conv_wgrad_tc(const uint8_t* bases, const void* dy_bf16, int32_t B, int32_t C1, int32_t k, float* dw, void* stream) {
    if (emb_device_count() < 1) return set_error(EMB_E_NO_DEVICE, "no sm_100 device");
    if (B < 1 || !onehot_wgrad_tc_ok(dy_bf16, bases, C1, k, C1)) return set_error(EMB_E_ARG, "bad shape or alignment");
    cudaStream_t st = (cudaStream_t)stream;
    EMB_CUDA_OK(cudaMemsetAsync(dw, 0, (size_t)C1 * 4 * k * sizeof(float), st));
    int rc = onehot_conv_wgrad_tc(bases, (const bf16*)dy_bf16, dw, B, C1, k, C1, st, 0);
    return rc ? rc : EMB_OK;
}

// One GEMM-shaped op of the step on either back end (0 = SIMT, 1 = tcgen05), fp32 in / fp32 out; the
// inputs are rounded to bf16 exactly as the bf16 precision does.  kind = TcKind; shapes as documented at tc_gemm():
//   0 linear fwd   a[M,K] b[N,K]            -> out[M,N]        3 conv fwd   a[B,L,Cin]  b = W[Cout,Cin,taps] -> out[B*L,Cout]
//   1 linear dgrad a[M,K] b[K,N]            -> out[M,N]        4 conv dgrad a[B,L,Cout] b = W[Cout,Cin,taps] -> out[B*L,Cin]
//   2 linear wgrad a[K,M] b[K,N]            -> out[M,N]        5 conv wgrad a[B,L,Cout] b = act[B,L,Cin]     -> out = dW[Cout,Cin,taps]
// (all inner widths must be multiples of 8).  Synchronises the stream.
static int k_gemm_impl(int32_t kind, int32_t backend, const float* a, const float* b, float* out, int32_t M, int32_t N, int32_t K,
                       int32_t B, int32_t L, int32_t Cin, int32_t Cout, int32_t taps, void* stream, int reps, float* ms_out) {
    if (emb_device_count() < 1) return set_error(EMB_E_NO_DEVICE, "no sm_100 device");
    cudaStream_t st = (cudaStream_t)stream;
    const bool conv = kind >= 3;
    const int pad = (taps - 1) / 2;
    size_t na, nb, nout;
    if (kind == 0) { na = (size_t)M * K; nb = (size_t)N * K; nout = (size_t)M * N; }
    else if (kind == 1) { na = (size_t)M * K; nb = (size_t)K * N; nout = (size_t)M * N; }
    else if (kind == 2) { na = (size_t)K * M; nb = (size_t)K * N; nout = (size_t)M * N; }
    else if (kind == 3) { na = (size_t)B * L * Cin; nb = (size_t)Cout * Cin * taps; nout = (size_t)B * L * Cout; }
    else if (kind == 4) { na = (size_t)B * L * Cout; nb = (size_t)Cout * Cin * taps; nout = (size_t)B * L * Cin; }
    else if (kind == 5) { na = (size_t)B * L * Cout; nb = (size_t)B * L * Cin; nout = (size_t)Cout * Cin * taps; }
    else return set_error(EMB_E_ARG, "bad kind");
    bf16 *a16 = nullptr, *b16 = nullptr;
    EMB_CUDA_OK(cudaMalloc(&a16, na * 2 + 16));
    EMB_CUDA_OK(cudaMalloc(&b16, nb * 2 + 16));
    cast_f32_kernel<bf16><<<cdiv(na, 256), 256, 0, st>>>(a, a16, na);
    const bool w_is_b = kind == 3 || kind == 4;
    if (w_is_b && backend == 1) wcache_conv_kernel<<<cdiv(nb, 256), 256, 0, st>>>(b, b16, Cout, Cin, taps, Cin);
    else cast_f32_kernel<bf16><<<cdiv(nb, 256), 256, 0, st>>>(b, b16, nb);
    EMB_CHECK_LAUNCH();
    Epilogue ep = {};
    ep.scale = 1.f;
    ep.out = out; ep.out_dtype = 0;
    if (kind == 2 || kind == 5) {
        EMB_CUDA_OK(cudaMemsetAsync(out, 0, nout * sizeof(float), st));
        ep.mode = EPI_ATOMIC; ep.ldo = N;
        if (kind == 5) { ep.map = MAP_CONV_W; ep.mapC = Cin; ep.map_taps = taps; }
    } else {
        ep.mode = EPI_LINEAR;
        ep.ldo = kind == 3 ? Cout : kind == 4 ? Cin : N;
    }
    int rc = EMB_OK;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (ms_out) { cudaEventCreate(&ev0); cudaEventCreate(&ev1); }
    for (int rep = 0; rep < reps && rc == EMB_OK; ++rep) {
    if (ms_out && rep == 1) cudaEventRecord(ev0, st);      // rep 0 is the warm-up
    if (backend == 1) {
        TcProblem pr = {};
        pr.kind = kind; pr.a = a16; pr.b = b16;
        pr.M = M; pr.N = N; pr.K = K; pr.B = B; pr.L = L; pr.Cin = Cin; pr.Cout = Cout; pr.taps = taps; pr.pad = pad;
        if (kind == 0) { pr.lda = K; pr.ldb = K; }
        else if (kind == 1) { pr.lda = K; pr.ldb = N; }
        else if (kind == 2) { pr.lda = M; pr.ldb = N; }
        else if (kind == 3) { pr.lda = Cin; pr.ldb = Cin; pr.M = B * L; pr.N = Cout; }
        else if (kind == 4) { pr.lda = Cout; pr.ldb = Cin; pr.M = B * L; pr.N = Cin; }
        else { pr.lda = Cout; pr.ldb = Cin; pr.M = Cout; pr.N = Cin; }
        rc = tc_dispatch(pr, ep, st);
    } else {
        Operand A, Bo;
        int gm, gn, gk, split = 1;
        if (kind == 0) { A = make_operand(a16, 1, OP_ROWMAJOR, K, M, K); Bo = make_operand(b16, 1, OP_ROWMAJOR, K, N, K); gm = M; gn = N; gk = K; }
        else if (kind == 1) { A = make_operand(a16, 1, OP_ROWMAJOR, K, M, K); Bo = make_operand(b16, 1, OP_TRANSPOSED, N, N, K); gm = M; gn = N; gk = K; }
        else if (kind == 2) { A = make_operand(a16, 1, OP_TRANSPOSED, M, M, K); Bo = make_operand(b16, 1, OP_TRANSPOSED, N, N, K); gm = M; gn = N; gk = K; split = 8; }
        else if (kind == 3) {
            A = make_operand(a16, 1, OP_CONV_SHIFT, Cin, B * L, taps * Cin); A.L = L; A.C = Cin; A.taps = taps; A.pad = pad; A.sign = 1;
            Bo = make_operand(b16, 1, OP_CONV_W_FWD, 0, Cout, taps * Cin); Bo.C = Cin; Bo.taps = taps;
            gm = B * L; gn = Cout; gk = taps * Cin;
        } else if (kind == 4) {
            A = make_operand(a16, 1, OP_CONV_SHIFT, Cout, B * L, taps * Cout); A.L = L; A.C = Cout; A.taps = taps; A.pad = pad; A.sign = -1;
            Bo = make_operand(b16, 1, OP_CONV_W_DGRAD, 0, Cin, taps * Cout); Bo.C = Cout; Bo.taps = taps; Bo.wrows = Cin;
            gm = B * L; gn = Cin; gk = taps * Cout;
        } else {
            A = make_operand(a16, 1, OP_TRANSPOSED, Cout, Cout, B * L);
            Bo = make_operand(b16, 1, OP_CONV_SHIFT_T, Cin, taps * Cin, B * L); Bo.L = L; Bo.C = Cin; Bo.taps = taps; Bo.pad = pad; Bo.sign = 1;
            gm = Cout; gn = taps * Cin; gk = B * L; split = 8;
        }
        cudaError_t err = launch_gemm_simt(A, Bo, ep, gm, gn, gk, split, st);
        if (err != cudaSuccess) rc = set_error(EMB_E_CUDA, "gemm launch: %s", cudaGetErrorString(err));
    }
    }
    if (ms_out) cudaEventRecord(ev1, st);
    cudaError_t serr = cudaStreamSynchronize(st);
    if (ms_out) {
        float ms = 0.f;
        if (serr == cudaSuccess && reps > 1) cudaEventElapsedTime(&ms, ev0, ev1);
        *ms_out = reps > 1 ? ms / (reps - 1) : 0.f;
        cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    }
    cudaFree(a16);
    cudaFree(b16);
    if (rc) return rc;
    if (serr != cudaSuccess) return set_error(EMB_E_CUDA, "emb_k_gemm: %s", cudaGetErrorString(serr));
    return EMB_OK;
}

int emb_k_gemm(int32_t kind, int32_t backend, const float* a, const float* b, float* out, int32_t M, int32_t N, int32_t K,
               int32_t B, int32_t L, int32_t Cin, int32_t Cout, int32_t taps, void* stream) {
    return k_gemm_impl(kind, backend, a, b, out, M, N, K, B, L, Cin, Cout, taps, stream, 1, nullptr);
}

// the same launch repeated `reps` times (the first is a warm-up), average device time per launch in *ms_out
// (CUDA events on `stream`); wgrad kinds keep accumulating into `out`, which is only meaningful for reps == 1
int emb_k_gemm_time(int32_t kind, int32_t backend, const float* a, const float* b, float* out, int32_t M, int32_t N, int32_t K,
                    int32_t B, int32_t L, int32_t Cin, int32_t Cout, int32_t taps, int32_t reps, float* ms_out, void* stream) {
    if (reps < 2 || !ms_out) return set_error(EMB_E_ARG, "reps >= 2 and ms_out required");
    return k_gemm_impl(kind, backend, a, b, out, M, N, K, B, L, Cin, Cout, taps, stream, reps, ms_out);
}

// test-only: UMMA descriptor row-shift probe (csrc/probe.cuh).  a, b fp32 host-visible device arrays:
//   mode 0: a [144,64], b [64,64] -> out[128,64] = a[shift:shift+128] @ b^T
//   mode 1: a [128,128] (k,m), b [144,64] (k,n) -> out[128,64] = sum_k a[k,m] * b[k+shift,n]
int emb_k_umma_shift_probe(int32_t mode, int32_t shift, int32_t use_base_offset, const float* a, const float* b, float* out, void* stream) {
    if (emb_device_count() < 1) return set_error(EMB_E_NO_DEVICE, "no sm_100 device");
    int rc = tc_init();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t na = mode == 0 ? 144 * 64 : 128 * 128, nb = mode == 0 ? 64 * 64 : 144 * 64;
    bf16 *a16, *b16;
    EMB_CUDA_OK(cudaMalloc(&a16, na * 2));
    EMB_CUDA_OK(cudaMalloc(&b16, nb * 2));
    cast_f32_kernel<bf16><<<cdiv(na, 256), 256, 0, st>>>(a, a16, na);
    cast_f32_kernel<bf16><<<cdiv(nb, 256), 256, 0, st>>>(b, b16, nb);
    CUtensorMap ma, mb;
    if (mode == 0) {
        if ((rc = make_map(&ma, a16, 64, 144, 1, 64, 144 * 64, 64, 144, 1))) return rc;
        if ((rc = make_map(&mb, b16, 64, 64, 1, 64, 64 * 64, 64, 64, 1))) return rc;
    } else {
        if ((rc = make_map(&ma, a16, 128, 128, 1, 128, 128 * 128, 64, 128, 1))) return rc;
        if ((rc = make_map(&mb, b16, 64, 144, 1, 64, 144 * 64, 64, 144, 1))) return rc;
    }
    cudaFuncSetAttribute(umma_shift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
    umma_shift_probe_kernel<<<1, 128, 68000, st>>>(ma, mb, mode, shift, use_base_offset, out);
    cudaError_t err = cudaStreamSynchronize(st);
    cudaFree(a16);
    cudaFree(b16);
    if (err != cudaSuccess) return set_error(EMB_E_CUDA, "probe: %s", cudaGetErrorString(err));
    return EMB_OK;
}

}  // extern "C"
