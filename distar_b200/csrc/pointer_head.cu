// Selected-units pointer network, one sampling step per launch (K12) — replaces the ~25 ATen launches per step of the
// python loop in SelectedUnitsHead._query (head/action_arg_head.py:262-306): query MLP (1024 -> 256 -> 32), the 32-wide
// LayerNorm-LSTM cell, dot products against the 513 keys, mask, temperature, softmax + multinomial (argmax(p/q)),
// end-flag / selected-count bookkeeping, masked mean of the selected keys and the embedding MLP (32 -> 256 -> 1024)
// that produces the next auto-regressive embedding.  One CTA per batch row; the row's whole state lives in shared
// memory, weights stream from L2 (2 MB per row and step).
//
// Mask recurrence exactly as the reference: step 0 has the end slot disabled; from step 1 on the end slot is enabled
// and every previously sampled slot is disabled; rows whose action type selects no units start with end_flag set.
#include <math_constants.h>
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kIn = 1024, kFunc = 256, kKey = 32, kG = 128;
constexpr float kEps = 1e-5f;

struct PtrWeights {
    const float *q1_w, *q1_b, *q2_w, *q2_b;                 // query_fc1 [256,1024], query_fc2 [32,256]
    const float *w_ih, *w_hh;                               // [128,32] each
    const float *lni_w, *lni_b, *lnh_w, *lnh_b, *lnc_w, *lnc_b;
    const float *e1_w, *e1_b, *e2_w, *e2_b;                 // embed_fc1 [256,32], embed_fc2 [1024,256]
};

struct PtrState {
    const float* emb0;        // [N,1024] embedding entering the head
    float* ae;                // [N,1024] current auto-regressive embedding (in/out)
    const float* key;         // [N,S,32]  (slot entity_num holds the learned end token)
    float* h; float* c;       // [N,32] LSTM state (in/out)
    uint8_t* mask;            // [N,S] selectable slots (in/out)
    float* ksum;              // [N,32] running sum of the selected keys (in/out)
    int* count;               // [N] number of selected units (in/out)
    uint8_t* end_flag;        // [N] (in/out)
    int64_t* num;             // [N] selected_units_num (in/out)
    const int64_t* entity_num;
    const int64_t* prev;      // [N] result of the previous step (ignored at step 0)
    const float* q;           // [N,S] Exp(1) draws for this step
    float* logits_out;        // [N,S] this step's logits (after mask and temperature)
    int64_t* result;          // [N]
    int* all_ended;           // [1] set to 0 by any row that has not ended after this step (caller presets 1)
    int N, S, step;
    float inv_temperature;
};

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = dsb::warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) t += red[i];
    return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
    v = dsb::warp_max(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float t = red[0];
#pragma unroll
    for (int i = 1; i < kWarps; ++i) t = fmaxf(t, red[i]);
    return t;
}

// out[o] = act(W[o,:] . x + b[o]) for o in [0, n_out): each warp owns outputs o = warp, warp + 8, ...; lanes stride over k
__device__ __forceinline__ void matvec(const float* __restrict__ W, const float* __restrict__ b, const float* x, float* out,
                                       int n_out, int n_in, bool relu) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = warp; o < n_out; o += kWarps) {
        const float* wr = W + (size_t)o * n_in;
        float acc = 0.f;
        for (int k = lane * 4; k < n_in; k += 128) {
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(wr + k));
            acc += w4.x * x[k] + w4.y * x[k + 1] + w4.z * x[k + 2] + w4.w * x[k + 3];
        }
        acc = dsb::warp_sum(acc);
        if (lane == 0) {
            float v = acc + (b ? b[o] : 0.f);
            out[o] = relu ? fmaxf(v, 0.f) : v;
        }
    }
}

// LayerNorm over n (<= 128) values held in shared memory, in place, by the first warp(s)
__device__ __forceinline__ void layernorm_small(float* v, const float* g, const float* b, int n, float* red) {
    float x = (threadIdx.x < n) ? v[threadIdx.x] : 0.f;
    const float mean = block_sum(x, red) / n;
    const float d = (threadIdx.x < n) ? x - mean : 0.f;
    const float var = block_sum(d * d, red) / n;
    const float rstd = rsqrtf(var + kEps);
    if (threadIdx.x < n) v[threadIdx.x] = d * rstd * g[threadIdx.x] + b[threadIdx.x];
    __syncthreads();
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(kThreads)
su_step_kernel(const PtrWeights w, const PtrState s) {
    extern __shared__ float sm[];
    float* ae = sm;                    // 1024
    float* x1 = ae + kIn;              // 256
    float* qv = x1 + kFunc;            // 32
    float* ig = qv + kKey;             // 128
    float* hg = ig + kG;               // 128
    float* hc = hg + kG;               // 64: h (32) | c (32)
    float* lg = hc + 64;               // S logits / probabilities
    float* red = lg + ((s.S + 3) & ~3);    // 2 * kWarps scratch
    __shared__ int s_idx[kWarps];
    __shared__ int s_result;
    const int n = blockIdx.x, tid = threadIdx.x;
    const int en = (int)s.entity_num[n];
    uint8_t* mask = s.mask + (size_t)n * s.S;

    for (int i = tid; i < kIn; i += kThreads) ae[i] = s.ae[(size_t)n * kIn + i];
    if (tid < 64) hc[tid] = tid < 32 ? s.h[n * 32 + tid] : s.c[n * 32 + tid - 32];
    // mask recurrence of the reference (applied at the top of step i > 0)
    if (s.step > 0 && tid == 0) {
        if (s.step == 1) mask[en] = 1;
        mask[s.prev[n]] = 0;
    }
    __syncthreads();

    // query = query_fc2(relu(query_fc1(ae)))
    matvec(w.q1_w, w.q1_b, ae, x1, kFunc, kIn, true);
    __syncthreads();
    matvec(w.q2_w, w.q2_b, x1, qv, kKey, kFunc, false);
    __syncthreads();
    // LayerNorm-LSTM cell (hidden 32): gates = LN_i(W_ih q) + LN_h(W_hh h)
    matvec(w.w_ih, nullptr, qv, ig, kG, kKey, false);
    matvec(w.w_hh, nullptr, hc, hg, kG, kKey, false);
    __syncthreads();
    layernorm_small(ig, w.lni_w, w.lni_b, kG, red);
    layernorm_small(hg, w.lnh_w, w.lnh_b, kG, red);
    if (tid < 32) {
        const float gi = ig[tid] + hg[tid], gf = ig[32 + tid] + hg[32 + tid], gg = ig[64 + tid] + hg[64 + tid];
        qv[tid] = sigmoidf_(gf) * hc[32 + tid] + sigmoidf_(gi) * tanhf(gg);      // pre-LN cell state (reuse qv)
    }
    __syncthreads();
    layernorm_small(qv, w.lnc_w, w.lnc_b, 32, red);
    if (tid < 32) {
        const float go = ig[96 + tid] + hg[96 + tid];
        const float cy = qv[tid];
        const float hy = sigmoidf_(go) * tanhf(cy);
        hc[tid] = hy;
        hc[32 + tid] = cy;
        s.h[n * 32 + tid] = hy;
        s.c[n * 32 + tid] = cy;
    }
    __syncthreads();
    // logits over the slots, mask, temperature
    const float* key = s.key + (size_t)n * s.S * kKey;
    float lmax = -CUDART_INF_F;
    for (int j = tid; j < s.S; j += kThreads) {
        const float4* kr = reinterpret_cast<const float4*>(key + (size_t)j * kKey);
        float acc = 0.f;
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            const float4 k4 = __ldg(kr + d);
            acc += k4.x * hc[4 * d] + k4.y * hc[4 * d + 1] + k4.z * hc[4 * d + 2] + k4.w * hc[4 * d + 3];
        }
        const float v = (mask[j] ? acc : -1e9f) * s.inv_temperature;
        lg[j] = v;
        s.logits_out[(size_t)n * s.S + j] = v;
        lmax = fmaxf(lmax, v);
    }
    lmax = block_max(lmax, red);
    float lsum = 0.f;
    for (int j = tid; j < s.S; j += kThreads) {
        const float e = expf(lg[j] - lmax);
        lg[j] = e;
        lsum += e;
    }
    lsum = block_sum(lsum, red);
    // multinomial(p, 1) == argmax(p / q), first maximum wins
    float best = -1.f;
    int bi = 0x7fffffff;
    for (int j = tid; j < s.S; j += kThreads) {
        const float v = (lg[j] / lsum) / s.q[(size_t)n * s.S + j];
        if (v > best) { best = v; bi = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    __syncthreads();
    if ((tid & 31) == 0) { red[tid >> 5] = best; s_idx[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
        float b = red[0];
        int r = s_idx[0];
        for (int i = 1; i < kWarps; ++i)
            if (red[i] > b || (red[i] == b && s_idx[i] < r)) { b = red[i]; r = s_idx[i]; }
        s_result = r;
        s.result[n] = r;
        // selected_units_num / end_flag bookkeeping (action_arg_head.py:283-288)
        const bool was_ended = s.end_flag[n] != 0;
        const bool is_end = (r == en);
        if (is_end && !was_ended) s.num[n] = s.step + 1;
        const bool ended = was_ended || is_end;
        s.end_flag[n] = ended ? 1 : 0;
        if (!ended) { s.count[n] += 1; atomicAnd(s.all_ended, 0); }
    }
    __syncthreads();
    const int r = s_result;
    const bool ended_now = s.end_flag[n] != 0;
    // masked mean of the selected keys -> embed MLP -> next auto-regressive embedding
    if (tid < 32) {
        float ks = s.ksum[n * 32 + tid];
        if (!ended_now) { ks += key[(size_t)r * kKey + tid]; s.ksum[n * 32 + tid] = ks; }
        const int cnt = s.count[n];
        qv[tid] = cnt != 0 ? ks / (float)cnt : ks;
    }
    __syncthreads();
    matvec(w.e1_w, w.e1_b, qv, x1, kFunc, kKey, true);
    __syncthreads();
    matvec(w.e2_w, w.e2_b, x1, ae, kIn, kFunc, false);
    __syncthreads();
    for (int i = tid; i < kIn; i += kThreads) s.ae[(size_t)n * kIn + i] = s.emb0[(size_t)n * kIn + i] + ae[i];
}

}  // namespace

extern "C" int dsb_su_sample_step(const void* const* weights16, const float* emb0, float* ae, const float* key, float* h,
                                  float* c, uint8_t* mask, float* ksum, int* count, uint8_t* end_flag, int64_t* num,
                                  const int64_t* entity_num, const int64_t* prev, const float* q, float* logits_out,
                                  int64_t* result, int* all_ended, int N, int S, int step, float temperature,
                                  dsb_stream_t stream) {
    DSB_REQUIRE(weights16 && emb0 && ae && key && h && c && mask && ksum && count && end_flag && num && entity_num && q &&
                logits_out && result && all_ended, "su_sample_step: null pointer");
    DSB_REQUIRE(step == 0 || prev, "su_sample_step: step > 0 needs the previous result");
    DSB_REQUIRE(N >= 0 && S > 0 && S <= 4096 && temperature > 0.f, "su_sample_step: bad shape");
    if (N == 0) return DSB_OK;
    PtrWeights w;
    const float** wp = reinterpret_cast<const float**>(&w);
    for (int i = 0; i < 16; ++i) { DSB_REQUIRE(weights16[i], "su_sample_step: null weight %d", i); wp[i] = (const float*)weights16[i]; }
    PtrState s;
    s.emb0 = emb0; s.ae = ae; s.key = key; s.h = h; s.c = c; s.mask = mask; s.ksum = ksum; s.count = count;
    s.end_flag = end_flag; s.num = num; s.entity_num = entity_num; s.prev = prev; s.q = q; s.logits_out = logits_out;
    s.result = result; s.all_ended = all_ended; s.N = N; s.S = S; s.step = step; s.inv_temperature = 1.0f / temperature;
    const size_t smem = (size_t)(kIn + kFunc + kKey + kG + kG + 64 + ((S + 3) & ~3) + 2 * kWarps) * sizeof(float);
    su_step_kernel<<<N, kThreads, smem, (cudaStream_t)stream>>>(w, s);
    return dsb::check_launch("su_sample_step");
}
