// Teacher-forced selected-units pointer network (training path of SURVEY K12): the three non-GEMM pieces of
// SelectedUnitsHead._train_query (head/action_arg_head.py:168-216), each one launch forward and one backward for ALL rows and
// steps, instead of the reference's Python loop over max(selected_units_num) steps (~25 launches per step) or round 1's
// one-hot / cummax / matmul glue over [P, S, 513] tensors.
//
//   su_prefix_mean   mean (or sum) of the keys of the units selected up to step i  -> the input of embed_fc1        (:196-199)
//   su_lstm          the 32-wide LayerNorm-LSTM over the S steps of a row, state in registers (one warp per row)    (:187-189)
//   su_logits        h_i . key_e for every (step, slot) with the reference's mask recurrence folded in             (:190-195)
//
// Keys are read IN PLACE from the stacked key projection of the two pointer heads ([P, E, ld] fp32, this head's 32 columns
// first); the learned end token (slot entity_num, :118-129) is substituted on the fly, so neither the [P, 513, 32] key tensor
// nor its gradient copy exist.  Key gradients are ADDED into a caller-zeroed [P, E, ld] buffer shared with the target-unit head.
#include <math_constants.h>
#include "common.cuh"

namespace {

constexpr int kKey = 32;
constexpr int kMaxS = 64;                   // MAX_SELECTED_UNITS_NUM
constexpr float kEps = 1e-5f;

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + __expf(-x)); }

// ---------------------------------------------------------------------------------------------- prefix mean of selected keys
// one warp per row p, lane = key component.  included(j): step j adds a NEW unit (not after / at the end token, not a repeat).
__global__ void su_prefix_mean_fwd_kernel(const float* __restrict__ kfull, int ld, const int64_t* __restrict__ su, int su_ld,
                                          const int64_t* __restrict__ entity_num, const int64_t* __restrict__ num,
                                          float* __restrict__ mean, int* __restrict__ cnt_out, int64_t P, int E, int S) {
    const int lane = threadIdx.x & 31;
    const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (p >= P) return;
    const int en = (int)entity_num[p];
    const bool normalise = num[p] != 0;
    const int l0 = lane < S ? (int)su[p * su_ld + lane] : -1, l1 = lane + 32 < S ? (int)su[p * su_ld + lane + 32] : -1;
    float ksum = 0.f;
    int cnt = 0;
    bool ended = false;
    unsigned inc0 = 0, inc1 = 0;                           // bit j: step j (resp. 32 + j) was included
    for (int j = 0; j < S; ++j) {
        const int lab = __shfl_sync(0xffffffffu, j < 32 ? l0 : l1, j & 31);
        ended = ended || lab == en;
        // a repeat of an already included unit does not change the (set-valued) selection (cummax of one-hots in the reference)
        const unsigned same0 = __ballot_sync(0xffffffffu, l0 == lab) & inc0, same1 = __ballot_sync(0xffffffffu, l1 == lab) & inc1;
        const bool add = !ended && !(same0 | same1) && lab >= 0 && lab < E;
        if (add) {
            ksum += kfull[((int64_t)p * E + lab) * ld + lane];
            ++cnt;
            if (j < 32) inc0 |= 1u << j; else inc1 |= 1u << (j - 32);
        }
        mean[((int64_t)p * S + j) * kKey + lane] = normalise ? ksum / (float)cnt : ksum;      // 0/0 = NaN as in the reference
        if (lane == 0) cnt_out[p * S + j] = cnt;
    }
}
__global__ void su_prefix_mean_bwd_kernel(const float* __restrict__ g, const int64_t* __restrict__ su, int su_ld,
                                          const int64_t* __restrict__ entity_num, const int64_t* __restrict__ num,
                                          const int* __restrict__ cnt_in, float* __restrict__ dkfull, int ld, int64_t P, int E,
                                          int S) {
    const int lane = threadIdx.x & 31;
    const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (p >= P) return;
    const bool normalise = num[p] != 0;
    float acc = 0.f;
    for (int i = S - 1; i >= 0; --i) {
        const int cnt = cnt_in[p * S + i];
        const float gi = g[((int64_t)p * S + i) * kKey + lane];
        acc += normalise ? (cnt > 0 ? gi / (float)cnt : 0.f) : gi;
        const int before = i > 0 ? cnt_in[p * S + i - 1] : 0;
        if (cnt > before) {                                 // step i added a new unit: it is in every mean from i on
            const int lab = (int)su[p * su_ld + i];
            dkfull[((int64_t)p * E + lab) * ld + lane] += acc;
        }
    }
}

// ---------------------------------------------------------------------------------------------- 32-wide LN-LSTM over S steps
// one warp per row; lane u owns hidden unit u and the gate columns u, 32+u, 64+u, 96+u.  W_hh [128, 32] sits in shared memory
// (row pitch 33: conflict free both by row and by column).
struct SuLstmSaved { float* gates; float* hg; float* pre_c; float* st; };   // [P,S,128], [P,S,128], [P,S,32], [P,S,4]

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
su_lstm_fwd_kernel(const float* __restrict__ ig, const float* __restrict__ w_hh, const float* __restrict__ gam_h,
                   const float* __restrict__ bet_h, const float* __restrict__ gam_c, const float* __restrict__ bet_c,
                   float* __restrict__ hs, float* __restrict__ cs, SuLstmSaved sv, int64_t P, int S) {
    __shared__ float W[128 * 33];
    for (int i = threadIdx.x; i < 128 * 32; i += WARPS * 32) W[(i >> 5) * 33 + (i & 31)] = w_hh[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (p >= P) return;
    float gh[4], bh[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) { gh[g] = gam_h[32 * g + lane]; bh[g] = bet_h[32 * g + lane]; }
    const float gc = gam_c[lane], bc = bet_c[lane];
    float h = 0.f, c = 0.f;
    for (int i = 0; i < S; ++i) {
        float hg[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const float hk = __shfl_sync(0xffffffffu, h, k);
#pragma unroll
            for (int g = 0; g < 4; ++g) hg[g] = fmaf(hk, W[(32 * g + lane) * 33 + k], hg[g]);
        }
        const float mean = dsb::warp_sum((hg[0] + hg[1]) + (hg[2] + hg[3])) * (1.0f / 128);
        float q = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) { const float d = hg[g] - mean; q += d * d; }
        const float rstd = rsqrtf(dsb::warp_sum(q) * (1.0f / 128) + kEps);
        const int64_t row = p * S + i;
        float gate[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            gate[g] = ig[row * 128 + 32 * g + lane] + ((hg[g] - mean) * rstd * gh[g] + bh[g]);
            sv.gates[row * 128 + 32 * g + lane] = gate[g];
            sv.hg[row * 128 + 32 * g + lane] = hg[g];
        }
        const float pc = sigmoidf(gate[1]) * c + sigmoidf(gate[0]) * tanhf(gate[2]);
        const float mean_c = dsb::warp_sum(pc) * (1.0f / 32);
        const float dc = pc - mean_c;
        const float rstd_c = rsqrtf(dsb::warp_sum(dc * dc) * (1.0f / 32) + kEps);
        c = dc * rstd_c * gc + bc;
        h = sigmoidf(gate[3]) * tanhf(c);
        hs[row * 32 + lane] = h;
        cs[row * 32 + lane] = c;
        sv.pre_c[row * 32 + lane] = pc;
        if (lane == 0) { sv.st[row * 4] = mean; sv.st[row * 4 + 1] = rstd; sv.st[row * 4 + 2] = mean_c; sv.st[row * 4 + 3] = rstd_c; }
    }
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
su_lstm_bwd_kernel(const float* __restrict__ g_hs, const float* __restrict__ w_hh, const float* __restrict__ gam_h,
                   const float* __restrict__ gam_c, const float* __restrict__ bet_c, const float* __restrict__ cs,
                   SuLstmSaved sv, float* __restrict__ d_ig, float* __restrict__ d_hg, float* __restrict__ dgam_h,
                   float* __restrict__ dbet_h, float* __restrict__ dgam_c, float* __restrict__ dbet_c, int64_t P, int S) {
    __shared__ float W[128 * 33];
    __shared__ float acc_s[WARPS][10][32];
    for (int i = threadIdx.x; i < 128 * 32; i += WARPS * 32) W[(i >> 5) * 33 + (i & 31)] = w_hh[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t p = (int64_t)blockIdx.x * WARPS + warp;
    float a_gh[4] = {0.f, 0.f, 0.f, 0.f}, a_bh[4] = {0.f, 0.f, 0.f, 0.f}, a_gc = 0.f, a_bc = 0.f;
    if (p < P) {
        float gmh[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) gmh[g] = gam_h[32 * g + lane];
        const float gc = gam_c[lane], bc = bet_c[lane];
        float dh_next = 0.f, dc_next = 0.f;
        for (int i = S - 1; i >= 0; --i) {
            const int64_t row = p * S + i;
            const float mean_h = sv.st[row * 4], rstd_h = sv.st[row * 4 + 1], mean_c = sv.st[row * 4 + 2], rstd_c = sv.st[row * 4 + 3];
            const float gh = g_hs[row * 32 + lane] + dh_next;
            float gate[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) gate[g] = sv.gates[row * 128 + 32 * g + lane];
            const float xs = (sv.pre_c[row * 32 + lane] - mean_c) * rstd_c;
            const float cy = xs * gc + bc;
            const float th = tanhf(cy), so = sigmoidf(gate[3]);
            const float dgo = gh * th * so * (1.f - so);
            const float d = gh * so * (1.f - th * th) + dc_next;
            a_gc += d * xs;
            a_bc += d;
            const float dcy = d * gc;
            const float s1 = dsb::warp_sum(dcy) * (1.0f / 32), s2 = dsb::warp_sum(dcy * xs) * (1.0f / 32);
            const float dpc = rstd_c * (dcy - s1 - xs * s2);
            const float cin = i > 0 ? cs[(row - 1) * 32 + lane] : 0.f;
            const float si = sigmoidf(gate[0]), sf = sigmoidf(gate[1]), tg = tanhf(gate[2]);
            float dg[4] = {dpc * tg * si * (1.f - si), dpc * cin * sf * (1.f - sf), dpc * si * (1.f - tg * tg), dgo};
            dc_next = dpc * sf;
            float xh[4], t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                d_ig[row * 128 + 32 * g + lane] = dg[g];
                xh[g] = (sv.hg[row * 128 + 32 * g + lane] - mean_h) * rstd_h;
                a_gh[g] += dg[g] * xh[g];
                a_bh[g] += dg[g];
                dg[g] *= gmh[g];
                t1 += dg[g];
                t2 += dg[g] * xh[g];
            }
            t1 = dsb::warp_sum(t1) * (1.0f / 128);
            t2 = dsb::warp_sum(t2) * (1.0f / 128);
            float dhg[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                dhg[g] = rstd_h * (dg[g] - t1 - xh[g] * t2);
                d_hg[row * 128 + 32 * g + lane] = dhg[g];
            }
            // dh[i-1][k = lane] = sum_j d_hg[j] W_hh[j][k]
            float acc = 0.f;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
#pragma unroll
                for (int l = 0; l < 32; ++l) acc = fmaf(__shfl_sync(0xffffffffu, dhg[g], l), W[(32 * g + l) * 33 + lane], acc);
            }
            dh_next = acc;
        }
    }
    // LayerNorm parameter gradients: fold the block's warps, then one atomic per parameter element and block
#pragma unroll
    for (int g = 0; g < 4; ++g) { acc_s[warp][g][lane] = a_gh[g]; acc_s[warp][4 + g][lane] = a_bh[g]; }
    acc_s[warp][8][lane] = a_gc;
    acc_s[warp][9][lane] = a_bc;
    __syncthreads();
    for (int i = threadIdx.x; i < 10 * 32; i += WARPS * 32) {
        const int which = i >> 5, l = i & 31;
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) t += acc_s[w][which][l];
        if (which < 4) atomicAdd(dgam_h + 32 * which + l, t);
        else if (which < 8) atomicAdd(dbet_h + 32 * (which - 4) + l, t);
        else if (which == 8) atomicAdd(dgam_c + l, t);
        else atomicAdd(dbet_c + l, t);
    }
}

// ---------------------------------------------------------------------------------------------- logits with the mask recurrence
// one CTA per row p; thread e strides over the E + 1 slots with the slot's key in registers.
constexpr int kLogitThreads = 256;

__device__ __forceinline__ void load_key(float (&k)[kKey], const float* __restrict__ kfull, int ld, const float* __restrict__ end_emb,
                                         int64_t p, int E, int e, int en) {
    const float* src = e == en ? end_emb : (e < E ? kfull + ((int64_t)p * E + e) * ld : nullptr);
#pragma unroll
    for (int d = 0; d < kKey; d += 4) {
        const float4 v = src ? *reinterpret_cast<const float4*>(src + d) : make_float4(0.f, 0.f, 0.f, 0.f);
        k[d] = v.x; k[d + 1] = v.y; k[d + 2] = v.z; k[d + 3] = v.w;
    }
}

// step mask of the reference (action_arg_head.py:179-186): slot valid (<= entity_num), not chosen at an earlier step, and the end
// token not selectable at step 0
__device__ __forceinline__ bool slot_open(int i, int e, int en, int first_chosen) {
    return e <= en && !(first_chosen < i) && !(i == 0 && e == en);
}

__global__ void __launch_bounds__(kLogitThreads)
su_logits_fwd_kernel(const float* __restrict__ hs, const float* __restrict__ kfull, int ld, const float* __restrict__ end_emb,
                     const int64_t* __restrict__ su, int su_ld, const int64_t* __restrict__ entity_num,
                     float* __restrict__ logits, int E, int S) {
    extern __shared__ int sm_i[];
    int* first = sm_i;                                           // [E + 1] first step at which the slot was chosen
    float* h_s = reinterpret_cast<float*>(sm_i + ((E + 1 + 3) & ~3));   // [S, 32]
    const int64_t p = blockIdx.x;
    const int en = (int)entity_num[p];
    for (int e = threadIdx.x; e <= E; e += kLogitThreads) first[e] = 0x7fffffff;
    for (int i = threadIdx.x; i < S * kKey; i += kLogitThreads) h_s[i] = hs[p * S * kKey + i];
    __syncthreads();
    if (threadIdx.x < S) {
        const int64_t lab = su[p * su_ld + threadIdx.x];
        if (lab >= 0 && lab <= E) atomicMin(&first[lab], (int)threadIdx.x);
    }
    __syncthreads();
    for (int e = threadIdx.x; e <= E; e += kLogitThreads) {
        float k[kKey];
        load_key(k, kfull, ld, end_emb, p, E, e, en);
        const int fc = first[e];
        for (int i = 0; i < S; ++i) {
            float acc = 0.f;
#pragma unroll
            for (int d = 0; d < kKey; ++d) acc = fmaf(h_s[i * kKey + d], k[d], acc);
            logits[((int64_t)p * S + i) * (E + 1) + e] = slot_open(i, e, en, fc) ? acc : -1e9f;
        }
    }
}

// dh[i, :] = sum_e g[i, e] key_e and dkey[p, e, :] += sum_i g[i, e] h[i, :] (slot entity_num -> the end embedding's gradient),
// g = grad_logits at open slots, 0 at masked ones.  Both are tiny per-row matrix products; threads are laid out over the OUTPUT
// (8 steps x 32 components, then 8 slots x 32 components), so neither needs a cross-thread reduction.
__global__ void __launch_bounds__(kLogitThreads)
su_logits_bwd_kernel(const float* __restrict__ gl, const float* __restrict__ hs, const float* __restrict__ kfull, int ld,
                     const float* __restrict__ end_emb, const int64_t* __restrict__ su, int su_ld,
                     const int64_t* __restrict__ entity_num, float* __restrict__ dhs, float* __restrict__ dkfull,
                     float* __restrict__ d_end_emb, int E, int S) {
    extern __shared__ int sm_i[];
    int* first = sm_i;
    float* h_s = reinterpret_cast<float*>(sm_i + ((E + 1 + 3) & ~3));   // [S, 32]
    const int64_t p = blockIdx.x;
    const int en = (int)entity_num[p];
    const int d = threadIdx.x & 31, grp = threadIdx.x >> 5;
    for (int e = threadIdx.x; e <= E; e += kLogitThreads) first[e] = 0x7fffffff;
    for (int i = threadIdx.x; i < S * kKey; i += kLogitThreads) h_s[i] = hs[p * S * kKey + i];
    __syncthreads();
    if (threadIdx.x < S) {
        const int64_t lab = su[p * su_ld + threadIdx.x];
        if (lab >= 0 && lab <= E) atomicMin(&first[lab], (int)threadIdx.x);
    }
    __syncthreads();
    const float* g_row = gl + (int64_t)p * S * (E + 1);
    const float end_d = end_emb[d];
    // ---- dh
    for (int i = grp; i < S; i += kLogitThreads / 32) {
        float acc = 0.f;
        const int last = en < E ? en : E;                       // slots above entity_num are never open
        for (int e = 0; e <= last; ++e) {
            if (!slot_open(i, e, en, first[e])) continue;
            const float kv = e == en ? end_d : kfull[((int64_t)p * E + e) * ld + d];
            acc = fmaf(g_row[(int64_t)i * (E + 1) + e], kv, acc);
        }
        dhs[((int64_t)p * S + i) * kKey + d] = acc;
    }
    // ---- dkey (+ the end embedding's gradient)
    const int last = en < E ? en : E;
    for (int e = grp; e <= last; e += kLogitThreads / 32) {
        const int fc = first[e];
        float acc = 0.f;
        for (int i = 0; i < S; ++i)
            if (slot_open(i, e, en, fc)) acc = fmaf(g_row[(int64_t)i * (E + 1) + e], h_s[i * kKey + d], acc);
        if (e == en) atomicAdd(d_end_emb + d, acc);
        else dkfull[((int64_t)p * E + e) * ld + d] += acc;
    }
}

constexpr int kLstmWarps = 8;

}  // namespace

extern "C" int dsb_su_prefix_mean_fwd(const float* kfull, int ld, const int64_t* selected_units, int su_ld,
                                      const int64_t* entity_num, const int64_t* selected_units_num, float* mean, int* count,
                                      int64_t P, int E, int S, dsb_stream_t stream) {
    DSB_REQUIRE(kfull && selected_units && entity_num && selected_units_num && mean && count && P >= 0 && E > 0 && S > 0 &&
                S <= kMaxS && ld >= kKey && su_ld >= S, "su_prefix_mean_fwd: bad argument");
    if (P == 0) return DSB_OK;
    su_prefix_mean_fwd_kernel<<<(unsigned)((P * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        kfull, ld, selected_units, su_ld, entity_num, selected_units_num, mean, count, P, E, S);
    return dsb::check_launch("su_prefix_mean_fwd");
}
extern "C" int dsb_su_prefix_mean_bwd(const float* grad_mean, const int64_t* selected_units, int su_ld, const int64_t* entity_num,
                                      const int64_t* selected_units_num, const int* count, float* grad_kfull, int ld, int64_t P,
                                      int E, int S, dsb_stream_t stream) {
    DSB_REQUIRE(grad_mean && selected_units && entity_num && selected_units_num && count && grad_kfull && P >= 0 && E > 0 &&
                S > 0 && S <= kMaxS && ld >= kKey && su_ld >= S, "su_prefix_mean_bwd: bad argument");
    if (P == 0) return DSB_OK;
    su_prefix_mean_bwd_kernel<<<(unsigned)((P * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        grad_mean, selected_units, su_ld, entity_num, selected_units_num, count, grad_kfull, ld, P, E, S);
    return dsb::check_launch("su_prefix_mean_bwd");
}

extern "C" int dsb_su_lstm_fwd(const float* ig, const float* w_hh, const float* gamma_h, const float* beta_h,
                               const float* gamma_c, const float* beta_c, float* hs, float* cs, float* gates, float* hg,
                               float* pre_c, float* stats, int64_t P, int S, dsb_stream_t stream) {
    DSB_REQUIRE(ig && w_hh && gamma_h && beta_h && gamma_c && beta_c && hs && cs && gates && hg && pre_c && stats && P >= 0 &&
                S > 0, "su_lstm_fwd: bad argument");
    if (P == 0) return DSB_OK;
    SuLstmSaved sv{gates, hg, pre_c, stats};
    su_lstm_fwd_kernel<kLstmWarps><<<(unsigned)((P + kLstmWarps - 1) / kLstmWarps), kLstmWarps * 32, 0, (cudaStream_t)stream>>>(
        ig, w_hh, gamma_h, beta_h, gamma_c, beta_c, hs, cs, sv, P, S);
    return dsb::check_launch("su_lstm_fwd");
}
extern "C" int dsb_su_lstm_bwd(const float* grad_hs, const float* w_hh, const float* gamma_h, const float* gamma_c,
                               const float* beta_c, const float* cs, const float* gates, const float* hg, const float* pre_c,
                               const float* stats, float* d_ig, float* d_hg, float* dgamma_h, float* dbeta_h, float* dgamma_c,
                               float* dbeta_c, int64_t P, int S, dsb_stream_t stream) {
    DSB_REQUIRE(grad_hs && w_hh && gamma_h && gamma_c && beta_c && cs && gates && hg && pre_c && stats && d_ig && d_hg &&
                dgamma_h && dbeta_h && dgamma_c && dbeta_c && P >= 0 && S > 0, "su_lstm_bwd: bad argument");
    if (P == 0) return DSB_OK;
    SuLstmSaved sv{const_cast<float*>(gates), const_cast<float*>(hg), const_cast<float*>(pre_c), const_cast<float*>(stats)};
    su_lstm_bwd_kernel<kLstmWarps><<<(unsigned)((P + kLstmWarps - 1) / kLstmWarps), kLstmWarps * 32, 0, (cudaStream_t)stream>>>(
        grad_hs, w_hh, gamma_h, gamma_c, beta_c, cs, sv, d_ig, d_hg, dgamma_h, dbeta_h, dgamma_c, dbeta_c, P, S);
    return dsb::check_launch("su_lstm_bwd");
}

extern "C" int dsb_su_logits_fwd(const float* hs, const float* kfull, int ld, const float* end_embedding,
                                 const int64_t* selected_units, int su_ld, const int64_t* entity_num, float* logits, int64_t P,
                                 int E, int S, dsb_stream_t stream) {
    DSB_REQUIRE(hs && kfull && end_embedding && selected_units && entity_num && logits && P >= 0 && E > 0 && E + 1 <= 3 * 256 &&
                S > 0 && S <= kMaxS && ld >= kKey && ld % 4 == 0 && su_ld >= S, "su_logits_fwd: bad argument");
    if (P == 0) return DSB_OK;
    const size_t smem = (size_t)(((E + 1 + 3) & ~3) + S * kKey) * 4;
    su_logits_fwd_kernel<<<(unsigned)P, kLogitThreads, smem, (cudaStream_t)stream>>>(hs, kfull, ld, end_embedding, selected_units,
                                                                                   su_ld, entity_num, logits, E, S);
    return dsb::check_launch("su_logits_fwd");
}
extern "C" int dsb_su_logits_bwd(const float* grad_logits, const float* hs, const float* kfull, int ld,
                                 const float* end_embedding, const int64_t* selected_units, int su_ld,
                                 const int64_t* entity_num, float* grad_hs, float* grad_kfull, float* grad_end_embedding,
                                 int64_t P, int E, int S, dsb_stream_t stream) {
    DSB_REQUIRE(grad_logits && hs && kfull && end_embedding && selected_units && entity_num && grad_hs && grad_kfull &&
                grad_end_embedding && P >= 0 && E > 0 && E + 1 <= 3 * 256 && S > 0 && S <= kMaxS && ld >= kKey && ld % 4 == 0 &&
                su_ld >= S, "su_logits_bwd: bad argument");
    if (P == 0) return DSB_OK;
    const size_t smem = (size_t)(((E + 1 + 3) & ~3) + S * kKey) * 4;
    su_logits_bwd_kernel<<<(unsigned)P, kLogitThreads, smem, (cudaStream_t)stream>>>(
        grad_logits, hs, kfull, ld, end_embedding, selected_units, su_ld, entity_num, grad_hs, grad_kfull, grad_end_embedding, E, S);
    return dsb::check_launch("su_logits_bwd");
}
