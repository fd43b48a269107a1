// Per-row categorical statistics (forward + backward) and masked categorical sampling.
//
// Forward replaces, per head, the reference's Categorical(logits).probs/.logits/.log_prob + entropy_loss + kl_loss
// passes (rl_loss.py:63-90, as_rl_utils.py:52-103): >= 10 full passes over [rows, C] tensors become ONE read of
// the target logits and ONE read of the teacher logits; nothing of size [rows, C] is written.
// Backward is the closed form of SURVEY.md Appendix C and re-reads the same two tensors once.
// Rows are independent: C <= 1024 -> one warp per row; larger C (16384 location logits) -> one CTA per row.
#include <math_constants.h>
#include "common.cuh"

namespace {

struct RowStats { float m, s; };  // running max and sum(exp(x - m))

__device__ __forceinline__ void online_update(RowStats& a, float x) {
    if (x > a.m) { a.s = a.s * __expf(a.m - x) + 1.f; a.m = x; }      // a.m = -inf, a.s = 0 on the first element: 0 * 0 + 1
    else if (x > -CUDART_INF_F) a.s += __expf(x - a.m);               // a logit of -inf contributes nothing (never -inf - -inf)
}
__device__ __forceinline__ RowStats merge(RowStats a, RowStats b) {
    if (b.m > a.m) { RowStats t = a; a = b; b = t; }
    if (b.m > -CUDART_INF_F) a.s += b.s * __expf(b.m - a.m);
    return a;
}

// Block-wide (or warp-wide when blockDim.x == 32 per row) reductions over `nthreads` cooperating threads.
template <int WARPS>
struct Coop {
    float* red;  // [WARPS * 2] scratch in shared memory (unused when WARPS == 1)
    __device__ __forceinline__ float sum(float v) {
        v = dsb::warp_sum(v);
        if (WARPS == 1) return v;
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        __syncthreads();
        if (lane == 0) red[w] = v;
        __syncthreads();
        float t = 0.f;
        for (int i = 0; i < WARPS; ++i) t += red[i];
        return t;
    }
    __device__ __forceinline__ RowStats stats(RowStats a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            RowStats b;
            b.m = __shfl_xor_sync(0xffffffffu, a.m, o);
            b.s = __shfl_xor_sync(0xffffffffu, a.s, o);
            a = merge(a, b);
        }
        if (WARPS == 1) return a;
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        __syncthreads();
        if (lane == 0) { red[2 * w] = a.m; red[2 * w + 1] = a.s; }
        __syncthreads();
        RowStats t{red[0], red[1]};
        for (int i = 1; i < WARPS; ++i) t = merge(t, RowStats{red[2 * i], red[2 * i + 1]});
        return t;
    }
};

// WARPS == 1: blockDim = (32, rows_per_block); WARPS > 1: blockDim = (32*WARPS, 1), one row per block.
template <int WARPS>
__global__ void stats_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ teacher,
                                 const int64_t* __restrict__ action, float* __restrict__ lse_out,
                                 float* __restrict__ logp_out, float* __restrict__ ent_out,
                                 float* __restrict__ kl_out, float* __restrict__ lse_t_out,
                                 float* __restrict__ mean_lp_out, int* __restrict__ error_flag, int64_t rows, int C) {
    __shared__ float red[WARPS * 2 + 2];
    Coop<WARPS> coop{red};
    const int64_t r = (WARPS == 1) ? (int64_t)blockIdx.x * blockDim.y + threadIdx.y : blockIdx.x;
    if (r >= rows) return;   // whole warp / block exits together
    const int nth = 32 * WARPS;
    const int t = threadIdx.x;
    const float* z = logits + r * C;
    const float* tz = teacher ? teacher + r * C : nullptr;
    RowStats a{-CUDART_INF_F, 0.f}, b{-CUDART_INF_F, 0.f};
    for (int j = t; j < C; j += nth) {
        online_update(a, z[j]);
        if (tz) online_update(b, tz[j]);
    }
    a = coop.stats(a);
    // log-softmax as (z - max) - log(sum): a row that is entirely -1e9 (padded selected-units step) then gives
    // -log(C) like the reference instead of losing log(sum) in the rounding of max + log(sum).
    const float zm = a.m, zl = logf(a.s);
    float tm = 0.f, tl = 0.f;
    if (tz) { b = coop.stats(b); tm = b.m; tl = logf(b.s); }
    // second pass (row is L1/L2 resident): entropy and KL
    float ent = 0.f, kl = 0.f, slp = 0.f;
    for (int j = t; j < C; j += nth) {
        const float l = (z[j] - zm) - zl;
        const float p = __expf(l);
        ent -= p * l;
        slp += l;
        if (tz) {
            const float lt = (tz[j] - tm) - tl;
            kl += __expf(lt) * (lt - l);
        }
    }
    ent = coop.sum(ent);
    if (tz) kl = coop.sum(kl);
    if (mean_lp_out) slp = coop.sum(slp);
    if (t == 0) {
        lse_out[2 * r] = zm;
        lse_out[2 * r + 1] = zl;
        if (logp_out) {
            // torch's Categorical.log_prob / cross_entropy fail on a label outside [0, C); here the label is clamped (no
            // out-of-bounds read) and the error recorded for the host to raise
            int64_t a = action[r];
            if (a < 0 || a >= C) { if (error_flag) atomicOr(error_flag, 2); a = a < 0 ? 0 : C - 1; }
            logp_out[r] = (z[a] - zm) - zl;
        }
        if (mean_lp_out) mean_lp_out[r] = slp / (float)C;
        if (ent_out) ent_out[r] = ent;
        if (tz) { kl_out[r] = kl; lse_t_out[2 * r] = tm; lse_t_out[2 * r + 1] = tl; }
    }
}

template <int WARPS>
__global__ void stats_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ teacher,
                                 const int64_t* __restrict__ action, const float* __restrict__ lse_in,
                                 const float* __restrict__ ent_in, const float* __restrict__ lse_t_in,
                                 const float* __restrict__ g_logp, const float* __restrict__ g_ent,
                                 const float* __restrict__ g_kl, const float* __restrict__ g_mean,
                                 float* __restrict__ grad, int64_t rows, int C) {
    const int64_t r = (WARPS == 1) ? (int64_t)blockIdx.x * blockDim.y + threadIdx.y : blockIdx.x;
    if (r >= rows) return;
    const int nth = 32 * WARPS;
    const float* z = logits + r * C;
    const float* tz = teacher ? teacher + r * C : nullptr;
    float* g = grad + r * C;
    const float gm = g_mean ? g_mean[r] : 0.f;           // d mean_j(log p_j) / dz_j = 1/C - p_j
    const float gm_c = gm / (float)C;
    const float zm = lse_in[2 * r], zl = lse_in[2 * r + 1];
    const float gl = g_logp ? g_logp[r] : 0.f;
    const float ge = g_ent ? g_ent[r] : 0.f;
    const float gk = (g_kl && tz) ? g_kl[r] : 0.f;
    const float H = ge != 0.f ? ent_in[r] : 0.f;
    const float tm = tz ? lse_t_in[2 * r] : 0.f, tl = tz ? lse_t_in[2 * r + 1] : 0.f;
    int64_t a64 = action[r];
    const int a = a64 < 0 ? 0 : (a64 >= C ? C - 1 : (int)a64);      // clamped as in the forward (which flagged it)
    for (int j = threadIdx.x; j < C; j += nth) {
        const float l = (z[j] - zm) - zl;
        const float p = __expf(l);
        float v = gl * ((j == a ? 1.f : 0.f) - p) + (gm_c - gm * p);
        if (ge != 0.f) v -= ge * p * (l + H);
        if (gk != 0.f) v += gk * (p - __expf((tz[j] - tm) - tl));
        g[j] = v;
    }
}

// index = argmax_j softmax(z)_j / q_j, first maximum wins (torch.multinomial n=1 == argmax(p / q), q ~ Exp(1)).
template <int WARPS>
__global__ void sample_kernel(const float* __restrict__ logits, const float* __restrict__ q,
                              int64_t* __restrict__ index, float* __restrict__ logp, int64_t rows, int C) {
    __shared__ float red[WARPS * 2 + 2];
    __shared__ int redi[WARPS + 1];
    Coop<WARPS> coop{red};
    const int64_t r = (WARPS == 1) ? (int64_t)blockIdx.x * blockDim.y + threadIdx.y : blockIdx.x;
    if (r >= rows) return;
    const int nth = 32 * WARPS;
    const int t = threadIdx.x;
    const float* z = logits + r * C;
    const float* qr = q + r * C;
    RowStats a{-CUDART_INF_F, 0.f};
    for (int j = t; j < C; j += nth) online_update(a, z[j]);
    a = coop.stats(a);
    // ATen softmax: exp(x - max) / sum ; multinomial: (p / q).argmax
    const float inv = 1.0f;  // the division by the row sum is a positive per-row constant: argmax unchanged,
    (void)inv;               // but near-ties could round differently, so apply it exactly as ATen does:
    float best = -1.f;
    int bi = 0x7fffffff;
    for (int j = t; j < C; j += nth) {
        const float p = expf(z[j] - a.m) / a.s;
        const float v = p / qr[j];
        if (v > best) { best = v; bi = j; }     // strided ascending j per thread: keeps the first max
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (WARPS > 1) {
        const int lane = t & 31, w = t >> 5;
        __syncthreads();
        if (lane == 0) { red[w] = best; redi[w] = bi; }
        __syncthreads();
        best = red[0]; bi = redi[0];
        for (int i = 1; i < WARPS; ++i)
            if (red[i] > best || (red[i] == best && redi[i] < bi)) { best = red[i]; bi = redi[i]; }
    }
    if (t == 0) {
        index[r] = bi;
        if (logp) logp[r] = (z[bi] - a.m) - logf(a.s);
    }
}

constexpr int kBigWarps = 8;
constexpr int kRowsPerBlock = 8;

}  // namespace

extern "C" int dsb_categorical_stats_fwd(const float* logits, const float* teacher, const int64_t* action,
                                         float* lse, float* logp, float* entropy, float* kl, float* lse_t,
                                         float* mean_logp, int* error_flag, int64_t rows, int C, dsb_stream_t stream) {
    DSB_REQUIRE(logits && lse && C > 0 && rows >= 0, "categorical_stats_fwd: bad argument");
    DSB_REQUIRE(!logp || action, "categorical_stats_fwd: logp needs action");
    DSB_REQUIRE(!teacher || (kl && lse_t), "categorical_stats_fwd: teacher needs kl and lse_t outputs");
    if (rows == 0) return DSB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (C <= 1024) {
        dim3 block(32, kRowsPerBlock);
        stats_fwd_kernel<1><<<(unsigned)((rows + kRowsPerBlock - 1) / kRowsPerBlock), block, 0, s>>>(
            logits, teacher, action, lse, logp, entropy, kl, lse_t, mean_logp, error_flag, rows, C);
    } else {
        stats_fwd_kernel<kBigWarps><<<(unsigned)rows, 32 * kBigWarps, 0, s>>>(logits, teacher, action, lse, logp, entropy,
                                                                              kl, lse_t, mean_logp, error_flag, rows, C);
    }
    return dsb::check_launch("categorical_stats_fwd");
}

extern "C" int dsb_categorical_stats_bwd(const float* logits, const float* teacher, const int64_t* action,
                                         const float* lse, const float* entropy, const float* lse_t,
                                         const float* g_logp, const float* g_ent, const float* g_kl,
                                         const float* g_mean, float* grad_logits, int64_t rows, int C,
                                         dsb_stream_t stream) {
    DSB_REQUIRE(logits && action && lse && grad_logits && C > 0, "categorical_stats_bwd: bad argument");
    DSB_REQUIRE(!g_ent || entropy, "categorical_stats_bwd: g_ent needs the forward entropy");
    DSB_REQUIRE(!(g_kl && teacher) || lse_t, "categorical_stats_bwd: g_kl needs lse_t");
    if (rows == 0) return DSB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (C <= 1024) {
        dim3 block(32, kRowsPerBlock);
        stats_bwd_kernel<1><<<(unsigned)((rows + kRowsPerBlock - 1) / kRowsPerBlock), block, 0, s>>>(
            logits, teacher, action, lse, entropy, lse_t, g_logp, g_ent, g_kl, g_mean, grad_logits, rows, C);
    } else {
        stats_bwd_kernel<kBigWarps><<<(unsigned)rows, 32 * kBigWarps, 0, s>>>(
            logits, teacher, action, lse, entropy, lse_t, g_logp, g_ent, g_kl, g_mean, grad_logits, rows, C);
    }
    return dsb::check_launch("categorical_stats_bwd");
}

extern "C" int dsb_sample_categorical(const float* logits, const float* q, int64_t* index, float* logp, int64_t rows,
                                      int C, dsb_stream_t stream) {
    DSB_REQUIRE(logits && q && index && C > 0, "sample_categorical: bad argument");
    if (rows == 0) return DSB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (C <= 1024) {
        dim3 block(32, kRowsPerBlock);
        sample_kernel<1><<<(unsigned)((rows + kRowsPerBlock - 1) / kRowsPerBlock), block, 0, s>>>(logits, q, index,
                                                                                                 logp, rows, C);
    } else {
        sample_kernel<kBigWarps><<<(unsigned)rows, 32 * kBigWarps, 0, s>>>(logits, q, index, logp, rows, C);
    }
    return dsb::check_launch("sample_categorical");
}
