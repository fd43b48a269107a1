// Small operators of the scalar encoder and the action heads: the pieces between the tensor-core
// GEMMs that the reference runs as strings of tiny ATen launches.
//   pack_pair        any small / oddly sized / integer-typed activation -> the zero-padded bf16 (hi, lo) operand pair of the
//                    tcgen05 GEMM (scalar_encoder.py:99-132 `.float()` + fc_block inputs; head MLPs with K or N not tileable)
//   glu_gate         GLU gate  sigmoid(g) * x   (module_utils.py:508-524) forward / backward
//   onehot_linear    act(W . one_hot(idx) + b) as a gather of one weight column (or row) per sample: the action embeddings of the
//                    auto-regressive heads (action_type_head.py:61-63, action_arg_head.py:49-52,82-85) and the scalar
//                    encoder's nn.Embedding lookups (scalar_encoder.py:105-116); backward scatters into the weight gradient
//   target_unit      TargetUnitHead logits  key[p,e,:] . query[p,:]  with the entity mask and temperature
//                    (action_arg_head.py:343-363) forward / backward, one warp per 4 entities
#include <cuda_fp16.h>
#include <math_constants.h>
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float load_as_float(const void* p, int dtype, int64_t i) {
    switch (dtype) {
        case 0: return (float)reinterpret_cast<const uint8_t*>(p)[i];
        case 1: return (float)reinterpret_cast<const int16_t*>(p)[i];
        case 2: return (float)reinterpret_cast<const int8_t*>(p)[i];
        case 3: return __half2float(reinterpret_cast<const __half*>(p)[i]);
        case 5: return (float)reinterpret_cast<const int64_t*>(p)[i];
        default: return reinterpret_cast<const float*>(p)[i];
    }
}

// one thread per (row, padded column pair)
__global__ void pack_pair_kernel(const void* __restrict__ x, int dtype, int64_t rows, int K, int64_t ld_in,
                                 __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int Kp) {
    const int64_t total = rows * (Kp / 2);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / (Kp / 2);
        const int c = (int)(i - r * (Kp / 2)) * 2;
        const float a = c < K ? load_as_float(x, dtype, r * ld_in + c) : 0.f;
        const float b = c + 1 < K ? load_as_float(x, dtype, r * ld_in + c + 1) : 0.f;
        const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        reinterpret_cast<__nv_bfloat162*>(hi)[i] = h;
        if (lo) {
            const float2 hf = __bfloat1622float2(h);
            reinterpret_cast<__nv_bfloat162*>(lo)[i] = __floats2bfloat162_rn(a - hf.x, b - hf.y);
        }
    }
}

__device__ __forceinline__ float sigmoid_(float v) { return 1.0f / (1.0f + __expf(-v)); }

__global__ void glu_fwd_kernel(const float4* __restrict__ g, const float4* __restrict__ x, float4* __restrict__ out, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 a = g[i], b = x[i];
        out[i] = make_float4(sigmoid_(a.x) * b.x, sigmoid_(a.y) * b.y, sigmoid_(a.z) * b.z, sigmoid_(a.w) * b.w);
    }
}
__global__ void glu_bwd_kernel(const float4* __restrict__ go, const float4* __restrict__ g, const float4* __restrict__ x,
                               float4* __restrict__ dg, float4* __restrict__ dx, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 o = go[i], a = g[i], b = x[i];
        const float s0 = sigmoid_(a.x), s1 = sigmoid_(a.y), s2 = sigmoid_(a.z), s3 = sigmoid_(a.w);
        dx[i] = make_float4(o.x * s0, o.y * s1, o.z * s2, o.w * s3);
        dg[i] = make_float4(o.x * b.x * s0 * (1.f - s0), o.y * b.y * s1 * (1.f - s1), o.z * b.z * s2 * (1.f - s2),
                            o.w * b.w * s3 * (1.f - s3));
    }
}

// out[p, j] = act(W[j * sj + idx_p * si] + b[j]);  idx clamped into [0, C) (torch's F.one_hot / embedding raise outside it:
// bit 4 of error_flag records that; a negative id additionally keeps the reference's clamp-max semantics of
// scalar_encoder.py:110-114 out of the picture: the caller decides with clamp_max whether >= C is an error or a clamp)
__global__ void onehot_fwd_kernel(const float* __restrict__ W, const float* __restrict__ b, const int64_t* __restrict__ idx,
                                  float* __restrict__ out, int64_t P, int N, int C, int64_t sj, int64_t si, int relu,
                                  int clamp_max, int* __restrict__ error_flag) {
    const int64_t total = P * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / N;
        const int j = (int)(i - p * N);
        int64_t a = idx[p];
        if (a < 0 || a >= C) {
            if (!(clamp_max && a >= C) && error_flag && j == 0) atomicOr(error_flag, 4);
            a = a < 0 ? 0 : C - 1;
        }
        float v = W[j * sj + a * si] + (b ? b[j] : 0.f);
        out[i] = relu ? fmaxf(v, 0.f) : v;
    }
}
// dW[j * sj + idx_p * si] += g[p, j] * (out > 0), db[j] += the same; rows of the same class collide -> atomics
__global__ void onehot_bwd_kernel(const float* __restrict__ go, const float* __restrict__ out, const int64_t* __restrict__ idx,
                                  float* __restrict__ dW, float* __restrict__ db, int64_t P, int N, int C, int64_t sj,
                                  int64_t si, int relu) {
    const int64_t total = P * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / N;
        const int j = (int)(i - p * N);
        float g = go[i];
        if (relu && !(out[i] > 0.f)) g = 0.f;
        if (g == 0.f) continue;
        int64_t a = idx[p];
        a = a < 0 ? 0 : (a >= C ? C - 1 : a);
        atomicAdd(dW + j * sj + a * si, g);
        if (db) atomicAdd(db + j, g);
    }
}

// ---- target-unit head: logits[p, e] = (e < entity_num[p] ? key[p, e, :] . q[p, :] : -1e9) / T, key rows of 32 floats with
// row pitch ldk (the key projection shares its GEMM output with the selected-units head: ldk = 64)
constexpr int kKey = 32;
__global__ void target_unit_fwd_kernel(const float* __restrict__ key, int ldk, const float* __restrict__ q,
                                       const int64_t* __restrict__ entity_num, float* __restrict__ logits, int64_t P, int E,
                                       float inv_t) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int groups = E / 4;                                   // 4 entities per warp iteration: 8 lanes x float4 each
    const int64_t total = P * groups;
    if (warp >= total) return;
    const int64_t p = warp / groups;
    const int e = (int)(warp - p * groups) * 4 + (lane >> 3);
    const int part = lane & 7;
    const float4 k4 = *reinterpret_cast<const float4*>(key + ((int64_t)p * E + e) * ldk + part * 4);
    const float4 q4 = *reinterpret_cast<const float4*>(q + p * kKey + part * 4);
    float acc = k4.x * q4.x + k4.y * q4.y + k4.z * q4.z + k4.w * q4.w;
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (part == 0) logits[p * E + e] = (e < entity_num[p] ? acc : -1e9f) * inv_t;
}
// dq[p, :] = sum_e g[p, e] key[p, e, :],  dkey[p, e, :] = g[p, e] q[p, :]  with g = grad_logits / T at valid entities, 0 elsewhere.
// One CTA per p: 8 warps stride over the entities (4 per iteration), partial dq reduced through shared memory.
__global__ void __launch_bounds__(kThreads)
target_unit_bwd_kernel(const float* __restrict__ gl, const float* __restrict__ key, int ldk, const float* __restrict__ q,
                       const int64_t* __restrict__ entity_num, float* __restrict__ dkey, int lddk, float* __restrict__ dq,
                       int E, float inv_t) {
    __shared__ float red[kThreads / 32][kKey];
    const int64_t p = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int part = lane & 7, sub = lane >> 3;
    const int en = (int)entity_num[p];
    const float4 q4 = *reinterpret_cast<const float4*>(q + p * kKey + part * 4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e0 = warp * 4; e0 < E; e0 += (kThreads / 32) * 4) {
        const int e = e0 + sub;
        const float g = e < en ? gl[p * E + e] * inv_t : 0.f;
        const float4 k4 = *reinterpret_cast<const float4*>(key + ((int64_t)p * E + e) * ldk + part * 4);
        acc.x += g * k4.x; acc.y += g * k4.y; acc.z += g * k4.z; acc.w += g * k4.w;
        *reinterpret_cast<float4*>(dkey + ((int64_t)p * E + e) * lddk + part * 4) = make_float4(g * q4.x, g * q4.y, g * q4.z, g * q4.w);
    }
    // lanes with the same `part` hold partial sums of the same 4 query components: fold the 4 sub-groups, then the warps
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 8);  acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 8);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 8);  acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 8);
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 16); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 16);
    if (lane < 8) {
        red[warp][part * 4 + 0] = acc.x; red[warp][part * 4 + 1] = acc.y;
        red[warp][part * 4 + 2] = acc.z; red[warp][part * 4 + 3] = acc.w;
    }
    __syncthreads();
    if (threadIdx.x < kKey) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) t += red[w][threadIdx.x];
        dq[p * kKey + threadIdx.x] = t;
    }
}

// ---- LayerNorm over narrow rows (D = 32 * NPL <= 128): the 64-wide pre-LN transformer of the beginning-build-order encoder
// (scalar_encoder.py:19-57, module_utils.py:130-151).  One warp per row.  Backward adds gamma / beta gradients with atomics.
template <int NPL>
__global__ void ln_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                    float* __restrict__ y, float* __restrict__ stats, int64_t rows, float eps) {
    constexpr int D = 32 * NPL;
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= rows) return;
    float v[NPL], s = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) { v[i] = x[r * D + i * 32 + lane]; s += v[i]; }
    const float mean = dsb::warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) { const float d = v[i] - mean; q += d * d; }
    const float rstd = rsqrtf(dsb::warp_sum(q) * (1.0f / D) + eps);
#pragma unroll
    for (int i = 0; i < NPL; ++i) y[r * D + i * 32 + lane] = (v[i] - mean) * rstd * gamma[i * 32 + lane] + beta[i * 32 + lane];
    if (lane == 0) { stats[2 * r] = mean; stats[2 * r + 1] = rstd; }
}
template <int NPL>
__global__ void ln_small_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x, const float* __restrict__ gamma,
                                    const float* __restrict__ stats, float* __restrict__ gx, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, int64_t rows, int rows_per_warp) {
    constexpr int D = 32 * NPL;
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float ag[NPL], ab[NPL];
#pragma unroll
    for (int i = 0; i < NPL; ++i) { ag[i] = 0.f; ab[i] = 0.f; }
    for (int64_t r = w * rows_per_warp; r < (w + 1) * rows_per_warp && r < rows; ++r) {
        const float mean = stats[2 * r], rstd = stats[2 * r + 1];
        float g[NPL], xh[NPL], s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            const float go = gy[r * D + i * 32 + lane];
            xh[i] = (x[r * D + i * 32 + lane] - mean) * rstd;
            ag[i] += go * xh[i];
            ab[i] += go;
            g[i] = go * gamma[i * 32 + lane];
            s1 += g[i];
            s2 += g[i] * xh[i];
        }
        s1 = dsb::warp_sum(s1) * (1.0f / D);
        s2 = dsb::warp_sum(s2) * (1.0f / D);
#pragma unroll
        for (int i = 0; i < NPL; ++i) gx[r * D + i * 32 + lane] = rstd * (g[i] - s1 - xh[i] * s2);
    }
#pragma unroll
    for (int i = 0; i < NPL; ++i) { atomicAdd(dgamma + i * 32 + lane, ag[i]); atomicAdd(dbeta + i * 32 + lane, ab[i]); }
}

// ---- unmasked multi-head self-attention over short sequences (S <= 32 tokens, head_dim <= 16): the 20-token transformer of
// the beginning-build-order encoder (module_utils.py:88-111 with heads = 2, head_dim = 8).  qkv [B, S, 3 * H * HD] (q | k | v,
// each head-major), out [B, S, H * HD].  One warp per (sequence, head); lane = query token; K and V of the head in shared memory.
template <int HD>
__global__ void attn_small_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out, int64_t B, int S, int H) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    float* ks = sm + wib * 2 * 32 * HD;
    float* vs = ks + 32 * HD;
    if (w >= B * H) return;
    const int64_t b = w / H;
    const int h = (int)(w - b * H);
    const int ld = 3 * H * HD;
    const float scale = rsqrtf((float)HD);
    float q[HD];
    if (lane < S) {
        const float* row = qkv + (b * S + lane) * ld;
#pragma unroll
        for (int d = 0; d < HD; ++d) {
            q[d] = row[h * HD + d];
            ks[lane * HD + d] = row[H * HD + h * HD + d];
            vs[lane * HD + d] = row[2 * H * HD + h * HD + d];
        }
    }
    __syncwarp();
    if (lane >= S) return;
    float sc[32], mx = -CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if (j < S) {
            float a = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) a = fmaf(q[d], ks[j * HD + d], a);
            sc[j] = a * scale;
            mx = fmaxf(mx, sc[j]);
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) if (j < S) { sc[j] = __expf(sc[j] - mx); sum += sc[j]; }
    const float inv = 1.0f / sum;
    float o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) if (j < S) {
        const float pj = sc[j] * inv;
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] = fmaf(pj, vs[j * HD + d], o[d]);
    }
#pragma unroll
    for (int d = 0; d < HD; ++d) out[(b * S + lane) * (H * HD) + h * HD + d] = o[d];
}
template <int HD>
__global__ void attn_small_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ gout, float* __restrict__ gqkv,
                                      int64_t B, int S, int H) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    float* ks = sm + wib * 2 * 32 * HD;
    float* vs = ks + 32 * HD;
    if (w >= B * H) return;
    const int64_t b = w / H;
    const int h = (int)(w - b * H);
    const int ld = 3 * H * HD;
    const float scale = rsqrtf((float)HD);
    const bool live = lane < S;
    float q[HD], go[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) { q[d] = 0.f; go[d] = 0.f; }
    if (live) {
        const float* row = qkv + (b * S + lane) * ld;
#pragma unroll
        for (int d = 0; d < HD; ++d) {
            q[d] = row[h * HD + d];
            ks[lane * HD + d] = row[H * HD + h * HD + d];
            vs[lane * HD + d] = row[2 * H * HD + h * HD + d];
            go[d] = gout[(b * S + lane) * (H * HD) + h * HD + d];
        }
    }
    __syncwarp();
    float p[32], mx = -CUDART_INF_F, sum = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        p[j] = 0.f;
        if (j < S) {
            float a = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) a = fmaf(q[d], ks[j * HD + d], a);
            p[j] = a * scale;
            mx = fmaxf(mx, p[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) if (j < S) { p[j] = __expf(p[j] - mx); sum += p[j]; }
    const float inv = live ? 1.0f / sum : 0.f;
    float dsum = 0.f, dp[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        dp[j] = 0.f;
        if (j < S) {
            p[j] *= inv;
            float a = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) a = fmaf(go[d], vs[j * HD + d], a);
            dp[j] = a;
            dsum = fmaf(p[j], a, dsum);
        }
    }
    float dq[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) dq[d] = 0.f;
    // per key j: ds_ij for this lane's query; dk_j / dv_j are sums over the queries (= lanes)
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if (j >= S) continue;                                   // (S is warp-uniform)
        const float ds = live ? p[j] * (dp[j] - dsum) * scale : 0.f;
        const float pj = live ? p[j] : 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) {
            dq[d] = fmaf(ds, ks[j * HD + d], dq[d]);
            const float dk = dsb::warp_sum(ds * q[d]);
            const float dv = dsb::warp_sum(pj * go[d]);
            if (lane == 0) {
                gqkv[(b * S + j) * ld + H * HD + h * HD + d] = dk;
                gqkv[(b * S + j) * ld + 2 * H * HD + h * HD + d] = dv;
            }
        }
    }
    if (live) {
#pragma unroll
        for (int d = 0; d < HD; ++d) gqkv[(b * S + lane) * ld + h * HD + d] = dq[d];
    }
}

// ---- token features of the beginning-build-order encoder (scalar_encoder.py:33-45): one-hot(action, 174) | one-hot(position, 20) |
// 10-bit binary x | 10-bit binary y of the build location, written as the EXACT bf16 A operand [B * 20, Kp] of the embedding GEMM
__global__ void bo_tokens_kernel(const int16_t* __restrict__ order, const int16_t* __restrict__ loc, int spatial_x,
                                 __nv_bfloat16* __restrict__ hi, int64_t B, int L, int A, int Kp) {
    const int64_t total = B * L * Kp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t tok = i / Kp;
        const int c = (int)(i - tok * Kp);
        const int pos = (int)(tok % L);
        float v = 0.f;
        if (c < A) {
            int a = order[tok];
            a = a < 0 ? 0 : (a >= A ? A - 1 : a);
            v = c == a ? 1.f : 0.f;
        } else if (c < A + L) {
            v = (c - A) == pos ? 1.f : 0.f;
        } else if (c < A + L + 20) {
            const int l = loc[tok];
            const int bit = c - A - L;
            const int val = bit < 10 ? l % spatial_x : l / spatial_x;
            v = (float)((val >> (9 - (bit % 10))) & 1);
        }
        hi[i] = __float2bfloat16_rn(v);
    }
}

inline unsigned grid_for(int64_t n) {
    int64_t blocks = (n + kThreads - 1) / kThreads;
    const int64_t cap = 148 * 16;
    return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace

extern "C" int dsb_pack_pair(const void* x, int dtype, int64_t rows, int K, int64_t ld_in, void* hi, void* lo, int Kp,
                             dsb_stream_t stream) {
    DSB_REQUIRE(x && hi && rows >= 0 && K > 0 && Kp >= K && Kp % 8 == 0 && ld_in >= K, "pack_pair: bad argument");
    DSB_REQUIRE(dtype >= 0 && dtype <= 5, "pack_pair: dtype must be 0 u8, 1 i16, 2 i8, 3 f16, 4 f32, 5 i64");
    if (rows == 0) return DSB_OK;
    pack_pair_kernel<<<grid_for(rows * (Kp / 2)), kThreads, 0, (cudaStream_t)stream>>>(
        x, dtype, rows, K, ld_in, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, Kp);
    return dsb::check_launch("pack_pair");
}

extern "C" int dsb_glu_gate_fwd(const float* gate, const float* x, float* out, int64_t n, dsb_stream_t stream) {
    DSB_REQUIRE(gate && x && out && n >= 0 && n % 4 == 0, "glu_gate_fwd: bad argument (n %% 4)");
    if (n == 0) return DSB_OK;
    glu_fwd_kernel<<<grid_for(n / 4), kThreads, 0, (cudaStream_t)stream>>>((const float4*)gate, (const float4*)x, (float4*)out, n / 4);
    return dsb::check_launch("glu_gate_fwd");
}
extern "C" int dsb_glu_gate_bwd(const float* grad_out, const float* gate, const float* x, float* grad_gate, float* grad_x,
                                int64_t n, dsb_stream_t stream) {
    DSB_REQUIRE(grad_out && gate && x && grad_gate && grad_x && n >= 0 && n % 4 == 0, "glu_gate_bwd: bad argument (n %% 4)");
    if (n == 0) return DSB_OK;
    glu_bwd_kernel<<<grid_for(n / 4), kThreads, 0, (cudaStream_t)stream>>>((const float4*)grad_out, (const float4*)gate,
                                                                          (const float4*)x, (float4*)grad_gate, (float4*)grad_x, n / 4);
    return dsb::check_launch("glu_gate_bwd");
}

extern "C" int dsb_onehot_linear_fwd(const float* W, const float* bias, const int64_t* idx, float* out, int64_t P, int N,
                                     int C, int64_t stride_out, int64_t stride_class, int relu, int clamp_max,
                                     int* error_flag, dsb_stream_t stream) {
    DSB_REQUIRE(W && idx && out && P >= 0 && N > 0 && C > 0, "onehot_linear_fwd: bad argument");
    if (P == 0) return DSB_OK;
    onehot_fwd_kernel<<<grid_for(P * N), kThreads, 0, (cudaStream_t)stream>>>(W, bias, idx, out, P, N, C, stride_out,
                                                                             stride_class, relu, clamp_max, error_flag);
    return dsb::check_launch("onehot_linear_fwd");
}
extern "C" int dsb_onehot_linear_bwd(const float* grad_out, const float* out, const int64_t* idx, float* grad_W,
                                     float* grad_bias, int64_t P, int N, int C, int64_t stride_out, int64_t stride_class,
                                     int relu, dsb_stream_t stream) {
    DSB_REQUIRE(grad_out && idx && grad_W && (!relu || out) && P >= 0 && N > 0 && C > 0, "onehot_linear_bwd: bad argument");
    if (P == 0) return DSB_OK;
    onehot_bwd_kernel<<<grid_for(P * N), kThreads, 0, (cudaStream_t)stream>>>(grad_out, out, idx, grad_W, grad_bias, P, N, C,
                                                                             stride_out, stride_class, relu);
    return dsb::check_launch("onehot_linear_bwd");
}

extern "C" int dsb_target_unit_fwd(const float* key, int ldk, const float* query, const int64_t* entity_num, float* logits,
                                   int64_t P, int E, float temperature, dsb_stream_t stream) {
    DSB_REQUIRE(key && query && entity_num && logits && P >= 0 && E > 0 && E % 4 == 0 && ldk >= kKey && ldk % 4 == 0 &&
                temperature > 0.f, "target_unit_fwd: bad argument");
    if (P == 0) return DSB_OK;
    const int64_t warps = P * (E / 4);
    target_unit_fwd_kernel<<<(unsigned)((warps * 32 + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        key, ldk, query, entity_num, logits, P, E, 1.0f / temperature);
    return dsb::check_launch("target_unit_fwd");
}
extern "C" int dsb_target_unit_bwd(const float* grad_logits, const float* key, int ldk, const float* query,
                                   const int64_t* entity_num, float* grad_key, int ldgk, float* grad_query, int64_t P, int E,
                                   float temperature, dsb_stream_t stream) {
    DSB_REQUIRE(grad_logits && key && query && entity_num && grad_key && grad_query && P >= 0 && E > 0 && E % 4 == 0 &&
                ldk >= kKey && ldk % 4 == 0 && ldgk >= kKey && ldgk % 4 == 0 && temperature > 0.f, "target_unit_bwd: bad argument");
    if (P == 0) return DSB_OK;
    target_unit_bwd_kernel<<<(unsigned)P, kThreads, 0, (cudaStream_t)stream>>>(grad_logits, key, ldk, query, entity_num, grad_key,
                                                                             ldgk, grad_query, E, 1.0f / temperature);
    return dsb::check_launch("target_unit_bwd");
}

extern "C" int dsb_ln_small_supported(int D) { return (D == 32 || D == 64 || D == 96) ? 1 : 0; }

extern "C" int dsb_ln_small_fwd(const float* x, const float* gamma, const float* beta, float* y, float* stats, int64_t rows, int D,
                                float eps, dsb_stream_t stream) {
    DSB_REQUIRE(x && gamma && beta && y && stats && rows >= 0 && dsb_ln_small_supported(D), "ln_small_fwd: bad argument (D %d)", D);
    if (rows == 0) return DSB_OK;
    const unsigned grid = (unsigned)((rows * 32 + kThreads - 1) / kThreads);
    cudaStream_t s = (cudaStream_t)stream;
    if (D == 32) ln_small_fwd_kernel<1><<<grid, kThreads, 0, s>>>(x, gamma, beta, y, stats, rows, eps);
    else if (D == 64) ln_small_fwd_kernel<2><<<grid, kThreads, 0, s>>>(x, gamma, beta, y, stats, rows, eps);
    else ln_small_fwd_kernel<3><<<grid, kThreads, 0, s>>>(x, gamma, beta, y, stats, rows, eps);
    return dsb::check_launch("ln_small_fwd");
}
extern "C" int dsb_ln_small_bwd(const float* gy, const float* x, const float* gamma, const float* stats, float* gx, float* dgamma,
                                float* dbeta, int64_t rows, int D, dsb_stream_t stream) {
    DSB_REQUIRE(gy && x && gamma && stats && gx && dgamma && dbeta && rows >= 0 && dsb_ln_small_supported(D),
                "ln_small_bwd: bad argument (D %d)", D);
    if (rows == 0) return DSB_OK;
    const int64_t warps = rows < 148 * 32 ? rows : 148 * 32;
    const int rpw = (int)((rows + warps - 1) / warps);
    const unsigned grid = (unsigned)((warps * 32 + kThreads - 1) / kThreads);
    cudaStream_t s = (cudaStream_t)stream;
    if (D == 32) ln_small_bwd_kernel<1><<<grid, kThreads, 0, s>>>(gy, x, gamma, stats, gx, dgamma, dbeta, rows, rpw);
    else if (D == 64) ln_small_bwd_kernel<2><<<grid, kThreads, 0, s>>>(gy, x, gamma, stats, gx, dgamma, dbeta, rows, rpw);
    else ln_small_bwd_kernel<3><<<grid, kThreads, 0, s>>>(gy, x, gamma, stats, gx, dgamma, dbeta, rows, rpw);
    return dsb::check_launch("ln_small_bwd");
}

extern "C" int dsb_attn_small_fwd(const float* qkv, float* out, int64_t B, int S, int H, int HD, dsb_stream_t stream) {
    DSB_REQUIRE(qkv && out && B >= 0 && S > 0 && S <= 32 && H > 0 && (HD == 8 || HD == 16), "attn_small_fwd: bad argument");
    if (B == 0) return DSB_OK;
    const int wpb = kThreads / 32;
    const unsigned grid = (unsigned)((B * H + wpb - 1) / wpb);
    const size_t smem = (size_t)wpb * 2 * 32 * HD * sizeof(float);
    cudaStream_t s = (cudaStream_t)stream;
    if (HD == 8) attn_small_fwd_kernel<8><<<grid, kThreads, smem, s>>>(qkv, out, B, S, H);
    else attn_small_fwd_kernel<16><<<grid, kThreads, smem, s>>>(qkv, out, B, S, H);
    return dsb::check_launch("attn_small_fwd");
}
extern "C" int dsb_attn_small_bwd(const float* qkv, const float* grad_out, float* grad_qkv, int64_t B, int S, int H, int HD,
                                  dsb_stream_t stream) {
    DSB_REQUIRE(qkv && grad_out && grad_qkv && B >= 0 && S > 0 && S <= 32 && H > 0 && (HD == 8 || HD == 16), "attn_small_bwd: bad argument");
    if (B == 0) return DSB_OK;
    const int wpb = kThreads / 32;
    const unsigned grid = (unsigned)((B * H + wpb - 1) / wpb);
    const size_t smem = (size_t)wpb * 2 * 32 * HD * sizeof(float);
    cudaStream_t s = (cudaStream_t)stream;
    if (HD == 8) attn_small_bwd_kernel<8><<<grid, kThreads, smem, s>>>(qkv, grad_out, grad_qkv, B, S, H);
    else attn_small_bwd_kernel<16><<<grid, kThreads, smem, s>>>(qkv, grad_out, grad_qkv, B, S, H);
    return dsb::check_launch("attn_small_bwd");
}

extern "C" int dsb_bo_tokens(const int16_t* beginning_order, const int16_t* bo_location, int spatial_x, void* hi, int64_t B, int L,
                             int num_actions, int Kp, dsb_stream_t stream) {
    DSB_REQUIRE(beginning_order && bo_location && hi && B >= 0 && L > 0 && spatial_x > 0 && Kp >= num_actions + L + 20,
                "bo_tokens: bad argument");
    if (B == 0) return DSB_OK;
    bo_tokens_kernel<<<grid_for(B * L * Kp), kThreads, 0, (cudaStream_t)stream>>>(beginning_order, bo_location, spatial_x,
                                                                                  (__nv_bfloat16*)hi, B, L, num_actions, Kp);
    return dsb::check_launch("bo_tokens");
}
