// Masked softmax (forward / backward) between the two tensor-core products of the entity self-attention
// (model/module_utils.py:99-107: `score /= sqrt(d); masked_fill_(~mask, -1e9); softmax; matmul(score, value)`).
// Scores arrive fp32 from the Q.K^T GEMM (scale already applied in its epilogue); probabilities leave as the
// bf16 (hi, lo) pair the P.V GEMM consumes, so P is written once and never in fp32.  One warp per row; a row is
// S = 128*k keys (512 entities).  Keys >= entity_num[obs] are masked to -1e9 exactly like the reference, which
// makes their probability exactly 0 (and a fully masked row uniform).
#include <math_constants.h>
#include "common.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kMaxVec = 8;           // S <= 8 * 128 = 1024

__device__ __forceinline__ void store_split4(__nv_bfloat16* hi, __nv_bfloat16* lo, int64_t off, float4 v) {
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
    const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y);
    const __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
    *reinterpret_cast<uint2*>(hi + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    *reinterpret_cast<uint2*>(lo + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
}

__device__ __forceinline__ float4 load_split4(const __nv_bfloat16* hi, const __nv_bfloat16* lo, int64_t off) {
    const uint2 h = *reinterpret_cast<const uint2*>(hi + off), l = *reinterpret_cast<const uint2*>(lo + off);
    const float2 h0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h.x));
    const float2 h1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h.y));
    const float2 l0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&l.x));
    const float2 l1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&l.y));
    return make_float4(h0.x + l0.x, h0.y + l0.y, h1.x + l1.x, h1.y + l1.y);
}

__global__ void attn_softmax_fwd_kernel(const float* __restrict__ scores, const int64_t* __restrict__ entity_num,
                                        int rows_per_obs, __nv_bfloat16* __restrict__ p_hi,
                                        __nv_bfloat16* __restrict__ p_lo, int64_t rows, int S) {
    const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int lane = threadIdx.x & 31;
    const int nk = entity_num ? (int)entity_num[r / rows_per_obs] : S;
    const int nvec = S / 128;
    float4 v[kMaxVec];
    float m = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
        if (i < nvec) {
            const int c = i * 128 + lane * 4;
            float4 x = *reinterpret_cast<const float4*>(scores + r * S + c);
            if (c + 0 >= nk) x.x = -1e9f;
            if (c + 1 >= nk) x.y = -1e9f;
            if (c + 2 >= nk) x.z = -1e9f;
            if (c + 3 >= nk) x.w = -1e9f;
            v[i] = x;
            m = fmaxf(m, fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w)));
        }
    }
    m = dsb::warp_max(m);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
        if (i < nvec) {
            v[i].x = expf(v[i].x - m); v[i].y = expf(v[i].y - m); v[i].z = expf(v[i].z - m); v[i].w = expf(v[i].w - m);
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    s = dsb::warp_sum(s);
    const float inv = 1.0f / s;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
        if (i < nvec) {
            const float4 p = make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv);
            store_split4(p_hi, p_lo, r * S + i * 128 + lane * 4, p);
        }
    }
}

// dS = P * (dP - sum_k dP_k P_k)   (gradient wrt the already scaled scores)
__global__ void attn_softmax_bwd_kernel(const __nv_bfloat16* __restrict__ p_hi, const __nv_bfloat16* __restrict__ p_lo,
                                        const float* __restrict__ dp, const int64_t* __restrict__ entity_num,
                                        int rows_per_obs, __nv_bfloat16* __restrict__ ds_hi,
                                        __nv_bfloat16* __restrict__ ds_lo, int64_t rows, int S) {
    const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int lane = threadIdx.x & 31;
    const int nvec = S / 128;
    // masked_fill_ blocks the gradient of masked keys (only visible for a fully masked row, where P is uniform)
    const int nk = entity_num ? (int)entity_num[r / rows_per_obs] : S;
    float4 p[kMaxVec], g[kMaxVec];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
        if (i < nvec) {
            const int64_t off = r * S + i * 128 + lane * 4;
            p[i] = load_split4(p_hi, p_lo, off);
            g[i] = *reinterpret_cast<const float4*>(dp + off);
            dot += (p[i].x * g[i].x + p[i].y * g[i].y) + (p[i].z * g[i].z + p[i].w * g[i].w);
        }
    }
    dot = dsb::warp_sum(dot);
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
        if (i < nvec) {
            const int c = i * 128 + lane * 4;
            float4 d = make_float4(p[i].x * (g[i].x - dot), p[i].y * (g[i].y - dot), p[i].z * (g[i].z - dot),
                                   p[i].w * (g[i].w - dot));
            if (c + 0 >= nk) d.x = 0.f;
            if (c + 1 >= nk) d.y = 0.f;
            if (c + 2 >= nk) d.z = 0.f;
            if (c + 3 >= nk) d.w = 0.f;
            store_split4(ds_hi, ds_lo, r * S + i * 128 + lane * 4, d);
        }
    }
}

}  // namespace

extern "C" int dsb_attn_softmax_fwd(const float* scores, const int64_t* entity_num, int rows_per_obs, void* p_hi,
                                    void* p_lo, int64_t rows, int S, dsb_stream_t stream) {
    DSB_REQUIRE(scores && p_hi && p_lo && rows >= 0 && rows_per_obs > 0, "attn_softmax_fwd: bad argument");
    DSB_REQUIRE(S % 128 == 0 && S <= 128 * kMaxVec, "attn_softmax_fwd: S must be a multiple of 128, <= %d", 128 * kMaxVec);
    if (rows == 0) return DSB_OK;
    const int64_t blocks = (rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
    DSB_REQUIRE(blocks < (1ll << 31), "attn_softmax_fwd: too many rows");
    attn_softmax_fwd_kernel<<<(unsigned)blocks, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        scores, entity_num, rows_per_obs, (__nv_bfloat16*)p_hi, (__nv_bfloat16*)p_lo, rows, S);
    return dsb::check_launch("attn_softmax_fwd");
}

extern "C" int dsb_attn_softmax_bwd(const void* p_hi, const void* p_lo, const float* dp, const int64_t* entity_num,
                                    int rows_per_obs, void* ds_hi, void* ds_lo, int64_t rows, int S,
                                    dsb_stream_t stream) {
    DSB_REQUIRE(p_hi && p_lo && dp && ds_hi && ds_lo && rows >= 0, "attn_softmax_bwd: bad argument");
    DSB_REQUIRE(S % 128 == 0 && S <= 128 * kMaxVec, "attn_softmax_bwd: S must be a multiple of 128, <= %d", 128 * kMaxVec);
    if (rows == 0) return DSB_OK;
    const int64_t blocks = (rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
    DSB_REQUIRE(blocks < (1ll << 31), "attn_softmax_bwd: too many rows");
    attn_softmax_bwd_kernel<<<(unsigned)blocks, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)p_hi, (const __nv_bfloat16*)p_lo, dp, entity_num, rows_per_obs > 0 ? rows_per_obs : 1,
        (__nv_bfloat16*)ds_hi, (__nv_bfloat16*)ds_lo, rows, S);
    return dsb::check_launch("attn_softmax_bwd");
}
