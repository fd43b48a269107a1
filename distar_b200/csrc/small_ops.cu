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
