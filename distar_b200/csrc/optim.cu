// Flat-arena optimiser step: global grad L2 norm -> clip -> Adam, plus the fp32 -> bf16 (hi, lo) split.
// Replaces torch.nn.utils.clip_grad_norm_ (ctools/torch_utils/grad_clip.py:141-144) + torch.optim.Adam.step
// (rl_learner.py:73-80,132: betas (0, 0.99), eps 1e-5, no weight decay) — ~1.2 k foreach launches in the
// reference — by two streaming passes over one contiguous fp32 arena; the DP average 1/world of
// DistModule.sync_gradients (ctools/utils/dist_helper.py:421-431) is folded into the same pass.
#include "common.cuh"

namespace {

constexpr int kPartials = 1024;
constexpr int kThreads = 256;

__global__ void sumsq_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ partial) {
    float acc = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = n / 4;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = g4[i];
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += g[i] * g[i];
    __shared__ float red[kThreads / 32];
    acc = dsb::warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < kThreads / 32; ++i) t += red[i];
        partial[blockIdx.x] = t;
    }
}

__global__ void norm_finish_kernel(const float* __restrict__ partial, int np, float scale, float* __restrict__ norm_out) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) acc += (double)partial[i];
    __shared__ double red[32];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        norm_out[0] = (float)sqrt(t) * scale;      // scale = 1/world: the norm of the AVERAGED gradient (dist_helper.py:421-431)
    }
}

__device__ __forceinline__ void split_store(float x, __nv_bfloat16* hi, __nv_bfloat16* lo, int64_t i) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(x - __bfloat162float(h));
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, const float* __restrict__ norm, float max_norm,
                            float grad_scale, float lr, float b1, float b2, float eps, float weight_decay, float bc1,
                            float bc2_sqrt, __nv_bfloat16* __restrict__ sh, __nv_bfloat16* __restrict__ sl,
                            const float* __restrict__ skip) {
    if (skip && skip[0] != 0.f) return;      // some rank flagged its batch as invalid: no weight moves anywhere
    float scale = grad_scale;
    if (norm && max_norm > 0.f) {
        // clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1 (norm[0] is already the averaged norm)
        const float total = norm[0];
        scale *= fminf(1.0f, max_norm / (total + 1e-6f));
    }
    const float step_size = lr / bc1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        // torch.optim.Adam (not AdamW): the L2 term joins the (already clipped) gradient, adam.py `grad.add(param, alpha=wd)`
        const float gi = fmaf(weight_decay, p[i], g[i] * scale);
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        const float pi = p[i] - step_size * (mi / denom);
        p[i] = pi;
        if (sh) split_store(pi, sh, sl, i);
    }
}

__global__ void split_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                             __nv_bfloat16* __restrict__ lo, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = n / 4;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    uint2* h2 = reinterpret_cast<uint2*>(hi);
    uint2* l2 = reinterpret_cast<uint2*>(lo);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = x4[i];
        const __nv_bfloat162 ha = __floats2bfloat162_rn(v.x, v.y), hb = __floats2bfloat162_rn(v.z, v.w);
        const float2 fa = __bfloat1622float2(ha), fb = __bfloat1622float2(hb);
        const __nv_bfloat162 la = __floats2bfloat162_rn(v.x - fa.x, v.y - fa.y);
        const __nv_bfloat162 lb = __floats2bfloat162_rn(v.z - fb.x, v.w - fb.y);
        h2[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&ha), *reinterpret_cast<const uint32_t*>(&hb));
        l2[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&la), *reinterpret_cast<const uint32_t*>(&lb));
    }
    for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        split_store(x[i], hi, lo, i);
}

// column sums of (hi + lo) over a band of rows per block, added atomically into out[N]; thread = 4 consecutive columns
__global__ void colsum_pair_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                   float* __restrict__ out, int64_t rows, int N, int rows_per_block) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (c >= N) return;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int64_t r = r0; r < r1; ++r) {
        const uint2 h = *reinterpret_cast<const uint2*>(hi + r * N + c), l = *reinterpret_cast<const uint2*>(lo + r * N + c);
        const float2 h0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h.x));
        const float2 h1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h.y));
        const float2 l0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&l.x));
        const float2 l1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&l.y));
        a0 += h0.x + l0.x; a1 += h0.y + l0.y; a2 += h1.x + l1.x; a3 += h1.y + l1.y;
    }
    atomicAdd(out + c, a0); atomicAdd(out + c + 1, a1); atomicAdd(out + c + 2, a2); atomicAdd(out + c + 3, a3);
}

inline unsigned grid_for(int64_t n, int per_thread) {
    int64_t blocks = (n / per_thread + kThreads - 1) / kThreads;
    const int64_t cap = 148 * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace

extern "C" int dsb_sumsq_partials(void) { return kPartials; }

extern "C" int dsb_grad_norm(const float* grad, int64_t n, float* partial, float* norm_out, float scale,
                             dsb_stream_t stream) {
    DSB_REQUIRE(grad && partial && norm_out && n > 0, "grad_norm: bad argument");
    DSB_REQUIRE((reinterpret_cast<uintptr_t>(grad) & 15) == 0, "grad_norm: grad must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    sumsq_kernel<<<kPartials, kThreads, 0, s>>>(grad, n, partial);
    int rc = dsb::check_launch("grad_norm/sumsq");
    if (rc) return rc;
    norm_finish_kernel<<<1, 256, 0, s>>>(partial, kPartials, scale, norm_out);
    return dsb::check_launch("grad_norm/finish");
}

extern "C" int dsb_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                             const float* norm, float max_norm, float grad_scale, float lr, float beta1, float beta2,
                             float eps, float weight_decay, int t, void* shadow_hi, void* shadow_lo, const float* skip_flag,
                             dsb_stream_t stream) {
    DSB_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0 && t >= 1, "adam_step: bad argument");
    DSB_REQUIRE(!shadow_hi == !shadow_lo, "adam_step: shadow_hi and shadow_lo go together");
    const float bc1 = 1.0f - powf(beta1, (float)t);
    const float bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)t));
    adam_kernel<<<grid_for(n, 4), kThreads, 0, (cudaStream_t)stream>>>(
        param, grad, exp_avg, exp_avg_sq, n, norm, max_norm, grad_scale, lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt,
        (__nv_bfloat16*)shadow_hi, (__nv_bfloat16*)shadow_lo, skip_flag);
    return dsb::check_launch("adam_step");
}

extern "C" int dsb_split_bf16(const float* x, void* hi, void* lo, int64_t n, dsb_stream_t stream) {
    DSB_REQUIRE(x && hi && lo && n >= 0, "split_bf16: bad argument");
    if (n == 0) return DSB_OK;
    DSB_REQUIRE(((reinterpret_cast<uintptr_t>(x) & 15) | (reinterpret_cast<uintptr_t>(hi) & 7) |
                 (reinterpret_cast<uintptr_t>(lo) & 7)) == 0, "split_bf16: misaligned pointer");
    split_kernel<<<grid_for(n, 8), kThreads, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, n);
    return dsb::check_launch("split_bf16");
}

// ---- ReLU backward fused with the bf16 split and the bias gradient -------------------------------------------------
// g = gy * (y > 0);  writes the (hi, lo) pair the dX / dW tensor-core GEMMs read, optionally g itself, and per-block
// partial column sums (bias gradient) — replaces a compare, a multiply, a split and a column reduction.
// Thread layout: min(N/4, 256) column threads (one float4 column group each) x as many row groups as fit in the block, so
// narrow activations (64 / 128 channels of the conv layers) still fill the block; rows are walked grid-stride with four
// independent 16-byte loads in flight per thread; the grid is capped so the partial-sum matrix stays tiny.
namespace {
constexpr int kRbThreads = 256;
constexpr int kRbMaxBlocks = 148 * 8;
constexpr int kRbRowsPerGroup = 8;            // rows one row group handles per grid-stride step

template <bool kMaskBf16>
__global__ void __launch_bounds__(kRbThreads)
relu_bwd_split_kernel(const float* __restrict__ gy, const void* __restrict__ yv, float* __restrict__ g_out,
                      __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, float* __restrict__ colsum,
                      int colsum_atomic, int64_t rows, int N) {
    __shared__ float4 part[kRbThreads];
    const float* y = kMaskBf16 ? nullptr : reinterpret_cast<const float*>(yv);
    const __nv_bfloat16* yb = kMaskBf16 ? reinterpret_cast<const __nv_bfloat16*>(yv) : nullptr;
    const int quads = N / 4;
    const int cthreads = quads < kRbThreads ? quads : kRbThreads;
    const int rgroups = kRbThreads / cthreads;
    const int ct = threadIdx.x % cthreads, rg = threadIdx.x / cthreads;
    const bool active = rg < rgroups;
    const int64_t rows_per_step = (int64_t)rgroups * kRbRowsPerGroup;
    for (int c = ct * 4; c < N; c += cthreads * 4) {           // uniform trip count over the block (ct < cthreads)
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) {
            for (int64_t r0 = (int64_t)blockIdx.x * rows_per_step; r0 < rows; r0 += (int64_t)gridDim.x * rows_per_step) {
#pragma unroll 4
                for (int i = 0; i < kRbRowsPerGroup; ++i) {
                    const int64_t r = r0 + (int64_t)i * rgroups + rg;
                    if (r >= rows) break;
                    const int64_t off = r * N + c;
                    float4 g = __ldcs(reinterpret_cast<const float4*>(gy + off));
                    if (kMaskBf16) {
                        // the ReLU output only survives as the hi half of its bf16 pair: bf16 rounding keeps sign and zero
                        const uint2 m = *reinterpret_cast<const uint2*>(yb + off);
                        g.x = ((m.x & 0xFFFFu) != 0u && !(m.x & 0x8000u)) ? g.x : 0.f;
                        g.y = ((m.x >> 16) != 0u && !(m.x & 0x80000000u)) ? g.y : 0.f;
                        g.z = ((m.y & 0xFFFFu) != 0u && !(m.y & 0x8000u)) ? g.z : 0.f;
                        g.w = ((m.y >> 16) != 0u && !(m.y & 0x80000000u)) ? g.w : 0.f;
                    } else if (y) {
                        const float4 yy = *reinterpret_cast<const float4*>(y + off);
                        g.x = yy.x > 0.f ? g.x : 0.f; g.y = yy.y > 0.f ? g.y : 0.f;
                        g.z = yy.z > 0.f ? g.z : 0.f; g.w = yy.w > 0.f ? g.w : 0.f;
                    }
                    if (g_out) *reinterpret_cast<float4*>(g_out + off) = g;
                    if (hi) {
                        const __nv_bfloat162 h0 = __floats2bfloat162_rn(g.x, g.y), h1 = __floats2bfloat162_rn(g.z, g.w);
                        const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
                        const __nv_bfloat162 l0 = __floats2bfloat162_rn(g.x - f0.x, g.y - f0.y);
                        const __nv_bfloat162 l1 = __floats2bfloat162_rn(g.z - f1.x, g.w - f1.y);
                        *reinterpret_cast<uint2*>(hi + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
                        *reinterpret_cast<uint2*>(lo + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
                    }
                    acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
                }
            }
        }
        if (colsum) {                                          // warp-uniform: colsum is a kernel argument
            if (rgroups > 1) {
                part[threadIdx.x] = acc;
                __syncthreads();
                if (rg == 0) {
                    for (int j = 1; j < rgroups; ++j) {
                        const float4 o = part[j * cthreads + ct];
                        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
                    }
                }
                __syncthreads();
            }
            if (rg == 0) {
                if (colsum_atomic) {            // accumulate straight into the bias gradient [N] (<= kRbMaxBlocks adds per column)
                    atomicAdd(reinterpret_cast<float4*>(colsum + c), acc);      // one 16-byte RED (sm_90+)
                } else {
                    *reinterpret_cast<float4*>(colsum + (int64_t)blockIdx.x * N + c) = acc;
                }
            }
        }
    }
}

constexpr int kRbMaxBlocksAtomic = 148 * 4;   // every block ends with N/4 vector REDs onto the same N floats: keep them few

inline int64_t rb_blocks(int64_t rows, int N, bool atomic = false) {
    const int quads = N / 4;
    const int cthreads = quads < kRbThreads ? quads : kRbThreads;
    const int64_t rows_per_step = (int64_t)(kRbThreads / cthreads) * kRbRowsPerGroup;
    const int64_t b = (rows + rows_per_step - 1) / rows_per_step;
    const int64_t cap = atomic ? kRbMaxBlocksAtomic : kRbMaxBlocks;
    return b < cap ? (b > 0 ? b : 1) : cap;
}
}  // namespace

extern "C" int dsb_relu_bwd_split_blocks(int64_t rows, int N) { return (int)rb_blocks(rows, N > 0 ? N : 4); }

extern "C" int dsb_relu_bwd_split(const float* gy, const void* y, int y_is_bf16, float* g_out, void* hi, void* lo,
                                  float* colsum, int colsum_atomic, int64_t rows, int N, dsb_stream_t stream) {
    DSB_REQUIRE(gy && (!hi == !lo) && (hi || g_out || colsum) && rows >= 0 && N > 0 && N % 4 == 0,
                "relu_bwd_split: bad argument (N %% 4 == 0 required)");
    if (rows == 0) return DSB_OK;
    const unsigned blocks = (unsigned)rb_blocks(rows, N, colsum && colsum_atomic);
    if (y && y_is_bf16)
        relu_bwd_split_kernel<true><<<blocks, kRbThreads, 0, (cudaStream_t)stream>>>(
            gy, y, g_out, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, colsum, colsum_atomic, rows, N);
    else
        relu_bwd_split_kernel<false><<<blocks, kRbThreads, 0, (cudaStream_t)stream>>>(
            gy, y, g_out, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, colsum, colsum_atomic, rows, N);
    return dsb::check_launch("relu_bwd_split");
}


// ---- GatedResBlock tail (module_utils.py:228-229) in one pass: out = relu(tanh(r * sigmoid(g)) * sp + x) [+ skip] ---------
// Replaces sigmoid, mul, tanh, mul, add, relu (and the next block's `x + map_skip`) = 7 elementwise launches over a
// [P,16,16,128] activation each way, and the five intermediates autograd would keep.  The backward recomputes the gate
// from (r, g, x); the scalar UpdateSP gradient is block-reduced and added atomically.
namespace {
constexpr int kGateThreads = 256;

__device__ __forceinline__ float sigmoid_f(float v) { return 1.f / (1.f + __expf(-v)); }

__global__ void __launch_bounds__(kGateThreads)
gate_fwd_kernel(const float4* __restrict__ r, const float4* __restrict__ g, const float4* __restrict__ x,
                const float4* __restrict__ skip, const float* __restrict__ sp, float4* __restrict__ out,
                uint2* __restrict__ out_hi, uint2* __restrict__ out_lo, int64_t n4) {
    const float s = __ldg(sp);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 rv = __ldcs(r + i), gv = __ldcs(g + i), xv = __ldcs(x + i);
        float4 o;
        o.x = fmaxf(tanhf(rv.x * sigmoid_f(gv.x)) * s + xv.x, 0.f);
        o.y = fmaxf(tanhf(rv.y * sigmoid_f(gv.y)) * s + xv.y, 0.f);
        o.z = fmaxf(tanhf(rv.z * sigmoid_f(gv.z)) * s + xv.z, 0.f);
        o.w = fmaxf(tanhf(rv.w * sigmoid_f(gv.w)) * s + xv.w, 0.f);
        if (skip) { const float4 k = __ldcs(skip + i); o.x += k.x; o.y += k.y; o.z += k.z; o.w += k.w; }
        out[i] = o;
        if (out_hi) {
            const __nv_bfloat162 h0 = __floats2bfloat162_rn(o.x, o.y), h1 = __floats2bfloat162_rn(o.z, o.w);
            const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
            const __nv_bfloat162 l0 = __floats2bfloat162_rn(o.x - f0.x, o.y - f0.y);
            const __nv_bfloat162 l1 = __floats2bfloat162_rn(o.z - f1.x, o.w - f1.y);
            out_hi[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
            out_lo[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
        }
    }
}

__device__ __forceinline__ void gate_bwd_one(float d, float rv, float gv, float xv, float s, float& dr, float& dg, float& dx,
                                             float& dsp) {
    const float u = sigmoid_f(gv), t = tanhf(rv * u);
    const float dy = (t * s + xv > 0.f) ? d : 0.f;
    dx = dy;
    dsp += dy * t;
    const float dz = dy * s * (1.f - t * t);        // d/d(r*u)
    dr = dz * u;
    dg = dz * rv * u * (1.f - u);
}

__global__ void __launch_bounds__(kGateThreads)
gate_bwd_kernel(const float4* __restrict__ dout, const float4* __restrict__ r, const float4* __restrict__ g,
                const float4* __restrict__ x, const float* __restrict__ sp, float4* __restrict__ dr, float4* __restrict__ dg,
                float4* __restrict__ dx, float* __restrict__ dsp, int64_t n4) {
    __shared__ float red[kGateThreads / 32];
    const float s = __ldg(sp);
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 d = __ldcs(dout + i), rv = __ldcs(r + i), gv = __ldcs(g + i), xv = __ldcs(x + i);
        float4 a, b, c;
        gate_bwd_one(d.x, rv.x, gv.x, xv.x, s, a.x, b.x, c.x, acc);
        gate_bwd_one(d.y, rv.y, gv.y, xv.y, s, a.y, b.y, c.y, acc);
        gate_bwd_one(d.z, rv.z, gv.z, xv.z, s, a.z, b.z, c.z, acc);
        gate_bwd_one(d.w, rv.w, gv.w, xv.w, s, a.w, b.w, c.w, acc);
        dr[i] = a; dg[i] = b; dx[i] = c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < kGateThreads / 32; ++w) t += red[w];
        atomicAdd(dsp, t);
    }
}
}  // namespace

extern "C" int dsb_gate_update_fwd(const float* r, const float* g, const float* x, const float* skip, const float* sp, float* out,
                                   void* out_hi, void* out_lo, int64_t n, dsb_stream_t stream) {
    DSB_REQUIRE(r && g && x && sp && out && (!out_hi == !out_lo) && n >= 0 && n % 4 == 0, "gate_update_fwd: bad argument (n %% 4 == 0)");
    if (n == 0) return DSB_OK;
    const int64_t n4 = n / 4;
    int64_t blocks = (n4 + kGateThreads - 1) / kGateThreads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    gate_fwd_kernel<<<(unsigned)blocks, kGateThreads, 0, (cudaStream_t)stream>>>(
        (const float4*)r, (const float4*)g, (const float4*)x, (const float4*)skip, sp, (float4*)out, (uint2*)out_hi, (uint2*)out_lo, n4);
    return dsb::check_launch("gate_update_fwd");
}

extern "C" int dsb_gate_update_bwd(const float* grad_out, const float* r, const float* g, const float* x, const float* sp,
                                   float* grad_r, float* grad_g, float* grad_x, float* grad_sp, int64_t n, dsb_stream_t stream) {
    DSB_REQUIRE(grad_out && r && g && x && sp && grad_r && grad_g && grad_x && grad_sp && n >= 0 && n % 4 == 0,
                "gate_update_bwd: bad argument (n %% 4 == 0; grad_sp is accumulated into)");
    if (n == 0) return DSB_OK;
    const int64_t n4 = n / 4;
    int64_t blocks = (n4 + kGateThreads - 1) / kGateThreads;
    if (blocks > 148 * 8) blocks = 148 * 8;
    gate_bwd_kernel<<<(unsigned)blocks, kGateThreads, 0, (cudaStream_t)stream>>>(
        (const float4*)grad_out, (const float4*)r, (const float4*)g, (const float4*)x, sp, (float4*)grad_r, (float4*)grad_g,
        (float4*)grad_x, grad_sp, n4);
    return dsb::check_launch("gate_update_bwd");
}

extern "C" int dsb_colsum_pair(const void* hi, const void* lo, float* out, int64_t rows, int N, dsb_stream_t stream) {
    DSB_REQUIRE(hi && lo && out && rows >= 0 && N > 0 && N % 4 == 0, "colsum_pair: bad argument (N %% 4)");
    if (rows == 0) return DSB_OK;
    const int threads = 128;
    const int bx = (N / 4 + threads - 1) / threads;
    int by = (int)((148 * 8 + bx - 1) / bx);
    if (by > rows) by = (int)rows;
    const int rpb = (int)((rows + by - 1) / by);
    by = (int)((rows + rpb - 1) / rpb);
    colsum_pair_kernel<<<dim3(bx, by), threads, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)hi, (const __nv_bfloat16*)lo, out,
                                                                        rows, N, rpb);
    return dsb::check_launch("colsum_pair");
}
