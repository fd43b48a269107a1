// Bilinear x2 up-sampling (align_corners = False) forward / backward for the location head decoder —
// replaces F.interpolate(x, scale_factor=2., mode='bilinear') at head/action_arg_head.py:439-440.
// ATen's kernel for this op runs at ~1 % of HBM speed on the [P,128,16,16] / [P,64,32,32] / [P,32,64,64] tensors of
// the learner batch (35 ms per call at P = 1024) and its backward launches an invalid grid beyond ~2 k rows.
// Forward: one thread per output element (coalesced 4 B stores, the 4 taps hit L1).  Backward: gather form —
// one thread per INPUT element sums its <= 16 contributing output gradients in a fixed order (deterministic,
// no atomics).  Index / weight arithmetic is ATen's (area_pixel_compute_source_index): src = 0.5*(dst+0.5)-0.5
// clamped at 0, i1 = (int)src, lambda = src - i1, second tap = i1 + (i1 < size-1).
#include "common.cuh"

namespace {

struct Tap { int i0, i1; float l0, l1; };

__device__ __forceinline__ Tap tap_of(int dst, int in_size) {
    float src = 0.5f * ((float)dst + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
    const int i0 = (int)src;
    const int p = (i0 < in_size - 1) ? 1 : 0;
    const float l1 = src - (float)i0;
    return Tap{i0, i0 + p, 1.f - l1, l1};
}

__global__ void upsample2x_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t total, int H,
                                      int W) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int OW = 2 * W, OH = 2 * H;
    const int ox = (int)(i % OW);
    const int oy = (int)((i / OW) % OH);
    const int64_t nc = i / ((int64_t)OW * OH);
    const Tap ty = tap_of(oy, H), tx = tap_of(ox, W);
    const float* p = in + nc * H * W;
    const float a = __ldg(p + ty.i0 * W + tx.i0), b = __ldg(p + ty.i0 * W + tx.i1);
    const float c = __ldg(p + ty.i1 * W + tx.i0), d = __ldg(p + ty.i1 * W + tx.i1);
    out[i] = ty.l0 * (tx.l0 * a + tx.l1 * b) + ty.l1 * (tx.l0 * c + tx.l1 * d);
}

__global__ void upsample2x_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gin, int64_t total, int H,
                                      int W) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int ix = (int)(i % W);
    const int iy = (int)((i / W) % H);
    const int64_t nc = i / ((int64_t)W * H);
    const int OW = 2 * W, OH = 2 * H;
    const float* g = gout + nc * OH * OW;
    float wy[4], wx[4];
    int oys[4], oxs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int oy = 2 * iy - 1 + k;
        oys[k] = oy;
        wy[k] = 0.f;
        if (oy >= 0 && oy < OH) {
            const Tap t = tap_of(oy, H);
            wy[k] = (t.i0 == iy ? t.l0 : 0.f) + (t.i1 == iy ? t.l1 : 0.f);
        }
        const int ox = 2 * ix - 1 + k;
        oxs[k] = ox;
        wx[k] = 0.f;
        if (ox >= 0 && ox < OW) {
            const Tap t = tap_of(ox, W);
            wx[k] = (t.i0 == ix ? t.l0 : 0.f) + (t.i1 == ix ? t.l1 : 0.f);
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        if (wy[a] == 0.f) continue;
        float row = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (wx[b] != 0.f) row += wx[b] * __ldg(g + (int64_t)oys[a] * OW + oxs[b]);
        acc += wy[a] * row;
    }
    gin[i] = acc;
}

// NHWC variants (channels innermost): same taps, consecutive threads walk the channel axis -> fully coalesced.
__global__ void upsample2x_nhwc_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t total, int H,
                                           int W, int C) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int OW = 2 * W, OH = 2 * H;
    const int c = (int)(i % C);
    const int ox = (int)((i / C) % OW);
    const int oy = (int)((i / ((int64_t)C * OW)) % OH);
    const int64_t n = i / ((int64_t)C * OW * OH);
    const Tap ty = tap_of(oy, H), tx = tap_of(ox, W);
    const float* p = in + n * H * W * C + c;
    const float a = __ldg(p + ((int64_t)ty.i0 * W + tx.i0) * C), b = __ldg(p + ((int64_t)ty.i0 * W + tx.i1) * C);
    const float cc = __ldg(p + ((int64_t)ty.i1 * W + tx.i0) * C), d = __ldg(p + ((int64_t)ty.i1 * W + tx.i1) * C);
    out[i] = ty.l0 * (tx.l0 * a + tx.l1 * b) + ty.l1 * (tx.l0 * cc + tx.l1 * d);
}

__global__ void upsample2x_nhwc_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gin, int64_t total, int H,
                                           int W, int C) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C);
    const int ix = (int)((i / C) % W);
    const int iy = (int)((i / ((int64_t)C * W)) % H);
    const int64_t n = i / ((int64_t)C * W * H);
    const int OW = 2 * W, OH = 2 * H;
    const float* g = gout + n * OH * OW * C + c;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int oy = 2 * iy - 1 + a;
        if (oy < 0 || oy >= OH) continue;
        const Tap ty = tap_of(oy, H);
        const float wy = (ty.i0 == iy ? ty.l0 : 0.f) + (ty.i1 == iy ? ty.l1 : 0.f);
        if (wy == 0.f) continue;
        float row = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int ox = 2 * ix - 1 + b;
            if (ox < 0 || ox >= OW) continue;
            const Tap tx = tap_of(ox, W);
            const float wx = (tx.i0 == ix ? tx.l0 : 0.f) + (tx.i1 == ix ? tx.l1 : 0.f);
            if (wx != 0.f) row += wx * __ldg(g + ((int64_t)oy * OW + ox) * C);
        }
        acc += wy * row;
    }
    gin[i] = acc;
}

}  // namespace

extern "C" int dsb_upsample_bilinear2x_nhwc_fwd(const float* in, float* out, int64_t N, int H, int W, int C,
                                                dsb_stream_t stream) {
    DSB_REQUIRE(in && out && N >= 0 && H > 0 && W > 0 && C > 0, "upsample_bilinear2x_nhwc_fwd: bad argument");
    const int64_t total = N * 4 * H * W * C;
    if (total == 0) return DSB_OK;
    const int64_t blocks = (total + 255) / 256;
    DSB_REQUIRE(blocks < (1ll << 31), "upsample_bilinear2x_nhwc_fwd: too large");
    upsample2x_nhwc_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, total, H, W, C);
    return dsb::check_launch("upsample_bilinear2x_nhwc_fwd");
}

extern "C" int dsb_upsample_bilinear2x_nhwc_bwd(const float* grad_out, float* grad_in, int64_t N, int H, int W, int C,
                                                dsb_stream_t stream) {
    DSB_REQUIRE(grad_out && grad_in && N >= 0 && H > 0 && W > 0 && C > 0, "upsample_bilinear2x_nhwc_bwd: bad argument");
    const int64_t total = N * H * W * C;
    if (total == 0) return DSB_OK;
    const int64_t blocks = (total + 255) / 256;
    DSB_REQUIRE(blocks < (1ll << 31), "upsample_bilinear2x_nhwc_bwd: too large");
    upsample2x_nhwc_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(grad_out, grad_in, total, H, W, C);
    return dsb::check_launch("upsample_bilinear2x_nhwc_bwd");
}

extern "C" int dsb_upsample_bilinear2x_fwd(const float* in, float* out, int64_t NC, int H, int W, dsb_stream_t stream) {
    DSB_REQUIRE(in && out && NC >= 0 && H > 0 && W > 0, "upsample_bilinear2x_fwd: bad argument");
    const int64_t total = NC * 4 * H * W;
    if (total == 0) return DSB_OK;
    const int threads = 256;
    const int64_t blocks = (total + threads - 1) / threads;
    DSB_REQUIRE(blocks < (1ll << 31), "upsample_bilinear2x_fwd: too large");
    upsample2x_fwd_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(in, out, total, H, W);
    return dsb::check_launch("upsample_bilinear2x_fwd");
}

extern "C" int dsb_upsample_bilinear2x_bwd(const float* grad_out, float* grad_in, int64_t NC, int H, int W,
                                           dsb_stream_t stream) {
    DSB_REQUIRE(grad_out && grad_in && NC >= 0 && H > 0 && W > 0, "upsample_bilinear2x_bwd: bad argument");
    const int64_t total = NC * H * W;
    if (total == 0) return DSB_OK;
    const int threads = 256;
    const int64_t blocks = (total + threads - 1) / threads;
    DSB_REQUIRE(blocks < (1ll << 31), "upsample_bilinear2x_bwd: too large");
    upsample2x_bwd_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(grad_out, grad_in, total, H, W);
    return dsb::check_launch("upsample_bilinear2x_bwd");
}

// ---- location-head tail: logits = conv3x3( upsample2x(x), w[1,C,3,3] ) + b  with ONE output channel ----------------
// (head/action_arg_head.py:436-443, last `upsample` stage).  Up-sampling is linear and channel independent, so
//     out(p) = b + sum_tap up(z_tap)(p + tap - 1),   z_tap = sum_c w[c,tap] * x[..,c]   (a 32 -> 9 projection at LOW resolution)
// which needs 4x fewer multiply-adds than convolving at 128x128 and never materialises the [P,128,128,32] tensor.
// This kernel is the second half: z [N, h, w, 9] -> out [N, 2h, 2w] (positions outside the up-sampled image count as
// the conv's zero padding).  Backward is the exact transpose as a gather over <= 16 output pixels per z element.
namespace {

__device__ __forceinline__ float up_at(const float* __restrict__ z, int H, int W, int uy, int ux, int tap) {
    if (uy < 0 || uy >= 2 * H || ux < 0 || ux >= 2 * W) return 0.f;      // zero padding of the 3x3 conv
    const Tap ty = tap_of(uy, H), tx = tap_of(ux, W);
    const float a = __ldg(z + ((int64_t)ty.i0 * W + tx.i0) * 9 + tap), b = __ldg(z + ((int64_t)ty.i0 * W + tx.i1) * 9 + tap);
    const float c = __ldg(z + ((int64_t)ty.i1 * W + tx.i0) * 9 + tap), d = __ldg(z + ((int64_t)ty.i1 * W + tx.i1) * 9 + tap);
    return ty.l0 * (tx.l0 * a + tx.l1 * b) + ty.l1 * (tx.l0 * c + tx.l1 * d);
}

__global__ void upshift9_fwd_kernel(const float* __restrict__ z, const float* __restrict__ bias, float* __restrict__ out,
                                    int64_t total, int H, int W) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int OW = 2 * W, OH = 2 * H;
    const int ox = (int)(i % OW), oy = (int)((i / OW) % OH);
    const int64_t n = i / ((int64_t)OW * OH);
    const float* zn = z + n * H * W * 9;
    float acc = bias ? __ldg(bias) : 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc += up_at(zn, H, W, oy + t / 3 - 1, ox + t % 3 - 1, t);
    out[i] = acc;
}

// dz[n, iy, ix, tap] = sum over up-sampled positions u that read input (iy, ix):  w_u * dOut[u - (tap offset)]
// One thread owns a low-resolution pixel and all nine taps: the 6x6 window of dOut they share is read once (36 loads for 9
// outputs instead of 16 each) and the nine results leave as one contiguous 36-byte run.
__global__ void upshift9_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gz, int64_t total, int H, int W) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int ix = (int)(i % W), iy = (int)((i / W) % H);
    const int64_t n = i / ((int64_t)W * H);
    const int OW = 2 * W, OH = 2 * H;
    const float* g = gout + n * OH * OW;
    float wy[4], wx[4];
#pragma unroll
    for (int ui = 0; ui < 4; ++ui) {
        const int uy = 2 * iy - 1 + ui, ux = 2 * ix - 1 + ui;
        wy[ui] = wx[ui] = 0.f;
        if (uy >= 0 && uy < OH) { const Tap t = tap_of(uy, H); wy[ui] = (t.i0 == iy ? t.l0 : 0.f) + (t.i1 == iy ? t.l1 : 0.f); }
        if (ux >= 0 && ux < OW) { const Tap t = tap_of(ux, W); wx[ui] = (t.i0 == ix ? t.l0 : 0.f) + (t.i1 == ix ? t.l1 : 0.f); }
    }
    float acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.f;
#pragma unroll
    for (int oi = 0; oi < 6; ++oi) {                 // output row o = 2 iy - 2 + oi meets (ui, d) with ui = oi - 1 + d
        const int oy = 2 * iy - 2 + oi;
        if (oy < 0 || oy >= OH) continue;
        float T[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int oj = 0; oj < 6; ++oj) {
            const int ox = 2 * ix - 2 + oj;
            if (ox < 0 || ox >= OW) continue;
            const float v = __ldg(g + (int64_t)oy * OW + ox);
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const int ui = oj - 1 + (d - 1);
                if (ui >= 0 && ui <= 3) T[d] += wx[ui] * v;
            }
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int ui = oi - 1 + (d - 1);
            if (ui < 0 || ui > 3) continue;
#pragma unroll
            for (int e = 0; e < 3; ++e) acc[d * 3 + e] += wy[ui] * T[e];
        }
    }
    float* out = gz + i * 9;
#pragma unroll
    for (int t = 0; t < 9; ++t) out[t] = acc[t];
}

// ---- location-head up-sampling stages with C output channels: y = act(conv3x3(upsample2x(x), w) + b) --------------------
// (head/action_arg_head.py:436-443, `upsample` stages 0 and 1).  Same factorisation as upshift9, generalised to C output
// channels: z[p, tap, co] = sum_ci w[co, ci, tap] x[p, ci] is ONE low-resolution GEMM on the tensor cores (K = Cin,
// N = 9*C), and y(o) = b + sum_tap up(z[., tap, .])(o + tap - 1) is this kernel.  Against convolving the up-sampled image
// it needs 4x fewer tensor-core flops and never materialises (or keeps for backward) the 4x larger up-sampled activation.
// One thread owns a low-resolution cell (iy, ix) x 4 channels = a 2x2 block of outputs, so the 3x3 low-resolution
// neighbourhood of every tap is loaded once for four outputs (81 16-byte loads instead of 144).
struct AxisW { float w[4][3]; };   // [u - (2i - 1)][low-res neighbour j = row - (i - 1)] bilinear weight (0 outside the image)

__device__ __forceinline__ AxisW axis_weights(int i, int size) {
    AxisW a;
#pragma unroll
    for (int ui = 0; ui < 4; ++ui) {
        a.w[ui][0] = a.w[ui][1] = a.w[ui][2] = 0.f;
        const int u = 2 * i - 1 + ui;
        if (u < 0 || u >= 2 * size) continue;               // the conv's zero padding
        const Tap t = tap_of(u, size);
        const int j0 = t.i0 - (i - 1), j1 = t.i1 - (i - 1);
#pragma unroll
        for (int j = 0; j < 3; ++j) a.w[ui][j] += (j == j0 ? t.l0 : 0.f) + (j == j1 ? t.l1 : 0.f);
    }
    return a;
}

template <int C>
__global__ void __launch_bounds__(256)
upconv_fwd_kernel(const float* __restrict__ z, int ldz, const float* __restrict__ bias, int relu, float* __restrict__ out,
                  __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, int ldo, int H, int W) {
    constexpr int Q = C / 4, TW = 8, TH = 512 / (Q * TW);         // 8x8 cells (C = 32) or 4x8 (C = 64) per block
    const int tiles_x = W / TW, tiles_y = H / TH;
    const int tile = blockIdx.x % (tiles_x * tiles_y);
    const int64_t n = blockIdx.x / (tiles_x * tiles_y);
    const int ty0 = (tile / tiles_x) * TH, tx0 = (tile % tiles_x) * TW;
    const float* zn = z + n * H * W * (int64_t)ldz;
    for (int idx = threadIdx.x; idx < TH * TW * Q; idx += 256) {
        const int q = idx % Q, cell = idx / Q;
        const int iy = ty0 + cell / TW, ix = tx0 + cell % TW;
        const AxisW wy = axis_weights(iy, H), wx = axis_weights(ix, W);
        float4 acc[2][2];
        const float4 b4 = bias ? __ldg(reinterpret_cast<const float4*>(bias) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        acc[0][0] = acc[0][1] = acc[1][0] = acc[1][1] = b4;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int dy = t / 3, dx = t % 3;                       // tap offset + 1
            float4 R[2][3];                                         // R[b][j] = sum_k wx[b + dx][k] * Z[j][k]
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                R[0][j] = R[1][j] = make_float4(0.f, 0.f, 0.f, 0.f);
                const int ry = iy - 1 + j;
                if (ry < 0 || ry >= H) continue;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int rx = ix - 1 + k;
                    if (rx < 0 || rx >= W) continue;
                    const float4 v = __ldg(reinterpret_cast<const float4*>(zn + ((int64_t)ry * W + rx) * ldz + t * C) + q);
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        const float w = wx.w[b + dx][k];
                        R[b][j].x += w * v.x; R[b][j].y += w * v.y; R[b][j].z += w * v.z; R[b][j].w += w * v.w;
                    }
                }
            }
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b)
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const float w = wy.w[a + dy][j];
                        acc[a][b].x += w * R[b][j].x; acc[a][b].y += w * R[b][j].y;
                        acc[a][b].z += w * R[b][j].z; acc[a][b].w += w * R[b][j].w;
                    }
        }
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                float4 v = acc[a][b];
                if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                const int64_t off = ((n * 2 * H + 2 * iy + a) * (int64_t)(2 * W) + 2 * ix + b) * ldo + 4 * q;
                if (out) *reinterpret_cast<float4*>(out + off) = v;
                if (out_hi) {
                    const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
                    const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
                    const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y);
                    const __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
                    *reinterpret_cast<uint2*>(out_hi + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
                    *reinterpret_cast<uint2*>(out_lo + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
                }
            }
    }
}

// Transpose of the above: gz[cell, tap, c] = sum over the (<= 4x4) up-sampled positions u that read cell with bilinear weight
// w(u):  w(u) * g[u - tap offset, c],  g = dL/dy already ReLU-masked.  One thread owns (cell, 4 channels) and all nine taps:
// the 6x6 window of g it needs is streamed row by row (36 16-byte loads for 9 outputs).  gz is written as the bf16
// (hi, lo) pair the dX / dW tensor-core GEMMs read; columns [9C, ldz) are zero filled.
template <int C>
__global__ void __launch_bounds__(256)
upconv_bwd_kernel(const float* __restrict__ g, int ldg, __nv_bfloat16* __restrict__ gz_hi, __nv_bfloat16* __restrict__ gz_lo,
                  int ldz, int H, int W) {
    constexpr int Q = C / 4, TW = 8, TH = 512 / (Q * TW);
    const int tiles_x = W / TW, tiles_y = H / TH;
    const int tile = blockIdx.x % (tiles_x * tiles_y);
    const int64_t n = blockIdx.x / (tiles_x * tiles_y);
    const int ty0 = (tile / tiles_x) * TH, tx0 = (tile % tiles_x) * TW;
    const int OH = 2 * H, OW = 2 * W;
    const float* gn = g + n * OH * OW * (int64_t)ldg;
    for (int idx = threadIdx.x; idx < TH * TW * Q; idx += 256) {
        const int q = idx % Q, cell = idx / Q;
        const int iy = ty0 + cell / TW, ix = tx0 + cell % TW;
        // weight of up-sampled position u = 2i - 1 + ui on this cell (0 when u is outside the image)
        float wy[4], wx[4];
#pragma unroll
        for (int ui = 0; ui < 4; ++ui) {
            const int uy = 2 * iy - 1 + ui, ux = 2 * ix - 1 + ui;
            wy[ui] = wx[ui] = 0.f;
            if (uy >= 0 && uy < OH) { const Tap t = tap_of(uy, H); wy[ui] = (t.i0 == iy ? t.l0 : 0.f) + (t.i1 == iy ? t.l1 : 0.f); }
            if (ux >= 0 && ux < OW) { const Tap t = tap_of(ux, W); wx[ui] = (t.i0 == ix ? t.l0 : 0.f) + (t.i1 == ix ? t.l1 : 0.f); }
        }
        float4 acc[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        // output pixel o = u - d, d in {-1,0,1}: o ranges over 2i - 2 .. 2i + 3 (index oi = o - (2i - 2) in 0..5); the pair
        // (ui, d) that meets o satisfies ui = oi - 1 + d  (d = tap offset)
#pragma unroll
        for (int oi = 0; oi < 6; ++oi) {
            const int oy = 2 * iy - 2 + oi;
            if (oy < 0 || oy >= OH) continue;
            float4 T[3];                                 // T[dx + 1] = sum_ui wx[ui] * g[oy, ux - dx]
            T[0] = T[1] = T[2] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int oj = 0; oj < 6; ++oj) {
                const int ox = 2 * ix - 2 + oj;
                if (ox < 0 || ox >= OW) continue;
                const float4 v = __ldg(reinterpret_cast<const float4*>(gn + ((int64_t)oy * OW + ox) * ldg) + q);
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const int ui = oj - 1 + (d - 1);     // ux = ox + dx
                    if (ui < 0 || ui > 3) continue;
                    const float w = wx[ui];
                    T[d].x += w * v.x; T[d].y += w * v.y; T[d].z += w * v.z; T[d].w += w * v.w;
                }
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const int ui = oi - 1 + (d - 1);
                if (ui < 0 || ui > 3) continue;
                const float w = wy[ui];
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    float4& a = acc[d * 3 + e];
                    a.x += w * T[e].x; a.y += w * T[e].y; a.z += w * T[e].z; a.w += w * T[e].w;
                }
            }
        }
        const int64_t row = ((n * H + iy) * (int64_t)W + ix) * ldz;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const float4 v = acc[t];
            const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
            const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
            const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y);
            const __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
            const int64_t off = row + t * C + 4 * q;
            *reinterpret_cast<uint2*>(gz_hi + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
            *reinterpret_cast<uint2*>(gz_lo + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
        }
        for (int c = 9 * C + 4 * q; c < ldz; c += C) {
            *reinterpret_cast<uint2*>(gz_hi + row + c) = make_uint2(0u, 0u);
            *reinterpret_cast<uint2*>(gz_lo + row + c) = make_uint2(0u, 0u);
        }
    }
}

}  // namespace

extern "C" int dsb_upconv_fwd(const float* z, int ldz, const float* bias, int relu, float* out, void* out_hi, void* out_lo,
                              int ldo, int64_t N, int H, int W, int C, dsb_stream_t stream) {
    DSB_REQUIRE(z && (out || out_hi) && (!out_hi == !out_lo) && N >= 0 && (C == 32 || C == 64) && ldz >= 9 * C && ldz % 4 == 0 &&
                ldo >= C && ldo % 4 == 0 && W % 8 == 0 && H % 8 == 0, "upconv_fwd: bad argument (C in {32,64}, H, W %% 8 == 0)");
    if (N == 0) return DSB_OK;
    const int64_t blocks = N * (W / 8) * (H / (C == 32 ? 8 : 4));
    DSB_REQUIRE(blocks < (1ll << 31), "upconv_fwd: too large");
    if (C == 32)
        upconv_fwd_kernel<32><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(z, ldz, bias, relu, out, (__nv_bfloat16*)out_hi,
                                                                                  (__nv_bfloat16*)out_lo, ldo, H, W);
    else
        upconv_fwd_kernel<64><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(z, ldz, bias, relu, out, (__nv_bfloat16*)out_hi,
                                                                                  (__nv_bfloat16*)out_lo, ldo, H, W);
    return dsb::check_launch("upconv_fwd");
}

extern "C" int dsb_upconv_bwd(const float* g, int ldg, void* gz_hi, void* gz_lo, int ldz, int64_t N, int H, int W, int C,
                              dsb_stream_t stream) {
    DSB_REQUIRE(g && gz_hi && gz_lo && N >= 0 && (C == 32 || C == 64) && ldz >= 9 * C && ldz % 4 == 0 && (ldz - 9 * C) % C == 0 &&
                ldg >= C && ldg % 4 == 0 && W % 8 == 0 && H % 8 == 0, "upconv_bwd: bad argument");
    if (N == 0) return DSB_OK;
    const int64_t blocks = N * (W / 8) * (H / (C == 32 ? 8 : 4));
    DSB_REQUIRE(blocks < (1ll << 31), "upconv_bwd: too large");
    if (C == 32)
        upconv_bwd_kernel<32><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g, ldg, (__nv_bfloat16*)gz_hi,
                                                                                  (__nv_bfloat16*)gz_lo, ldz, H, W);
    else
        upconv_bwd_kernel<64><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g, ldg, (__nv_bfloat16*)gz_hi,
                                                                                  (__nv_bfloat16*)gz_lo, ldz, H, W);
    return dsb::check_launch("upconv_bwd");
}


extern "C" int dsb_upshift9_fwd(const float* z, const float* bias, float* out, int64_t N, int H, int W, dsb_stream_t stream) {
    DSB_REQUIRE(z && out && N >= 0 && H > 0 && W > 0, "upshift9_fwd: bad argument");
    const int64_t total = N * 4 * H * W;
    if (total == 0) return DSB_OK;
    const int64_t blocks = (total + 255) / 256;
    DSB_REQUIRE(blocks < (1ll << 31), "upshift9_fwd: too large");
    upshift9_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(z, bias, out, total, H, W);
    return dsb::check_launch("upshift9_fwd");
}

extern "C" int dsb_upshift9_bwd(const float* grad_out, float* grad_z, int64_t N, int H, int W, dsb_stream_t stream) {
    DSB_REQUIRE(grad_out && grad_z && N >= 0 && H > 0 && W > 0, "upshift9_bwd: bad argument");
    const int64_t total = N * H * W;
    if (total == 0) return DSB_OK;
    const int64_t blocks = (total + 255) / 256;
    DSB_REQUIRE(blocks < (1ll << 31), "upshift9_bwd: too large");
    upshift9_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(grad_out, grad_z, total, H, W);
    return dsb::check_launch("upshift9_bwd");
}


// ---- 2x2 / stride-2 max-pool on channels-last activations (spatial encoder, between the down-sampling convolutions) ------
// out[n, y, x, c] = max over the 2x2 window (first maximum in scan order, as ATen); also writes the bf16 (hi, lo) pair the
// next convolution reads and a 2-bit argmax per element (one byte) so the backward does not re-read the input.
namespace {
__global__ void __launch_bounds__(256)
maxpool2_fwd_kernel(const float4* __restrict__ x, float4* __restrict__ out, uint2* __restrict__ out_hi, uint2* __restrict__ out_lo,
                    uchar4* __restrict__ idx, int64_t total, int H, int W, int C4) {
    const int OH = H / 2, OW = W / 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        const int ox = (int)((i / C4) % OW), oy = (int)((i / ((int64_t)C4 * OW)) % OH);
        const int64_t n = i / ((int64_t)C4 * OW * OH);
        const float4* base = x + ((n * H + 2 * oy) * (int64_t)W + 2 * ox) * C4 + c;
        const float4 v[4] = {__ldcs(base), __ldcs(base + C4), __ldcs(base + (int64_t)W * C4), __ldcs(base + (int64_t)W * C4 + C4)};
        float4 m = v[0];
        uchar4 k = make_uchar4(0, 0, 0, 0);
#pragma unroll
        for (int j = 1; j < 4; ++j) {
            if (v[j].x > m.x) { m.x = v[j].x; k.x = j; }
            if (v[j].y > m.y) { m.y = v[j].y; k.y = j; }
            if (v[j].z > m.z) { m.z = v[j].z; k.z = j; }
            if (v[j].w > m.w) { m.w = v[j].w; k.w = j; }
        }
        out[i] = m;
        idx[i] = k;
        if (out_hi) {
            const __nv_bfloat162 h0 = __floats2bfloat162_rn(m.x, m.y), h1 = __floats2bfloat162_rn(m.z, m.w);
            const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
            const __nv_bfloat162 l0 = __floats2bfloat162_rn(m.x - f0.x, m.y - f0.y);
            const __nv_bfloat162 l1 = __floats2bfloat162_rn(m.z - f1.x, m.w - f1.y);
            out_hi[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
            out_lo[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
        }
    }
}

__global__ void __launch_bounds__(256)
maxpool2_bwd_kernel(const float4* __restrict__ gout, const uchar4* __restrict__ idx, float4* __restrict__ gx, int64_t total,
                    int H, int W, int C4) {
    const int OH = H / 2, OW = W / 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        const int ox = (int)((i / C4) % OW), oy = (int)((i / ((int64_t)C4 * OW)) % OH);
        const int64_t n = i / ((int64_t)C4 * OW * OH);
        const float4 g = __ldcs(gout + i);
        const uchar4 k = idx[i];
        float4* base = gx + ((n * H + 2 * oy) * (int64_t)W + 2 * ox) * C4 + c;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float4 o;
            o.x = (k.x == j) ? g.x : 0.f; o.y = (k.y == j) ? g.y : 0.f;
            o.z = (k.z == j) ? g.z : 0.f; o.w = (k.w == j) ? g.w : 0.f;
            base[(j & 1) * C4 + (j >> 1) * (int64_t)W * C4] = o;
        }
    }
}
}  // namespace

extern "C" int dsb_maxpool2_nhwc_fwd(const float* x, float* out, void* out_hi, void* out_lo, uint8_t* argmax, int64_t N, int H,
                                     int W, int C, dsb_stream_t stream) {
    DSB_REQUIRE(x && out && argmax && (!out_hi == !out_lo) && N >= 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C % 4 == 0,
                "maxpool2_nhwc_fwd: bad argument (even H, W; C %% 4 == 0)");
    const int64_t total = N * (H / 2) * (W / 2) * (C / 4);
    if (total == 0) return DSB_OK;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    maxpool2_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)x, (float4*)out, (uint2*)out_hi,
                                                                            (uint2*)out_lo, (uchar4*)argmax, total, H, W, C / 4);
    return dsb::check_launch("maxpool2_nhwc_fwd");
}

extern "C" int dsb_maxpool2_nhwc_bwd(const float* grad_out, const uint8_t* argmax, float* grad_x, int64_t N, int H, int W, int C,
                                     dsb_stream_t stream) {
    DSB_REQUIRE(grad_out && argmax && grad_x && N >= 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C % 4 == 0,
                "maxpool2_nhwc_bwd: bad argument");
    const int64_t total = N * (H / 2) * (W / 2) * (C / 4);
    if (total == 0) return DSB_OK;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    maxpool2_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)grad_out, (const uchar4*)argmax,
                                                                            (float4*)grad_x, total, H, W, C / 4);
    return dsb::check_launch("maxpool2_nhwc_bwd");
}


// ---- location-head tail projection z[p, t] = sum_c x[p, c] w[c, t]  (32 channels -> the 9 taps of the last 3x3 conv) -------
// A [16.7 M x 32] x [32 x 9] product: far too skinny for the library GEMMs (the weight gradient ran as a 4.5 ms "large-k"
// kernel).  Forward: one thread per pixel.  Backward: one warp walks pixels with lane = channel; every lane keeps its 9
// weight-gradient accumulators in registers and writes its channel of dL/dx; partials are reduced per block and added
// atomically into gw [32, 9].
namespace {
constexpr int kP9C = 32, kP9T = 9;

__global__ void __launch_bounds__(256)
proj9_fwd_kernel(const float4* __restrict__ x, const float* __restrict__ w, float* __restrict__ z, int64_t pixels) {
    __shared__ float ws[kP9C * kP9T];
    for (int i = threadIdx.x; i < kP9C * kP9T; i += blockDim.x) ws[i] = w[i];
    __syncthreads();
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < pixels; p += (int64_t)gridDim.x * blockDim.x) {
        float acc[kP9T];
#pragma unroll
        for (int t = 0; t < kP9T; ++t) acc[t] = 0.f;
#pragma unroll
        for (int j = 0; j < kP9C / 4; ++j) {
            const float4 v = __ldcs(x + p * (kP9C / 4) + j);
#pragma unroll
            for (int t = 0; t < kP9T; ++t)
                acc[t] += v.x * ws[(4 * j) * kP9T + t] + v.y * ws[(4 * j + 1) * kP9T + t] + v.z * ws[(4 * j + 2) * kP9T + t] +
                          v.w * ws[(4 * j + 3) * kP9T + t];
        }
#pragma unroll
        for (int t = 0; t < kP9T; ++t) z[p * kP9T + t] = acc[t];
    }
}

__global__ void __launch_bounds__(256)
proj9_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gz, const float* __restrict__ w, float* __restrict__ gx,
                 float* __restrict__ gw, int64_t pixels) {
    __shared__ float red[8][kP9C * kP9T];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float wl[kP9T], acc[kP9T];
#pragma unroll
    for (int t = 0; t < kP9T; ++t) { wl[t] = __ldg(w + lane * kP9T + t); acc[t] = 0.f; }
    const int64_t warps = (int64_t)gridDim.x * 8;
    for (int64_t p = (int64_t)blockIdx.x * 8 + warp; p < pixels; p += warps) {
        const float xv = __ldcs(x + p * kP9C + lane);
        float g[kP9T];
#pragma unroll
        for (int t = 0; t < kP9T; ++t) g[t] = __ldg(gz + p * kP9T + t);       // same address in every lane: one broadcast load
        float d = 0.f;
#pragma unroll
        for (int t = 0; t < kP9T; ++t) { d += g[t] * wl[t]; acc[t] += xv * g[t]; }
        gx[p * kP9C + lane] = d;
    }
#pragma unroll
    for (int t = 0; t < kP9T; ++t) red[warp][lane * kP9T + t] = acc[t];
    __syncthreads();
    for (int i = threadIdx.x; i < kP9C * kP9T; i += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += red[k][i];
        atomicAdd(gw + i, s);
    }
}
}  // namespace

extern "C" int dsb_proj9_fwd(const float* x, const float* w, float* z, int64_t pixels, int C, dsb_stream_t stream) {
    DSB_REQUIRE(x && w && z && pixels >= 0 && C == kP9C, "proj9_fwd: bad argument (C must be 32)");
    if (pixels == 0) return DSB_OK;
    int64_t blocks = (pixels + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    proj9_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)x, w, z, pixels);
    return dsb::check_launch("proj9_fwd");
}

extern "C" int dsb_proj9_bwd(const float* x, const float* grad_z, const float* w, float* grad_x, float* grad_w, int64_t pixels,
                             int C, dsb_stream_t stream) {
    DSB_REQUIRE(x && grad_z && w && grad_x && grad_w && pixels >= 0 && C == kP9C, "proj9_bwd: bad argument (C must be 32)");
    if (pixels == 0) return DSB_OK;
    int64_t blocks = (pixels + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    proj9_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, grad_z, w, grad_x, grad_w, pixels);
    return dsb::check_launch("proj9_bwd");
}
