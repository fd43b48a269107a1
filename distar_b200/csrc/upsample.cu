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
__global__ void upshift9_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gz, int64_t total, int H, int W) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int t = (int)(i % 9);
    const int ix = (int)((i / 9) % W), iy = (int)((i / (9 * (int64_t)W)) % H);
    const int64_t n = i / (9 * (int64_t)W * H);
    const int OW = 2 * W, OH = 2 * H;
    const int dy = t / 3 - 1, dx = t % 3 - 1;
    const float* g = gout + n * OH * OW;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int uy = 2 * iy - 1 + a;
        if (uy < 0 || uy >= OH) continue;
        const Tap ty = tap_of(uy, H);
        const float wy = (ty.i0 == iy ? ty.l0 : 0.f) + (ty.i1 == iy ? ty.l1 : 0.f);
        const int oy = uy - dy;                       // output pixel whose tap (dy,dx) reads up-sampled position uy
        if (wy == 0.f || oy < 0 || oy >= OH) continue;
        float row = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int ux = 2 * ix - 1 + b;
            if (ux < 0 || ux >= OW) continue;
            const Tap tx = tap_of(ux, W);
            const float wx = (tx.i0 == ix ? tx.l0 : 0.f) + (tx.i1 == ix ? tx.l1 : 0.f);
            const int ox = ux - dx;
            if (wx != 0.f && ox >= 0 && ox < OW) row += wx * __ldg(g + (int64_t)oy * OW + ox);
        }
        acc += wy * row;
    }
    gz[i] = acc;
}

}  // namespace

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
    const int64_t total = N * H * W * 9;
    if (total == 0) return DSB_OK;
    const int64_t blocks = (total + 255) / 256;
    DSB_REQUIRE(blocks < (1ll << 31), "upshift9_bwd: too large");
    upshift9_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(grad_out, grad_z, total, H, W);
    return dsb::check_launch("upshift9_bwd");
}
