// Fused spatial-encoder stem: plane expansion + entity scatter + 1x1 conv (56 -> 32) + ReLU + 2x2 max-pool.
//
// Replaces, for the training/inference path, the chain  scatter_connection (module_utils.py:11-34) -> 6 one-hot
// embeddings + 6 effect-index scatters + cat (spatial_encoder.py:51-71) -> project conv + ReLU (:72) -> first
// max_pool2d (:75-79).  The 56-channel fp32 input (3.7 MB / obs) and the 2 MiB scatter map are never materialised:
//   pre[p, o] = b[o] + W[o,0] * height/256 + sum_k W[o, base_k + plane_k[p]]        (one-hot planes: ONE combined-table lookup)
//             + sum_j [p in effect_j] W[o, 18+j]                                     (zero padding makes pixel 0 always set)
//             + sum_{entities e at p} sum_c W[o, 24+c] * project[e, c]               (the scatter, pushed through the conv)
//   out[pooled p, o] = max over the 2x2 window of relu(pre)
// A CTA owns (obs, 4 input rows): lane = output channel, warps walk pixels; the pre-activation tile (512 px x 32 ch)
// lives in shared memory.  HBM traffic per obs: 7 plane bytes/pixel + 64 KiB of entity rows in, 32 B per POOLED pixel
// out (fp32 NHWC padded to 64 channels, plus the bf16 (hi, lo) pair the following 3x3 conv reads).
// Backward recomputes the tile, routes dOut through the max/ReLU decisions and produces dW [32,56], db [32] and
// d_project [N,E,32] directly (no dense gradient map).
#include "common.cuh"

namespace {

constexpr int kOC = 32;            // output channels of the project conv
constexpr int kIC = 56;            // input planes
constexpr int kRows = 4;           // input rows per tile
constexpr int kThreads = 512;          // 2 CTAs (2 x 93 KB of shared memory) x 16 warps per SM: the tile phases are latency bound
constexpr int kWarps = kThreads / 32;
constexpr int kPlanes = 7;         // height, visibility, creep, player_relative, alerts, pathable, buildable
constexpr int kEffects = 6;
constexpr int kEffLen = 100;

__device__ __constant__ int kPlaneBase[kPlanes] = {0, 1, 5, 7, 12, 14, 16};
__device__ __constant__ int kPlaneVocab[kPlanes] = {0, 4, 2, 5, 2, 2, 2};
// The six categorical planes of a pixel select one of 4*2*5*2*2*2 = 320 combinations; the sum of their six weight columns is
// looked up in ONE table row (lut [320][32], rebuilt from the current weights by stem_lut_kernel at every launch) instead of six
// shared-memory lookups + adds per pixel and lane: the stem is instruction-issue bound (ncu: 66 % issue-slot utilisation,
// 129 warp instructions per pixel before this change), not bandwidth bound.
constexpr int kCombos = 320;
__host__ __device__ constexpr int combo_stride(int k) {   // k = 1..6 -> multiplier of plane k's digit
    return k == 1 ? 1 : (k == 2 ? 4 : (k == 3 ? 8 : (k == 4 ? 40 : (k == 5 ? 80 : 160))));
}
__host__ __device__ constexpr int vocab_c(int k) { return k == 1 ? 4 : (k == 3 ? 5 : (k == 0 ? 0 : 2)); }   // compile-time vocab

struct StemArgs {
    const uint8_t* planes[kPlanes];     // each [N, H, W]
    const int16_t* effects[kEffects];   // each [N, 100]
    const float* project;               // [N, E, 32]  (already masked / ReLU'd scatter_project output)
    const uint8_t* ex;
    const uint8_t* ey;
    const int64_t* entity_num;
    const float* weight;                // [32, 56]
    const float* bias;                  // [32]
    const float* lut;                   // [320, 32] combined categorical table (workspace, written by stem_lut_kernel)
    int N, E, H, W;
};

struct Smem {
    float* pre;        // [npix][32]
    float* wt;         // [56][32]  (transposed weight)
    float* wo;         // [32][56]  (original layout, for the d_project product)
    uint16_t* cidx;    // [npix] combined index of the six categorical planes
    uint8_t* hgt;      // [npix] height_map
    uint32_t* eff;     // [npix/4] one byte per pixel: bit j = pixel is in effect list j
    uint32_t* list;    // [E] (e << 16) | pix
    int* counters;     // [kWarps + 1]
};

__device__ __forceinline__ Smem carve(unsigned char* raw, int npix, int E) {
    Smem s;
    s.pre = reinterpret_cast<float*>(raw);
    s.wt = s.pre + npix * kOC;
    s.wo = s.wt + kIC * kOC;
    s.cidx = reinterpret_cast<uint16_t*>(s.wo + kIC * kOC);
    s.hgt = reinterpret_cast<uint8_t*>(s.cidx + npix);
    s.eff = reinterpret_cast<uint32_t*>(s.hgt + npix);
    s.list = s.eff + npix / 4;
    s.counters = reinterpret_cast<int*>(s.list + E);
    return s;
}

__host__ __device__ inline size_t stem_smem_bytes(int npix, int E) {
    return (size_t)npix * kOC * 4 + (size_t)2 * kIC * kOC * 4 + (size_t)npix * 2 + (size_t)npix + (size_t)npix +
           (size_t)E * 4 + (kWarps + 1) * 4;
}

// Builds the pre-activation tile of (obs n, input rows [y0, y0+kRows)) in shared memory; returns the number of listed
// entities (s.list holds them in entity order).  All threads participate.
__device__ __forceinline__ int build_pre_tile(const StemArgs& a, const Smem& s, int n, int y0) {
    const int W = a.W, H = a.H, E = a.E, npix = kRows * W;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // stage this tile: the six categorical planes collapse into one combined index per pixel (ids clamped into their
    // vocabulary as before), the height byte is kept, the effect bytes are cleared
    {
        const size_t base = ((size_t)n * H + y0) * W;
        for (int i = tid; i < npix; i += kThreads) {
            int c = 0;
#pragma unroll
            for (int k = 1; k < kPlanes; ++k) c += min((int)a.planes[k][base + i], vocab_c(k) - 1) * combo_stride(k);
            s.cidx[i] = (uint16_t)c;
            s.hgt[i] = a.planes[0][base + i];
        }
    }
    for (int i = tid; i < npix / 4; i += kThreads) s.eff[i] = 0u;
    if (tid == 0) s.counters[kWarps] = 0;
    __syncthreads();
    // effect lists: every entry (including the zero padding) lights its pixel
    for (int i = tid; i < kEffects * kEffLen; i += kThreads) {
        const int j = i / kEffLen;
        const int idx = (int)a.effects[j][(size_t)n * kEffLen + (i - j * kEffLen)];
        const int p = idx - y0 * W;
        if (p >= 0 && p < npix) atomicOr(&s.eff[p >> 2], 1u << ((p & 3) * 8 + j));
    }
    // ordered list of the entities inside this tile (same ballot compaction as scatter_connection)
    const int en = a.entity_num ? min((int)a.entity_num[n], E) : E;
    for (int base = 0; base < en; base += kThreads) {
        const int e = base + tid;
        int pix = -1;
        if (e < en) {
            const int yy = min((int)a.ey[(size_t)n * E + e], H - 1), xx = min((int)a.ex[(size_t)n * E + e], W - 1);
            if (yy >= y0 && yy < y0 + kRows) pix = (yy - y0) * W + xx;
        }
        const unsigned m = __ballot_sync(0xffffffffu, pix >= 0);
        if (lane == 0) s.counters[warp] = __popc(m);
        __syncthreads();
        int off = s.counters[kWarps];
        for (int w = 0; w < warp; ++w) off += s.counters[w];
        if (pix >= 0) s.list[off + __popc(m & ((1u << lane) - 1))] = ((uint32_t)e << 16) | (uint32_t)pix;
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < kWarps; ++w) tot += s.counters[w];
            s.counters[kWarps] += tot;
        }
        __syncthreads();
    }
    // dense part: lane = output channel, each warp walks its pixels
    const float b = a.bias[lane], wh = s.wt[lane] * (1.0f / 256.0f);
    const uint8_t* effb = reinterpret_cast<const uint8_t*>(s.eff);
    for (int p = warp; p < npix; p += kWarps) {
        float v = fmaf(wh, (float)s.hgt[p], b) + __ldg(a.lut + (int)s.cidx[p] * kOC + lane);
        unsigned e = effb[p];                                  // almost always 0: <= 600 list entries per observation
        while (e) {
            const int j = __ffs(e) - 1;
            v += s.wt[(18 + j) * kOC + lane];
            e &= e - 1;
        }
        s.pre[p * kOC + lane] = v;
    }
    __syncthreads();
    // sparse part: entities, in entity order; a pixel is owned by warp (pix % kWarps)
    const int len = s.counters[kWarps];
    const float* prow = a.project + (size_t)n * E * kOC;
    for (int i = 0; i < len; ++i) {
        const uint32_t ent = s.list[i];
        const int pix = ent & 0xffff;
        if ((pix & (kWarps - 1)) != warp) continue;
        const float pe = __ldg(prow + (size_t)(ent >> 16) * kOC + lane);      // project[e, c = lane]
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < kOC; ++c) acc += s.wt[(24 + c) * kOC + lane] * __shfl_sync(0xffffffffu, pe, c);
        s.pre[pix * kOC + lane] += acc;
    }
    __syncthreads();
    return len;
}

__device__ __forceinline__ void load_weight_t(const StemArgs& a, const Smem& s) {
    for (int i = threadIdx.x; i < kIC * kOC; i += kThreads) {
        const int k = i / kOC, o = i - k * kOC;
        s.wt[i] = a.weight[o * kIC + k];
        s.wo[i] = a.weight[i];
    }
}

__global__ void __launch_bounds__(kThreads, 2)
stem_fwd_kernel(const StemArgs a, float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi,
                __nv_bfloat16* __restrict__ out_lo, int out_c) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int W = a.W, npix = kRows * W, bands = a.H / kRows;
    const Smem s = carve(smem_raw, npix, a.E);
    load_weight_t(a, s);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int PW = W / 2, PH = a.H / 2;
    for (int tile = blockIdx.x; tile < a.N * bands; tile += gridDim.x) {
        const int n = tile / bands, y0 = (tile - n * bands) * kRows;
        build_pre_tile(a, s, n, y0);
        // relu + 2x2 max-pool -> [n, y0/2 + {0,1}, 0..PW, 0..31] (channels >= 32 are the zero padding)
        for (int pp = warp; pp < (kRows / 2) * PW; pp += kWarps) {
            const int py = pp / PW, px = pp - py * PW;
            const int p00 = (2 * py) * W + 2 * px;
            float v = fmaxf(fmaxf(s.pre[p00 * kOC + lane], s.pre[(p00 + 1) * kOC + lane]),
                            fmaxf(s.pre[(p00 + W) * kOC + lane], s.pre[(p00 + W + 1) * kOC + lane]));
            v = fmaxf(v, 0.f);
            const size_t o = (((size_t)n * PH + y0 / 2 + py) * PW + px) * out_c;
            out[o + lane] = v;
            if (out_c > kOC) out[o + kOC + lane] = 0.f;
            if (out_hi) {
                const __nv_bfloat16 h = __float2bfloat16_rn(v);
                out_hi[o + lane] = h;
                out_lo[o + lane] = __float2bfloat16_rn(v - __bfloat162float(h));
                if (out_c > kOC) { out_hi[o + kOC + lane] = __float2bfloat16_rn(0.f); out_lo[o + kOC + lane] = __float2bfloat16_rn(0.f); }
            }
        }
        __syncthreads();
    }
}

// Backward: dW/db accumulate in shared memory across the tiles of a persistent CTA and are flushed once.
__global__ void __launch_bounds__(kThreads, 2)
stem_bwd_kernel(const StemArgs a, const float* __restrict__ gout, int out_c, float* __restrict__ gweight,
                float* __restrict__ gbias, float* __restrict__ gproject) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int W = a.W, npix = kRows * W, bands = a.H / kRows;
    const Smem s = carve(smem_raw, npix, a.E);
    float* gw = reinterpret_cast<float*>(smem_raw + stem_smem_bytes(npix, a.E) + 16 - (stem_smem_bytes(npix, a.E) & 15));   // [57][32]
    load_weight_t(a, s);
    for (int i = threadIdx.x; i < (kIC + 1) * kOC; i += kThreads) gw[i] = 0.f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int PW = W / 2, PH = a.H / 2;
    float racc[25];
#pragma unroll
    for (int i = 0; i < 25; ++i) racc[i] = 0.f;
    for (int tile = blockIdx.x; tile < a.N * bands; tile += gridDim.x) {
        const int n = tile / bands, y0 = (tile - n * bands) * kRows;
        const int len = build_pre_tile(a, s, n, y0);
        // route dOut through max-pool (first maximum in window scan order) and ReLU; overwrite pre with dpre.  Only the winning
        // pixel of a window receives a gradient, so the parameter gradients of the dense part are accumulated right here, once
        // per POOLED pixel and lane (25 register accumulators: bias, height, 17 one-hot columns, 6 effect planes; predicated adds,
        // no atomics) instead of in a second pass over all 4x as many input pixels.
        const uint8_t* effb = reinterpret_cast<const uint8_t*>(s.eff);
        for (int pp = warp; pp < (kRows / 2) * PW; pp += kWarps) {
            const int py = pp / PW, px = pp - py * PW;
            const int p00 = (2 * py) * W + 2 * px;
            const int cand[4] = {p00, p00 + 1, p00 + W, p00 + W + 1};
            float best = s.pre[cand[0] * kOC + lane];
            int bi = 0;
#pragma unroll
            for (int k = 1; k < 4; ++k) {
                const float v = s.pre[cand[k] * kOC + lane];
                if (v > best) { best = v; bi = k; }
            }
            const float g = (best > 0.f) ? gout[((((size_t)n * PH + y0 / 2 + py) * PW + px) * out_c) + lane] : 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) s.pre[cand[k] * kOC + lane] = (k == bi) ? g : 0.f;
            if (__ballot_sync(0xffffffffu, g != 0.f) == 0u) continue;
            const int pw = p00 + (bi & 1) + (bi >> 1) * W;           // this lane's winning pixel
            racc[0] += g;
            racc[1] += g * ((float)s.hgt[pw] * (1.0f / 256.0f));
            int slot = 2, c = (int)s.cidx[pw];
#pragma unroll
            for (int k = 1; k < kPlanes; ++k) {
                const int idx = c % vocab_c(k);                     // digits of the combined index, least significant plane first
                c /= vocab_c(k);
#pragma unroll
                for (int v = 0; v < vocab_c(k); ++v) racc[slot + v] += (idx == v) ? g : 0.f;
                slot += vocab_c(k);
            }
            const unsigned e = effb[pw];
            if (__ballot_sync(0xffffffffu, e != 0u) != 0u) {
#pragma unroll
                for (int j = 0; j < kEffects; ++j) racc[19 + j] += ((e >> j) & 1u) ? g : 0.f;
            }
        }
        __syncthreads();
        // entities: d_project[e, c] = sum_o W[o, 24+c] dpre[pix, o];  dW[o, 24+c] += dpre[pix, o] * project[e, c]
        const float* prow = a.project + (size_t)n * a.E * kOC;
        for (int i = warp; i < len; i += kWarps) {
            const uint32_t ent = s.list[i];
            const int pix = ent & 0xffff, e = ent >> 16;
            const float d = s.pre[pix * kOC + lane];                    // dpre[pix, o = lane]
            const float pe = __ldg(prow + (size_t)e * kOC + lane);      // project[e, c = lane]
            float acc = 0.f;
#pragma unroll
            for (int o = 0; o < kOC; ++o) acc += s.wo[o * kIC + 24 + lane] * __shfl_sync(0xffffffffu, d, o);
            gproject[((size_t)n * a.E + e) * kOC + lane] = acc;
            if (__ballot_sync(0xffffffffu, d != 0.f) != 0u) {
#pragma unroll
                for (int c = 0; c < kOC; ++c)
                    atomicAdd(&gw[(24 + c) * kOC + lane], d * __shfl_sync(0xffffffffu, pe, c));
            }
        }
        __syncthreads();
    }
    // fold the register accumulators: slot 0 = bias, 1 = height (column 0), 2..18 = one-hot columns 1..17, 19..24 = effects
    atomicAdd(&gw[kIC * kOC + lane], racc[0]);
    atomicAdd(&gw[lane], racc[1]);
#pragma unroll
    for (int i = 2; i < 25; ++i) atomicAdd(&gw[(i - 1) * kOC + lane], racc[i]);
    __syncthreads();
    for (int i = threadIdx.x; i < kIC * kOC; i += kThreads) {
        const int k = i / kOC, o = i - k * kOC;
        atomicAdd(&gweight[o * kIC + k], gw[i]);
    }
    if (threadIdx.x < kOC) atomicAdd(&gbias[threadIdx.x], gw[kIC * kOC + threadIdx.x]);
}

// lut[combo][o] = sum over the six categorical planes of W[o, base_k + digit_k(combo)]
__global__ void stem_lut_kernel(const float* __restrict__ weight, float* __restrict__ lut) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kCombos * kOC) return;
    const int o = i % kOC;
    int c = i / kOC;
    float v = 0.f;
#pragma unroll
    for (int k = 1; k < kPlanes; ++k) {
        const int digit = c % vocab_c(k);
        c /= vocab_c(k);
        v += weight[o * kIC + kPlaneBase[k] + digit];
    }
    lut[i] = v;
}

int fill_args(StemArgs& a, const void* const* planes, const void* const* effects, const float* project,
              const uint8_t* ex, const uint8_t* ey, const int64_t* entity_num, const float* weight, const float* bias,
              float* lut, int N, int E, int H, int W, cudaStream_t stream) {
    DSB_REQUIRE(planes && effects && project && ex && ey && weight && bias && lut, "spatial_stem: null pointer");
    DSB_REQUIRE(N >= 0 && E > 0 && E <= 65535 && H % kRows == 0 && W % 32 == 0 && W % 2 == 0 && kRows * W <= 65535,
                "spatial_stem: need H %% %d == 0 and W %% 32 == 0 (H=%d W=%d)", kRows, H, W);
    for (int k = 0; k < kPlanes; ++k) { DSB_REQUIRE(planes[k], "spatial_stem: null plane"); a.planes[k] = (const uint8_t*)planes[k]; }
    for (int j = 0; j < kEffects; ++j) { DSB_REQUIRE(effects[j], "spatial_stem: null effect list"); a.effects[j] = (const int16_t*)effects[j]; }
    a.project = project; a.ex = ex; a.ey = ey; a.entity_num = entity_num; a.weight = weight; a.bias = bias; a.lut = lut;
    a.N = N; a.E = E; a.H = H; a.W = W;
    if (N > 0) {
        stem_lut_kernel<<<(kCombos * kOC + 255) / 256, 256, 0, stream>>>(weight, lut);
        return dsb::check_launch("spatial_stem_lut");
    }
    return DSB_OK;
}

}  // namespace

extern "C" int dsb_spatial_stem_fwd(const void* const* planes, const void* const* effects, const float* project,
                                    const uint8_t* ex, const uint8_t* ey, const int64_t* entity_num, const float* weight,
                                    const float* bias, float* lut_workspace, float* out, void* out_hi, void* out_lo, int out_c,
                                    int N, int E, int H, int W, dsb_stream_t stream) {
    StemArgs a;
    int rc = fill_args(a, planes, effects, project, ex, ey, entity_num, weight, bias, lut_workspace, N, E, H, W,
                       (cudaStream_t)stream);
    if (rc) return rc;
    DSB_REQUIRE(out && (out_c == 32 || out_c == 64) && (!out_hi == !out_lo), "spatial_stem_fwd: bad output arguments");
    if (N == 0) return DSB_OK;
    const size_t smem = stem_smem_bytes(kRows * W, E);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(stem_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { dsb::set_error("spatial_stem_fwd smem attr: %s", cudaGetErrorString(e)); return DSB_ERR_CUDA; }
        configured = smem;
    }
    const int64_t tiles = (int64_t)N * (H / kRows);
    const unsigned grid = (unsigned)(tiles < 148 * 2 ? tiles : 148 * 2);
    stem_fwd_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(a, out, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, out_c);
    return dsb::check_launch("spatial_stem_fwd");
}

extern "C" int dsb_spatial_stem_bwd(const void* const* planes, const void* const* effects, const float* project,
                                    const uint8_t* ex, const uint8_t* ey, const int64_t* entity_num, const float* weight,
                                    const float* bias, float* lut_workspace, const float* grad_out, int out_c,
                                    float* grad_weight, float* grad_bias, float* grad_project, int N, int E, int H, int W,
                                    dsb_stream_t stream) {
    StemArgs a;
    int rc = fill_args(a, planes, effects, project, ex, ey, entity_num, weight, bias, lut_workspace, N, E, H, W,
                       (cudaStream_t)stream);
    if (rc) return rc;
    DSB_REQUIRE(grad_out && grad_weight && grad_bias && grad_project && (out_c == 32 || out_c == 64),
                "spatial_stem_bwd: bad arguments (grad_weight/grad_bias/grad_project must be zero-initialised by the caller)");
    if (N == 0) return DSB_OK;
    const size_t smem = stem_smem_bytes(kRows * W, E) + 16 + (size_t)(kIC + 1) * kOC * 4;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(stem_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { dsb::set_error("spatial_stem_bwd smem attr: %s", cudaGetErrorString(e)); return DSB_ERR_CUDA; }
        configured = smem;
    }
    const int64_t tiles = (int64_t)N * (H / kRows);
    const unsigned grid = (unsigned)(tiles < 148 * 2 ? tiles : 148 * 2);
    stem_bwd_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(a, grad_out, out_c, grad_weight, grad_bias, grad_project);
    return dsb::check_launch("spatial_stem_bwd");
}
