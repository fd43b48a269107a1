// scatter_connection forward/backward — replaces module_utils.py:11-34 (+ the entity mask of encoder.py:37-38).
//
// Forward: one CTA owns (obs n, band of ROWS map rows) for all 32 channels.  It zeroes a [ROWS*W][32] fp32 tile
// in shared memory, builds the ordered list of this band's entities (ballot compaction keeps entity order so
// the per-pixel sum order equals the reference's sequential scatter_add_ -> bit-exact vs the CPU oracle), adds
// each entity's 32-channel row (one coalesced 128 B read) into the tile, then streams the tile out so that every
// output element is written exactly once with fully coalesced 128 B warp stores.  No memset pass, no int64
// index tensor, no global atomics.  Algorithmic traffic per obs: 2 MiB written + 64 KiB + 1 KiB read.
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int kC = 32;        // scatter_dim (actor_critic_default_config.yaml: encoder.scatter.output_dim)
constexpr int kPad = 4;       // channel stride = npix + 4 floats: keeps 16 B alignment, 4-way (not 32-way) bank conflicts

// Forward, v2.  The map is >= 97 % zeros (<= 512 entities on 16384 pixels), so the shared-memory tile is never
// cleared nor fully read: an occupancy bitmap says which pixels hold data; the write-out streams 16-byte zero
// vectors straight from registers and only touches the tile where the bitmap is set.  Per CTA the instruction
// stream is essentially the 4096 coalesced float4 streaming stores of its 64 KiB output slab.
template <int ROWS, int WT, int kThreads>   // WT = compile-time map width (0 = runtime)
__global__ void __launch_bounds__(kThreads)
scatter_fwd_kernel(const float* __restrict__ project, const uint8_t* __restrict__ ex,
                   const uint8_t* __restrict__ ey, const int64_t* __restrict__ entity_num,
                   float* __restrict__ out, int E, int H, int Wrt) {
    constexpr int kWarps = kThreads / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int W = WT ? WT : Wrt;
    const int bands = H / ROWS;
    const int n = blockIdx.x / bands;
    const int band = blockIdx.x - n * bands;
    const int y0 = band * ROWS;
    const int npix = ROWS * W;
    const int cstride = npix + kPad;
    float* tile = reinterpret_cast<float*>(smem_raw);                       // [kC][cstride]
    uint32_t* list = reinterpret_cast<uint32_t*>(tile + kC * cstride);      // (e << 16) | pix, ordered by e
    uint32_t* bitmap = list + E;                                            // npix bits
    __shared__ int warp_cnt[kWarps];
    __shared__ int list_len;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwords = (npix + 31) / 32;
    for (int i = tid; i < nwords; i += kThreads) bitmap[i] = 0u;
    if (tid == 0) list_len = 0;
    const int en = entity_num ? min((int)entity_num[n], E) : E;
    __syncthreads();
    // phase 1: ordered compaction of the entities that fall in this band
    for (int base = 0; base < en; base += kThreads) {
        const int e = base + tid;
        int pix = -1;
        if (e < en) {
            const int yy = min((int)ey[(size_t)n * E + e], H - 1);
            const int xx = min((int)ex[(size_t)n * E + e], W - 1);
            if (yy >= y0 && yy < y0 + ROWS) pix = (yy - y0) * W + xx;
        }
        const unsigned m = __ballot_sync(0xffffffffu, pix >= 0);
        if (lane == 0) warp_cnt[warp] = __popc(m);
        __syncthreads();
        int off = list_len;
        for (int w = 0; w < warp; ++w) off += warp_cnt[w];
        if (pix >= 0) list[off + __popc(m & ((1u << lane) - 1))] = ((uint32_t)e << 16) | (uint32_t)pix;
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < kWarps; ++w) tot += warp_cnt[w];
            list_len += tot;
        }
        __syncthreads();
    }
    // phase 2: accumulate.  Warp w owns pixels with (pix % kWarps) == w, so a pixel's additions happen in entity
    // order inside one warp (bit-exact vs the sequential reference).  First touch stores, later touches add.
    const int len = list_len;
    const float* prow = project + (size_t)n * E * kC;
    for (int i = 0; i < len; ++i) {
        const uint32_t v = list[i];
        const int pix = v & 0xffff;
        if ((pix & (kWarps - 1)) != warp) continue;
        const int e = v >> 16;
        const float x = __ldg(prow + (size_t)e * kC + lane);
        const uint32_t bit = 1u << (pix & 31);
        const bool seen = (bitmap[pix >> 5] & bit) != 0u;
        float* slot = tile + lane * cstride + pix;
        *slot = seen ? (*slot + x) : x;
        __syncwarp();
        if (!seen && lane == 0) atomicOr(&bitmap[pix >> 5], bit);
        __syncwarp();
    }
    __syncthreads();
    // phase 3: stream out 16-byte vectors; zeros come from registers
    float* obase = out + (size_t)n * kC * H * W + (size_t)y0 * W;
    const int q = npix >> 2;                        // float4 per channel (W % 4 == 0 checked on the host)
    for (int idx = tid; idx < kC * q; idx += kThreads) {
        const int c = idx / q;
        const int p4 = idx - c * q;
        const int pix = p4 << 2;
        const uint32_t bits = (bitmap[pix >> 5] >> (pix & 31)) & 0xFu;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bits) {
            const float* t = tile + c * cstride + pix;
            if (bits & 1u) v.x = t[0];
            if (bits & 2u) v.y = t[1];
            if (bits & 4u) v.z = t[2];
            if (bits & 8u) v.w = t[3];
        }
        __stcs(reinterpret_cast<float4*>(obase + (size_t)c * H * W + pix), v);
    }
}

__global__ void scatter_bwd_kernel(const float* __restrict__ grad_out, const uint8_t* __restrict__ ex,
                                   const uint8_t* __restrict__ ey, const int64_t* __restrict__ entity_num,
                                   float* __restrict__ grad_project, int64_t total, int E, int H, int W) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % kC);
    const int64_t ne = i / kC;
    const int e = (int)(ne % E);
    const int64_t n = ne / E;
    const int en = entity_num ? (int)entity_num[n] : E;
    float g = 0.f;
    if (e < en) {
        const int yy = min((int)ey[ne], H - 1), xx = min((int)ex[ne], W - 1);
        g = __ldg(grad_out + ((size_t)(n * kC + c) * H + yy) * W + xx);
    }
    grad_project[i] = g;
}

}  // namespace

template <int ROWS, int WT, int kThreads>
static int launch_scatter_fwd(const float* project, const uint8_t* ex, const uint8_t* ey, const int64_t* entity_num,
                              float* out, int N, int E, int H, int W, cudaStream_t stream) {
    const size_t smem = (size_t)kC * (ROWS * W + kPad) * sizeof(float) + (size_t)E * sizeof(uint32_t) +
                        (size_t)((ROWS * W + 31) / 32) * sizeof(uint32_t);
    DSB_REQUIRE(smem <= 220 * 1024, "scatter_connection_fwd: tile does not fit shared memory");
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(scatter_fwd_kernel<ROWS, WT, kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) { dsb::set_error("scatter fwd smem attr: %s", cudaGetErrorString(e)); return DSB_ERR_CUDA; }
        configured = smem;
    }
    const int64_t grid = (int64_t)N * (H / ROWS);
    DSB_REQUIRE(grid < (1ll << 31), "scatter_connection_fwd: grid too large");
    scatter_fwd_kernel<ROWS, WT, kThreads><<<(unsigned)grid, kThreads, smem, stream>>>(project, ex, ey, entity_num, out, E, H, W);
    return dsb::check_launch("scatter_connection_fwd");
}

extern "C" int dsb_scatter_connection_fwd(const float* project, const uint8_t* ex, const uint8_t* ey,
                                          const int64_t* entity_num, float* out, int N, int E, int H, int W,
                                          dsb_stream_t stream) {
    DSB_REQUIRE(project && ex && ey && out, "scatter_connection_fwd: null pointer");
    DSB_REQUIRE(N >= 0 && E > 0 && E <= 65535 && H > 0 && W > 0, "scatter_connection_fwd: bad shape");
    if (N == 0) return DSB_OK;
    DSB_REQUIRE(H % 8 == 0 && W % 4 == 0 && 8 * W <= 65535, "scatter_connection_fwd: need H %% 8 == 0 and W %% 4 == 0");
    DSB_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "scatter_connection_fwd: out must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    static int variant = -1;                 // DEBUG: DSB_SCATTER_VARIANT picks the tile shape for tuning runs
    if (variant < 0) { const char* v = getenv("DSB_SCATTER_VARIANT"); variant = v ? atoi(v) : 0; }
    if (W == 128) {
        switch (variant) {
            case 1: return launch_scatter_fwd<4, 128, 512>(project, ex, ey, entity_num, out, N, E, H, W, s);
            case 6: return launch_scatter_fwd<1, 128, 128>(project, ex, ey, entity_num, out, N, E, H, W, s);
            case 7: return launch_scatter_fwd<1, 128, 256>(project, ex, ey, entity_num, out, N, E, H, W, s);
            case 8: return launch_scatter_fwd<2, 128, 128>(project, ex, ey, entity_num, out, N, E, H, W, s);
            case 2: return launch_scatter_fwd<4, 128, 256>(project, ex, ey, entity_num, out, N, E, H, W, s);
            case 3: return launch_scatter_fwd<8, 128, 512>(project, ex, ey, entity_num, out, N, E, H, W, s);
            case 4: return launch_scatter_fwd<2, 128, 512>(project, ex, ey, entity_num, out, N, E, H, W, s);
            case 5: return launch_scatter_fwd<8, 128, 1024>(project, ex, ey, entity_num, out, N, E, H, W, s);
            default: return launch_scatter_fwd<2, 128, 256>(project, ex, ey, entity_num, out, N, E, H, W, s);
        }
    }
    if (W == 160) return launch_scatter_fwd<2, 160, 256>(project, ex, ey, entity_num, out, N, E, H, W, s);
    return launch_scatter_fwd<2, 0, 256>(project, ex, ey, entity_num, out, N, E, H, W, s);
}

extern "C" int dsb_scatter_connection_bwd(const float* grad_out, const uint8_t* ex, const uint8_t* ey,
                                          const int64_t* entity_num, float* grad_project, int N, int E, int H,
                                          int W, dsb_stream_t stream) {
    DSB_REQUIRE(grad_out && ex && ey && grad_project, "scatter_connection_bwd: null pointer");
    if (N == 0) return DSB_OK;
    const int64_t total = (int64_t)N * E * kC;
    const int threads = 256;
    scatter_bwd_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
        grad_out, ex, ey, entity_num, grad_project, total, E, H, W);
    return dsb::check_launch("scatter_connection_bwd");
}
