// On-device assembly of the learner batch from compact trajectories.
//
// Reference: rl_training/rl_dataloader.py:45-76 (collate_fn) and :206-245 (padding_entity_info) run on the host for every
// trajectory step: entity fields are zero-padded to the batch maximum, the selected-units / target-unit teacher logits and
// behaviour log-probs are padded with -1e9 to [64, 513] / [512] / [64], three sequence masks are built, and the padded batch
// (1.47 GB at B=128 x T=32, 0.54 GB of it the [T,B,64,513] teacher logits that are -1e9 almost everywhere) crosses PCIe.
// Here the host ships the un-padded payload plus per-frame lengths, and three kernels expand it in HBM:
//   expand_ragged   dst[r, s, e] = (s < steps[r] && e < width[r]) ? src[off[r] + s * width[r] + e] : fill   (1-, 2-, 4-byte types)
//   sequence_mask   dst[r, j] = j < len[r]                                                      (rl_dataloader.py:212-213,240-243)
//   unpack_planes   the six categorical spatial planes travel bit-packed in one uint16 per pixel next to the uint8 height map
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

template <typename T>
__global__ void expand_ragged_kernel(const T* __restrict__ src, const int64_t* __restrict__ off, const int* __restrict__ steps,
                                     const int* __restrict__ width, T* __restrict__ dst, int64_t rows, int S, int W, T fill) {
    const int64_t per_row = (int64_t)S * W, total = rows * per_row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / per_row;
        const int64_t rem = i - r * per_row;
        const int s = (int)(rem / W), e = (int)(rem - (int64_t)s * W);
        const int w = width[r], st = steps ? steps[r] : 1;
        dst[i] = (s < st && e < w) ? src[off[r] + (int64_t)s * w + e] : fill;
    }
}

__global__ void sequence_mask_kernel(const int64_t* __restrict__ len, int add, uint8_t* __restrict__ dst, int64_t rows, int W) {
    const int64_t total = rows * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / W;
        dst[i] = (i - r * W) < len[r] + add ? 1 : 0;
    }
}

// packed bits: [0,2) visibility_map, [2] creep, [3,6) player_relative, [6] alerts, [7] pathable, [8] buildable
__global__ void unpack_planes_kernel(const uint16_t* __restrict__ packed, uint8_t* __restrict__ vis, uint8_t* __restrict__ creep,
                                     uint8_t* __restrict__ rel, uint8_t* __restrict__ alerts, uint8_t* __restrict__ path,
                                     uint8_t* __restrict__ build, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned v = packed[i];
        vis[i] = v & 3u; creep[i] = (v >> 2) & 1u; rel[i] = (v >> 3) & 7u; alerts[i] = (v >> 6) & 1u;
        path[i] = (v >> 7) & 1u; build[i] = (v >> 8) & 1u;
    }
}

inline unsigned grid_for(int64_t n) {
    int64_t blocks = (n + kThreads - 1) / kThreads;
    const int64_t cap = 148 * 16;
    return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace

extern "C" int dsb_expand_ragged(const void* src, const int64_t* row_offset, const int* steps, const int* width, void* dst,
                                 int64_t rows, int S, int W, int elem_bytes, double fill, int fill_is_float,
                                 dsb_stream_t stream) {
    DSB_REQUIRE(src && row_offset && width && dst && rows >= 0 && S > 0 && W > 0, "expand_ragged: bad argument");
    if (rows == 0) return DSB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = grid_for(rows * S * W);
    if (elem_bytes == 4) {
        if (fill_is_float) {
            expand_ragged_kernel<float><<<grid, kThreads, 0, s>>>((const float*)src, row_offset, steps, width, (float*)dst, rows, S,
                                                                 W, (float)fill);
        } else {
            expand_ragged_kernel<int32_t><<<grid, kThreads, 0, s>>>((const int32_t*)src, row_offset, steps, width, (int32_t*)dst,
                                                                   rows, S, W, (int32_t)fill);
        }
    } else if (elem_bytes == 2) {
        expand_ragged_kernel<uint16_t><<<grid, kThreads, 0, s>>>((const uint16_t*)src, row_offset, steps, width, (uint16_t*)dst, rows,
                                                                S, W, (uint16_t)(int64_t)fill);
    } else if (elem_bytes == 1) {
        expand_ragged_kernel<uint8_t><<<grid, kThreads, 0, s>>>((const uint8_t*)src, row_offset, steps, width, (uint8_t*)dst, rows, S,
                                                               W, (uint8_t)(int64_t)fill);
    } else {
        dsb::set_error("expand_ragged: elem_bytes must be 1, 2 or 4");
        return DSB_ERR_ARG;
    }
    return dsb::check_launch("expand_ragged");
}

extern "C" int dsb_sequence_mask(const int64_t* lengths, int add, uint8_t* dst, int64_t rows, int W, dsb_stream_t stream) {
    DSB_REQUIRE(lengths && dst && rows >= 0 && W > 0, "sequence_mask: bad argument");
    if (rows == 0) return DSB_OK;
    sequence_mask_kernel<<<grid_for(rows * W), kThreads, 0, (cudaStream_t)stream>>>(lengths, add, dst, rows, W);
    return dsb::check_launch("sequence_mask");
}

extern "C" int dsb_unpack_planes(const uint16_t* packed, uint8_t* visibility, uint8_t* creep, uint8_t* player_relative,
                                 uint8_t* alerts, uint8_t* pathable, uint8_t* buildable, int64_t n, dsb_stream_t stream) {
    DSB_REQUIRE(packed && visibility && creep && player_relative && alerts && pathable && buildable && n >= 0,
                "unpack_planes: bad argument");
    if (n == 0) return DSB_OK;
    unpack_planes_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(packed, visibility, creep, player_relative, alerts,
                                                                            pathable, buildable, n);
    return dsb::check_launch("unpack_planes");
}
