// Entity feature expansion (K1): 36 raw per-entity fields (uint8 / int16 / int8 / fp16) -> the 997-wide one-hot / binary /
// scalar feature row of EntityEncoder.forward (obs_encoder/entity_encoder.py:59-78), written DIRECTLY as the bf16
// (hi, lo) operand pair of the embedding GEMM, zero padded to 1024 columns.  The fp32 [tokens, 997] concat of the
// reference (2 KB/token written, read, split) is never materialised.  One warp per token; lane l owns columns
// [32 l, 32 l + 32): one-hot / binary bits are collected in a 32-bit mask, the 8 scalar fields carry a real (hi, lo).
// One-hot ids >= vocab are clamped (entity_encoder.py:73); a negative id raises the error flag (:69-72).
#include <cuda_fp16.h>
#include "common.cuh"

namespace {

constexpr int kFields = 36;
constexpr int kWidth = 1024;
constexpr int kWarpsPerBlock = 8;

struct FieldTable {
    const void* ptr[kFields];
    int kind[kFields];     // 0 one-hot, 1 binary (11 bits, MSB first), 2 scalar
    int offset[kFields];
    int vocab[kFields];
    int dtype[kFields];    // 0 u8, 1 i16, 2 i8, 3 f16
    int lo_col_base;       // >= 0: the lo halves of the scalar fields become extra hi columns [lo_col_base + j], no lo tensor
};

__device__ __forceinline__ float load_field(const FieldTable& t, int f, int64_t tok) {
    switch (t.dtype[f]) {
        case 0: return (float)reinterpret_cast<const uint8_t*>(t.ptr[f])[tok];
        case 1: return (float)reinterpret_cast<const int16_t*>(t.ptr[f])[tok];
        case 2: return (float)reinterpret_cast<const int8_t*>(t.ptr[f])[tok];
        default: return __half2float(reinterpret_cast<const __half*>(t.ptr[f])[tok]);
    }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
entity_features_kernel(const FieldTable t, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int64_t tokens,
                       int* __restrict__ error_flag) {
    const int64_t tok = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (tok >= tokens) return;
    const int lane = threadIdx.x & 31;
    // lanes 0..35 would be nicer; with 32 lanes each lane fetches field `lane` and field `lane + 32` (4 extra)
    const float v0 = load_field(t, lane, tok);
    const float v1 = (lane + 32 < kFields) ? load_field(t, lane + 32, tok) : 0.f;
    uint32_t mask = 0u;
    bool bad = false;
#pragma unroll
    for (int f = 0; f < kFields; ++f) {
        const float v = (f < 32) ? __shfl_sync(0xffffffffu, v0, f) : __shfl_sync(0xffffffffu, v1, f - 32);
        const int kind = t.kind[f], off = t.offset[f];
        if (kind == 0) {
            int id = (int)v;
            if (id < 0) { bad = true; id = 0; }
            id = min(id, t.vocab[f] - 1);
            const int col = off + id;
            if ((col >> 5) == lane) mask |= 1u << (col & 31);
        } else if (kind == 1) {
            const int id = (int)v;
#pragma unroll
            for (int b = 0; b < 11; ++b) {
                const int col = off + b;
                if ((col >> 5) == lane && ((id >> (10 - b)) & 1)) mask |= 1u << (col & 31);
            }
        }   // scalar fields are resolved per column below
    }
    if (bad && lane == 0) atomicOr(error_flag, 1);
    // build the 32 columns; scalar fields may share a lane (e.g. columns 274..277), so resolve them per column
    uint32_t h[16], l[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { h[j] = 0u; l[j] = 0u; }
    const uint32_t one = 0x3F80u;   // bf16 1.0
#pragma unroll
    for (int c = 0; c < 32; ++c)
        if ((mask >> c) & 1u) h[c >> 1] |= one << ((c & 1) * 16);
    int scalar_idx = 0;
#pragma unroll
    for (int f = 0; f < kFields; ++f) {
        if (t.kind[f] != 2) continue;
        const int off = t.offset[f];
        const float v = (f < 32) ? __shfl_sync(0xffffffffu, v0, f) : __shfl_sync(0xffffffffu, v1, f - 32);
        const __nv_bfloat16 hb = __float2bfloat16_rn(v);
        const __nv_bfloat16 lb = __float2bfloat16_rn(v - __bfloat162float(hb));
        const uint32_t hu = (uint32_t)(*reinterpret_cast<const uint16_t*>(&hb));
        const uint32_t lu = (uint32_t)(*reinterpret_cast<const uint16_t*>(&lb));
        if ((off >> 5) == lane) {
            const int c = off & 31;
            // dynamic register index avoided: select with a loop the compiler unrolls
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (j == (c >> 1)) { h[j] |= hu << ((c & 1) * 16); l[j] |= lu << ((c & 1) * 16); }
        }
        if (t.lo_col_base >= 0) {          // exact-operand layout: the residual rides in a spare column of the same row
            const int xc = t.lo_col_base + scalar_idx;
            if ((xc >> 5) == lane) {
                const int c = xc & 31;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j == (c >> 1)) h[j] |= lu << ((c & 1) * 16);
            }
        }
        ++scalar_idx;
    }
    uint4* dh = reinterpret_cast<uint4*>(hi + tok * kWidth + lane * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) dh[j] = make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
    if (lo) {
        uint4* dl = reinterpret_cast<uint4*>(lo + tok * kWidth + lane * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) dl[j] = make_uint4(l[4 * j], l[4 * j + 1], l[4 * j + 2], l[4 * j + 3]);
    }
}

}  // namespace

extern "C" int dsb_entity_features(const void* const* fields, const int* kind, const int* offset, const int* vocab,
                                   const int* dtype, int num_fields, void* hi, void* lo, int lo_col_base, int64_t tokens,
                                   int* error_flag, dsb_stream_t stream) {
    DSB_REQUIRE(fields && kind && offset && vocab && dtype && hi && error_flag, "entity_features: null pointer");
    DSB_REQUIRE(lo || lo_col_base >= 0, "entity_features: either a lo tensor or lo_col_base (exact-operand layout) is needed");
    DSB_REQUIRE(num_fields == kFields, "entity_features: expected %d fields, got %d", kFields, num_fields);
    if (tokens == 0) return DSB_OK;
    FieldTable t;
    t.lo_col_base = lo_col_base;
    int scalars = 0, row_end = 0;
    for (int f = 0; f < kFields; ++f) {
        DSB_REQUIRE(fields[f], "entity_features: null field %d", f);
        DSB_REQUIRE(kind[f] >= 0 && kind[f] <= 2 && dtype[f] >= 0 && dtype[f] <= 3, "entity_features: bad table entry %d", f);
        DSB_REQUIRE(offset[f] >= 0 && offset[f] + (kind[f] == 0 ? vocab[f] : (kind[f] == 1 ? 11 : 1)) <= kWidth,
                    "entity_features: field %d overflows the %d-wide row", f, kWidth);
        scalars += kind[f] == 2;
        { const int e = offset[f] + (kind[f] == 0 ? vocab[f] : (kind[f] == 1 ? 11 : 1)); row_end = e > row_end ? e : row_end; }
        t.ptr[f] = fields[f]; t.kind[f] = kind[f]; t.offset[f] = offset[f]; t.vocab[f] = vocab[f]; t.dtype[f] = dtype[f];
    }
    DSB_REQUIRE(lo_col_base < 0 || (lo_col_base >= row_end && lo_col_base + scalars <= kWidth),
                "entity_features: lo_col_base %d collides with the feature row (ends at %d, %d scalar fields)", lo_col_base, row_end, scalars);
    const int64_t blocks = (tokens + kWarpsPerBlock - 1) / kWarpsPerBlock;
    DSB_REQUIRE(blocks < (1ll << 31), "entity_features: too many tokens");
    entity_features_kernel<<<(unsigned)blocks, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        t, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, tokens, error_flag);
    return dsb::check_launch("entity_features");
}
