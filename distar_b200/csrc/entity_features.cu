// Entity feature expansion (K1): 36 raw per-entity fields (uint8 / int16 / int8 / fp16) -> the 997-wide one-hot / binary /
// scalar feature row of EntityEncoder.forward (obs_encoder/entity_encoder.py:59-78), written DIRECTLY as the bf16
// (hi, lo) operand pair of the embedding GEMM, zero padded to 1024 columns.  The fp32 [tokens, 997] concat of the
// reference (2 KB/token written, read, split) is never materialised.  One warp per token: lanes first own FIELDS (column bits
// into a shared 1024-bit mask), then COLUMNS [32 l, 32 l + 32) of the row; the 8 scalar fields carry a real (hi, lo).
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
    int scalar_idx[kFields];   // ordinal of a scalar field among the scalar fields (filled by the host wrapper)
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

// Lane f owns FIELD f (and f + 32): it turns its value into column bits of the token's 1024-bit row mask in shared memory
// (one-hot: one bit, binary: up to 11), lane l then expands mask word l into its 32 bf16 columns, the 8 scalar fields are patched
// in as 16-bit stores, and the row leaves as 16-byte stores.  ~120 warp instructions per token (the first version walked all 36
// fields in every lane: ~800, which made the kernel issue bound at 0.65 TB/s).
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
entity_features_kernel(const FieldTable t, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int64_t tokens,
                       int* __restrict__ error_flag) {
    __shared__ __align__(16) uint32_t bits[kWarpsPerBlock][32];
    __shared__ __align__(16) uint32_t row_hi[kWarpsPerBlock][kWidth / 2];
    __shared__ __align__(16) uint32_t row_lo[kWarpsPerBlock][kWidth / 2];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tok = (int64_t)blockIdx.x * kWarpsPerBlock + w;
    if (tok >= tokens) return;
    bits[w][lane] = 0u;
    __syncwarp();
    bool bad = false;
    // scalar fields of this lane (at most two: fields lane and lane + 32), remembered for the patch phase
    int sc_col[2] = {-1, -1}, sc_idx[2] = {0, 0};
    uint16_t sc_h[2] = {0, 0}, sc_l[2] = {0, 0};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int f = lane + 32 * r;
        if (f >= kFields) continue;
        const float v = load_field(t, f, tok);
        const int kind = t.kind[f], off = t.offset[f];
        if (kind == 0) {
            int id = (int)v;
            if (id < 0) { bad = true; id = 0; }
            id = min(id, t.vocab[f] - 1);
            const int col = off + id;
            atomicOr(&bits[w][col >> 5], 1u << (col & 31));
        } else if (kind == 1) {
            const int id = (int)v;
            for (int b = 0; b < 11; ++b)
                if ((id >> (10 - b)) & 1) { const int col = off + b; atomicOr(&bits[w][col >> 5], 1u << (col & 31)); }
        } else {
            const __nv_bfloat16 hb = __float2bfloat16_rn(v);
            const __nv_bfloat16 lb = __float2bfloat16_rn(v - __bfloat162float(hb));
            sc_col[r] = off;
            sc_h[r] = *reinterpret_cast<const uint16_t*>(&hb);
            sc_l[r] = *reinterpret_cast<const uint16_t*>(&lb);
            sc_idx[r] = t.scalar_idx[f];
        }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(error_flag, 1);
    __syncwarp();
    const uint32_t mask = bits[w][lane];
    uint4* rh = reinterpret_cast<uint4*>(&row_hi[w][lane * 16]);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t h[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = 4 * q + e;                      // columns 2j, 2j + 1 of this lane's 32
            h[e] = (((mask >> (2 * j)) & 1u) * 0x3F80u) | (((mask >> (2 * j + 1)) & 1u) * 0x3F800000u);
        }
        rh[q] = make_uint4(h[0], h[1], h[2], h[3]);
        if (lo) reinterpret_cast<uint4*>(&row_lo[w][lane * 16])[q] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncwarp();
    uint16_t* rh16 = reinterpret_cast<uint16_t*>(row_hi[w]);
    uint16_t* rl16 = reinterpret_cast<uint16_t*>(row_lo[w]);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (sc_col[r] < 0) continue;
        rh16[sc_col[r]] = sc_h[r];
        if (lo) rl16[sc_col[r]] = sc_l[r];
        if (t.lo_col_base >= 0) rh16[t.lo_col_base + sc_idx[r]] = sc_l[r];   // exact-operand layout: the residual rides in a spare column
    }
    __syncwarp();
    uint4* dh = reinterpret_cast<uint4*>(hi + tok * kWidth + lane * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) dh[j] = rh[j];
    if (lo) {
        uint4* dl = reinterpret_cast<uint4*>(lo + tok * kWidth + lane * 32);
        const uint4* rl = reinterpret_cast<const uint4*>(&row_lo[w][lane * 16]);
#pragma unroll
        for (int j = 0; j < 4; ++j) dl[j] = rl[j];
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
        t.scalar_idx[f] = scalars;
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
