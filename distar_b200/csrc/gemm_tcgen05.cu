// Split-precision tcgen05 GEMM:  C[M,N] = act(A[M,K] . W[N,K]^T + bias)   — the fc_block workhorse
// (ctools/torch_utils/network/nn_module.py:231-270) behind the entity transformer's QKV / out-proj / MLP
// layers (model/module_utils.py:88-139), the LSTM input projection and the value / head MLPs.
//
// Why "split": the parity contract is 1e-3 on logits against an fp32 reference; plain bf16 operands miss it by
// >10x and TF32 by ~3x (measured on the CPU oracle).  Each fp32 operand x is therefore
// carried as a bf16 pair (hi = bf16(x), lo = bf16(x - hi)) and the product is formed on the tensor cores as
// hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM (relative error ~2^-16).  terms = 1 gives plain bf16.
//
// Structure (one CTA per SM, persistent over 128x128 output tiles, m-major tile order so the CTAs that share
// an A tile run together and A streams from HBM once):
//   warp 0      TMA producer: cp.async.bulk.tensor 128x64 bf16 boxes (SWIZZLE_128B) for A_hi/A_lo/W_hi/W_lo
//   warp 1      MMA issuer  : one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M128 N128 K16),
//               tcgen05.commit releases smem stages and publishes the TMEM accumulator
//   warps 2..5  epilogue    : tcgen05.ld 32x32b -> +bias, ReLU -> swizzled smem box -> TMA store of fp32 C
//               (and an optional bf16 hi/lo split of C)
//   kStages-deep smem ring (mbarrier full/empty) and a 2-deep TMEM accumulator ring (tmem_full/tmem_empty) so
//   the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int kStages = 3;
constexpr int kAccStages = 2;
constexpr int kTileBytes = BM * BK * 2;                 // 16 KiB: one 128x64 bf16 operand tile
constexpr int kStageBytes = 4 * kTileBytes;             // A_hi, A_lo, W_hi, W_lo
constexpr int kStoreBufBytes = 32 * 128;                // one 32-row x 32-col fp32 staging box (128 B rows, SWIZZLE_128B)
constexpr int kStoreBytes = 4 * 2 * kStoreBufBytes;     // 4 epilogue warps x double buffer
constexpr int kSmemBytes = kStages * kStageBytes + kStoreBytes + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int kTmemCols = kAccStages * BN;              // 256
constexpr int kThreads = 192;
constexpr int UMMA_K = 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B operand tile (rows of 128 B, 8-row groups of 1024 B): cute::UMMA::SmemDescriptor
// (cute/arch/mma_sm100_desc.hpp): start>>4 [0,14), LBO>>4 [16,30) (=1, unused when swizzled), SBO>>4 [32,46)
// (=1024 B), version=1 [46,48), layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// cute::UMMA::InstrDescriptor: c_format F32=1 [4,6), a/b_format BF16=1 [7,10)/[10,13), K-major both,
// n_dim = N>>3 [17,23), m_dim = M>>4 [24,29).
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct GemmParams {
    const float* bias;
    float* c;
    __nv_bfloat16* c_hi;
    __nv_bfloat16* c_lo;
    int64_t M;
    int N, K, terms, relu;
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_split_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                  const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                  const __grid_constant__ CUtensorMap map_c, const GemmParams p) {
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B tiles need 1024 B alignment
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* store_bufs = smem + kStages * kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(store_bufs + kStoreBytes);
    uint64_t* full = bars;                       // [kStages]
    uint64_t* empty = bars + kStages;            // [kStages]
    uint64_t* tmem_full = bars + 2 * kStages;    // [kAccStages]
    uint64_t* tmem_empty = tmem_full + kAccStages;
    uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(tmem_empty + kAccStages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_m = (int)((p.M + BM - 1) / BM), num_n = p.N / BN, num_k = p.K / BK;
    const int num_tiles = num_m * num_n;
    const bool three = p.terms == 3;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w_hi) : "memory");
        if (three) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w_lo) : "memory");
        }
        for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < kAccStages; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM allocation is warp-collective; the same warp deallocates at the end
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_ptr)),
                     "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_ptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx = (three ? 4 : 2) * kTileBytes;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / num_n) * BM, n0 = (tile % num_n) * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char* st = smem + stage * kStageBytes;
                    mbar_expect_tx(&full[stage], tx);
                    tma_load_2d(&map_a_hi, &full[stage], st, kb * BK, m0);
                    tma_load_2d(&map_w_hi, &full[stage], st + 2 * kTileBytes, kb * BK, n0);
                    if (three) {
                        tma_load_2d(&map_a_lo, &full[stage], st + kTileBytes, kb * BK, m0);
                        tma_load_2d(&map_w_lo, &full[stage], st + 3 * kTileBytes, kb * BK, n0);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kb = 0; kb < num_k; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sa = smem_u32(smem + stage * kStageBytes);
                    const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + kTileBytes);
                    const uint64_t w_hi = make_desc(sa + 2 * kTileBytes), w_lo = make_desc(sa + 3 * kTileBytes);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);   // +32 B per K step inside the swizzle span
                        const uint32_t first = (kb | k) ? 1u : 0u;
                        if (three) {
                            umma(d_tmem, a_lo + adv, w_hi + adv, first);
                            umma(d_tmem, a_hi + adv, w_lo + adv, 1u);
                            umma(d_tmem, a_hi + adv, w_hi + adv, 1u);
                        } else {
                            umma(d_tmem, a_hi + adv, w_hi + adv, first);
                        }
                    }
                    umma_commit(&empty[stage]);                       // frees this smem stage when the MMAs retire
                    if (kb == num_k - 1) umma_commit(&tmem_full[acc]);  // accumulator complete
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        // TMEM -> registers -> (+bias, ReLU) -> swizzled smem box -> TMA store: every global write is a full,
        // coalesced 128 B line issued by the copy engine; the staging box is double buffered per warp.
        const int q = warp & 3;                    // TMEM lane quarter this warp may access
        unsigned char* my_bufs = store_bufs + (warp - 2) * 2 * kStoreBufBytes;
        int it = 0;
        uint32_t chunk_no = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int m0 = (tile / num_n) * BM, n0 = (tile % num_n) * BN;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const int64_t row = (int64_t)m0 + q * 32 + lane;
            const bool row_ok = row < p.M;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32, ++chunk_no) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), r);
                if (c0 == BN - 32) {               // accumulator fully read: hand TMEM back to the MMA warp early
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                }
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float x = __uint_as_float(r[j]);
                    if (p.bias) x += __ldg(p.bias + n0 + c0 + j);
                    if (p.relu) x = fmaxf(x, 0.f);
                    v[j] = x;
                }
                unsigned char* buf = my_bufs + (chunk_no & 1) * kStoreBufBytes;
                // the TMA store that last read this buffer (two chunks ago) must have finished reading it
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                const uint32_t rowbase = smem_u32(buf) + lane * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t addr = rowbase + (uint32_t)((j ^ (lane & 7)) << 4);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                                 "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) tma_store_2d(&map_c, buf, n0 + c0, m0 + q * 32);
                if (p.c_hi && row_ok) {
                    uint4* dh = reinterpret_cast<uint4*>(p.c_hi + row * p.N + n0 + c0);
                    uint4* dl = reinterpret_cast<uint4*>(p.c_lo + row * p.N + n0 + c0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t h[4], l[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float x0 = v[8 * j + 2 * e], x1 = v[8 * j + 2 * e + 1];
                            const __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
                            const float2 hf = __bfloat1622float2(hh);
                            const __nv_bfloat162 ll = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
                            h[e] = *reinterpret_cast<const uint32_t*>(&hh);
                            l[e] = *reinterpret_cast<const uint32_t*>(&ll);
                        }
                        dh[j] = make_uint4(h[0], h[1], h[2], h[3]);
                        dl[j] = make_uint4(l[0], l[1], l[2], l[3]);
                    }
                }
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// 2-D bf16 row-major [rows, cols] tensor, box = [BM rows, BK cols], 128 B swizzle
int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { dsb::set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return DSB_ERR_CUDA; }
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { dsb::set_error("cuTensorMapEncodeTiled failed: %d", (int)r); return DSB_ERR_CUDA; }
    return DSB_OK;
}

// fp32 row-major [rows, cols] output, box = 32 x 32, 128 B swizzle (matches the epilogue's staging layout)
int make_map_c(CUtensorMap* map, const void* base, int64_t rows, int64_t cols) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { dsb::set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return DSB_ERR_CUDA; }
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    const cuuint32_t box[2] = {32, 32};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { dsb::set_error("cuTensorMapEncodeTiled(C) failed: %d", (int)r); return DSB_ERR_CUDA; }
    return DSB_OK;
}

}  // namespace

extern "C" int dsb_gemm_bf16_split(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo,
                                   const float* bias, float* c, void* c_hi, void* c_lo, int64_t M, int N, int K,
                                   int terms, int relu, dsb_stream_t stream) {
    DSB_REQUIRE(a_hi && w_hi && c, "gemm: null pointer");
    DSB_REQUIRE(terms == 1 || terms == 3, "gemm: terms must be 1 or 3");
    DSB_REQUIRE(terms == 1 || (a_lo && w_lo), "gemm: terms=3 needs the lo halves");
    DSB_REQUIRE(!c_hi == !c_lo, "gemm: c_hi and c_lo go together");
    DSB_REQUIRE(M >= 0 && N > 0 && K > 0 && N % BN == 0 && K % BK == 0, "gemm: need N %% %d == 0 and K %% %d == 0 (N=%d K=%d)",
                BN, BK, N, K);
    DSB_REQUIRE((M + BM - 1) / BM * (int64_t)(N / BN) < (1ll << 31), "gemm: too many tiles");
    if (M == 0) return DSB_OK;
    CUtensorMap ma_hi, ma_lo, mw_hi, mw_lo, mc;
    int rc;
    if ((rc = make_map_c(&mc, c, M, N))) return rc;
    if ((rc = make_map(&ma_hi, a_hi, M, K))) return rc;
    if ((rc = make_map(&mw_hi, w_hi, N, K))) return rc;
    if ((rc = make_map(&ma_lo, terms == 3 ? a_lo : a_hi, M, K))) return rc;
    if ((rc = make_map(&mw_lo, terms == 3 ? w_lo : w_hi, N, K))) return rc;
    static bool configured = false;
    static int num_sms = 0;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) { dsb::set_error("gemm smem attr: %s", cudaGetErrorString(e)); return DSB_ERR_CUDA; }
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        configured = true;
    }
    GemmParams p;
    p.bias = bias; p.c = c; p.c_hi = (__nv_bfloat16*)c_hi; p.c_lo = (__nv_bfloat16*)c_lo;
    p.M = M; p.N = N; p.K = K; p.terms = terms; p.relu = relu;
    const int64_t tiles = (M + BM - 1) / BM * (int64_t)(N / BN);
    const unsigned grid = (unsigned)(tiles < num_sms ? tiles : num_sms);
    gemm_split_kernel<<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(ma_hi, ma_lo, mw_hi, mw_lo, mc, p);
    return dsb::check_launch("gemm_bf16_split");
}
