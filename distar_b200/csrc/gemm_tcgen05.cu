// Split-precision tcgen05 GEMM family:  C = alpha * op(A) . op(B)^T (+ bias, ReLU)
//
// One persistent warp-specialised kernel serves every GEMM-shaped piece of the hot path:
//   * fc_block forward  y = act(x W^T + b)           (ctools/torch_utils/network/nn_module.py:231-270; the entity
//     transformer's QKV / out-proj / MLP, model/module_utils.py:88-139, LSTM input projection, value/head MLPs)
//   * its input gradient dX = dY W  and weight gradient dW = dY^T X  (split-K over the token dimension)
//   * attention scores S = Q K^T, context O = P V and all five attention-backward products, batched over
//     (observation, head) by pure coordinate arithmetic on ONE tensor map (no per-head copies or transposes).
//
// Why "split": the parity contract is 1e-3 on logits against an fp32 reference; plain bf16 operands miss it by
// >10x and TF32 by ~3x (measured on the CPU oracle).  Each fp32 operand x is therefore
// carried as a bf16 pair (hi = bf16(x), lo = bf16(x - hi)) and the product is formed on the tensor cores as
// hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM (relative error ~2^-16).  terms = 1 gives plain bf16.
//
// Operand layouts.  A "K-major" operand is stored with the reduction index contiguous (x[M,K], W[N,K]): one TMA box
// of 128 rows x 64 k (SWIZZLE_128B).  An "MN-major" operand is stored with the reduction index strided (the natural
// layout of X and dY when reducing over tokens, of V in P.V, of P in P^T.dO ...): two TMA boxes of 64 k-rows x 64 mn,
// consumed through MN-major UMMA descriptors (LBO = 8 KiB between the 64-wide halves, SBO = 1 KiB between 8-row
// k groups).  No transposed copy of any activation is ever materialised.
//
// Structure (one CTA per SM, persistent over 128 x BN output tiles with BN in {64, 128, 256}, n-fastest tile order so CTAs
// that share an A tile run together and A streams from HBM once):
//   warp 0      TMA producer : cp.async.bulk.tensor boxes for A_hi/A_lo/B_hi/B_lo into a 2-3 deep smem ring (implicit-GEMM
//               convolutions read NHWC activations through 4-D boxes, halos are TMA zero fill)
//   warp 1      MMA issuer   : one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M128 N=BN K16), up to three per
//               k step (lo x hi, hi x lo, hi x hi; operands flagged exact skip their lo product and loads);
//               tcgen05.commit frees smem stages and publishes the TMEM accumulator (2-deep ring)
//   warps 2..9  epilogue     : tcgen05.ld 32x32b -> alpha, +bias, (+residual), ReLU -> swizzled smem box -> TMA store of fp32 C
//               and / or of the bf16 (hi, lo) pair of C (SWIZZLE_64B boxes), or TMA reduce-add (split-K, accumulation
//               into a gradient buffer)
// Cluster variants (template MC): 2 = two CTAs share every B tile by TMA multicast; 4 = CTA pair issuing
// tcgen05.mma.cta_group::2 (M = 256 across two SMs, each CTA stages half of B, barriers in the leader CTA; all cross-CTA
// barrier arrivals are .relaxed.cluster - a .release.cluster arrive per stage cost the producer ~0.7 us and made the pair
// slower than the single-CTA kernel).  The automatic choice (launch()) uses the pair for the plain 3-term products with
// >= 74 tile pairs and multicast for 1-term products.
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"

namespace {

constexpr int BM = 128, BK = 64;                        // BN (64, 128 or 256) is a template parameter
constexpr int kAccStages = 2;
constexpr int kTileBytes = BM * BK * 2;                 // 16 KiB: one 128x64 bf16 operand tile (B tiles use BN*128 B of it)
constexpr int kStoreBufBytes = 32 * 128;                // one 32-row x 32-col fp32 staging box (128 B rows, SWIZZLE_128B)
constexpr int kStoreBytes = 8 * kStoreBufBytes;         // 8 epilogue warps x one staging box each (x2 in CTA-pair mode)
// per-BN shared-memory plan: stage = A_hi | A_lo | B_hi | B_lo; B slots are 16 KiB (BN <= 128) or 32 KiB (BN = 256)
// MC = 4 (CTA pair, cta_group::2 MMAs): each CTA stages only its half of every B tile (64 KiB stages), which leaves room
// for a third ring stage (measured: FFN down-projection 182 us with 2 stages + double-buffered staging boxes, 166 us with
// 3 stages + one box; the epilogue alone is HBM-write bound either way).
template <int BN, int MC = 1> struct Plan {
    static constexpr int kBSlot = MC == 4 ? BN * 64 : (BN > 128 ? BN * 128 : kTileBytes);
    static constexpr int kStageBytes = 2 * kTileBytes + 2 * kBSlot;
    static constexpr int kStages = MC == 4 ? (BN > 128 ? 3 : 4) : (BN > 128 ? 2 : 3);
    static constexpr int kStoreBufs = 1;
    static constexpr int kStoreTotal = kStoreBufs * kStoreBytes;
    static constexpr int kSmemBytes = kStages * kStageBytes + kStoreTotal + 1024 /*align slack*/ + 256 /*barriers*/;
};
constexpr int kEpiWarps = 8;                            // two epilogue warps per scheduler: the epilogue is issue-bound
constexpr int kThreads = (2 + kEpiWarps) * 32;          // 320
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
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// multicast variants: the box lands at the same CTA-relative smem offset in every CTA of `mask` and completes tx bytes on
// the mbarrier at the same offset there
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void tma_load_4d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                               int c3, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(mask) : "memory");
}
// ---- CTA-pair (cta_group::2) plumbing: barriers live in the leader CTA (rank 0) and are addressed through the cluster
// window (mapa); loads issued by either CTA complete on the leader's barrier.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t bar_addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t bar_addr, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(const CUtensorMap* map, uint32_t bar_addr, void* dst, int c0, int c1, int c2,
                                                 int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// L2 prefetch of a box (no shared-memory destination, no barrier): issued one tile ahead for the streamed A operand so the
// real load later is an L2 hit - the kernel is bound by the latency of its (at most 2-3 stage) operand ring, not by bandwidth
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1, bool commit = true) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
    if (commit) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int kPending>
__device__ __forceinline__ void tma_store_wait_read() {      // at most kPending store groups may still be reading smem
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}
// split-K accumulation: the copy engine adds the fp32 box into global memory (no partial buffers, no reduce pass)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
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

// cute::UMMA::SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout SWIZZLE_128B=2 [61,64).
//   K-major tile  : rows of 128 B (64 k), 8-row groups of 1024 B  -> SBO = 1024, LBO unused (1)
//   MN-major tile : rows of 128 B (64 mn) indexed by k, 8-k groups of 1024 B -> SBO = 1024; second 64-wide mn half
//                   8192 B further -> LBO = 8192
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, bool mn_major) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(mn_major ? (8192 >> 4) : 1) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// cute::UMMA::InstrDescriptor: c_format F32=1 [4,6), a/b_format BF16=1 [7,10)/[10,13), a_major [15], b_major [16]
// (0 = K, 1 = MN), n_dim = N>>3 [17,23), m_dim = M>>4 [24,29).
__device__ __forceinline__ uint32_t make_idesc(bool a_mn, bool b_mn, int bn, int m = BM) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                     uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// CTA-pair MMA: one instruction drives the tensor cores of both SMs (M = 256: each CTA's 128 rows of A from its own shared
// memory, the N columns of B split half / half between the two CTAs' shared memories, accumulators in each CTA's TMEM).
__device__ __forceinline__ void umma_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {      // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
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

struct OperandMap {       // tile origin = (col_base + bi*col_inner, row_outer*bo + row_inner*bi), see header
    int col_base, col_inner, row_outer, row_inner, mn_major;
    // implicit-GEMM convolution operand (conv != 0): an NHWC tensor [imgs, H, W, C] read through a 4-D tensor map;
    // the halo of a 3x3 tap is whatever TMA zero-fills outside the image.
    int conv, H, W, C, taps;
};

struct GemmParams {
    const float* bias;
    const float* residual;   // optional fp32 [c_rows, ldc] added before the ReLU
    const __nv_bfloat16* mask;   // optional bf16 [c_rows, ldc]: result zeroed where mask == 0 (ReLU backward through a saved output)
    float* colsum;               // optional fp32 [N]: column sums of the (masked) result are ADDED here (a bias gradient)
    __nv_bfloat16* c_hi;
    __nv_bfloat16* c_lo;
    float alpha;
    int64_t M;               // valid rows per batch
    int64_t ldc;             // row stride (elements) of c_hi / c_lo
    int N, terms, relu, accumulate;
    int a_exact, b_exact;    // terms == 3: that operand has no lo half (exactly representable in bf16)
    int store_c;             // 0: only the bf16 (hi, lo) pair is written (no fp32 C)
    int prefetch;            // L2-prefetch the next tile's A operand
    int debug;               // DEV ONLY (env DSB_GEMM_DEBUG, tools/gemm_ablate.sh): 1 = skip global stores, 2 = skip MMAs, 4 = skip
                             // operand loads, 8 = all stores to the first tile (L2 resident), 16 = no staging / store, 32 = no TMEM read
    int num_m, num_n, num_k; // tiles per (batch, split); k blocks per split
    int batch, inner, splits;
    int c_row_outer, c_row_inner, c_row_split, c_col_base, c_col_inner;
    OperandMap a, b;
};

struct TileCoord { int b, s, m0, n0, bo, bi, valid; };

// MC = 1: `tile` enumerates output tiles.  MC = 2 (cluster of two CTAs sharing every B tile by TMA multicast): `tile`
// enumerates PAIRS of vertically adjacent tiles; CTA `rank` owns m tile 2*pair_m + rank (possibly past the end).
template <int BN, int MC>
__device__ __forceinline__ TileCoord decode_tile(int tile, const GemmParams& p, int rank) {
    TileCoord t;
    const int n_i = tile % p.num_n;
    int r = tile / p.num_n;
    const int num_mu = (MC == 1) ? p.num_m : (p.num_m + 1) / 2;
    int m_i = r % num_mu;
    r /= num_mu;
    if (MC != 1) m_i = 2 * m_i + rank;
    t.valid = m_i < p.num_m;
    t.s = r % p.splits;
    t.b = r / p.splits;
    t.m0 = m_i * BM;
    t.n0 = n_i * BN;
    t.bo = t.b / p.inner;
    t.bi = t.b - t.bo * p.inner;
    return t;
}

__device__ __forceinline__ void tap_offset(int tap, int taps, int& dy, int& dx) {
    if (taps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; } else { dy = 0; dx = 0; }
}

// mn0 = first row (A) / column (B) of the output tile this operand tile feeds, kg = first reduction index, width = 128
// for A, BN for B.  mc_rank >= 0: this is the B operand of a 2-CTA cluster: load only this CTA's half of the tile and
// multicast it to both CTAs (the other half arrives from the sibling).
// pair_bar != 0 (CTA pair): every load completes on the leader's barrier `pair_bar` (cluster address); for the B operand
// (mc_rank >= 0) this CTA loads only its half of the tile, into ITS OWN slot at offset 0.
__device__ __forceinline__ void load_operand(const CUtensorMap* map, uint64_t* bar, unsigned char* dst,
                                             const OperandMap& o, const TileCoord& t, int mn0, int kg, int width,
                                             int mc_rank = -1, uint32_t pair_bar = 0) {
    const int halves = width / 64;
    if (pair_bar) {
        const bool half = mc_rank >= 0;
        const int h_lo = half ? mc_rank * (halves / 2) : 0, h_hi = half ? (mc_rank + 1) * (halves / 2) : halves;
        if (o.conv) {
            const int cblocks = o.C / 64;
            if (!o.mn_major) {
                const int pix0 = mn0, hw = o.H * o.W;
                const int img = pix0 / hw, y0 = (pix0 - img * hw) / o.W;
                const int kb = kg / 64, tap = kb / cblocks, cb = kb - tap * cblocks;
                int dy, dx;
                tap_offset(tap, o.taps, dy, dx);
                tma_load_4d_pair(map, pair_bar, dst, cb * 64, dx, y0 + dy, img);
            } else {
                const int hw = o.H * o.W;
                const int img = kg / hw, y0 = (kg - img * hw) / o.W;
                for (int h = h_lo; h < h_hi; ++h) {
                    const int col = mn0 + 64 * h, tap = col / o.C, c = col - tap * o.C;
                    int dy, dx;
                    tap_offset(tap, o.taps, dy, dx);
                    tma_load_4d_pair(map, pair_bar, dst + (h - h_lo) * (kTileBytes / 2), c, dx, y0 + dy, img);
                }
            }
            return;
        }
        const int col0 = o.col_base + t.bi * o.col_inner;
        const int row0 = o.row_outer * t.bo + o.row_inner * t.bi;
        if (!o.mn_major) {      // box {64 k, 128 rows (A) or width/2 rows (B half)}
            tma_load_2d_pair(map, pair_bar, dst, col0 + kg, row0 + mn0 + (half ? mc_rank * (width / 2) : 0));
        } else {
            for (int h = h_lo; h < h_hi; ++h)
                tma_load_2d_pair(map, pair_bar, dst + (h - h_lo) * (kTileBytes / 2), col0 + mn0 + 64 * h, row0 + kg);
        }
        return;
    }
    const bool mc = mc_rank >= 0;
    const int h_lo = mc ? mc_rank * (halves / 2) : 0, h_hi = mc ? (mc_rank + 1) * (halves / 2) : halves;
    if (o.conv) {
        const int cblocks = o.C / 64;
        if (!o.mn_major) {
            // forward / input-gradient convolution: rows = 128 consecutive output pixels (whole image rows),
            // reduction index = (tap, channel)   (A operand only: never multicast)
            const int pix0 = mn0, hw = o.H * o.W;
            const int img = pix0 / hw, y0 = (pix0 - img * hw) / o.W;
            const int kb = kg / 64, tap = kb / cblocks, cb = kb - tap * cblocks;
            int dy, dx;
            tap_offset(tap, o.taps, dy, dx);
            tma_load_4d(map, bar, dst, cb * 64, dx, y0 + dy, img);               // box {64 c, W, 128/W, 1}
        } else {
            // weight-gradient convolution: reduction index = 64 consecutive pixels, columns = (tap, channel)
            const int hw = o.H * o.W;
            const int img = kg / hw, y0 = (kg - img * hw) / o.W;
            for (int h = h_lo; h < h_hi; ++h) {
                const int col = mn0 + 64 * h, tap = col / o.C, c = col - tap * o.C;
                int dy, dx;
                tap_offset(tap, o.taps, dy, dx);
                if (mc) tma_load_4d_mc(map, bar, dst + h * (kTileBytes / 2), c, dx, y0 + dy, img, 3);
                else tma_load_4d(map, bar, dst + h * (kTileBytes / 2), c, dx, y0 + dy, img);   // box {64 c, W, 64/W, 1}
            }
        }
        return;
    }
    const int col0 = o.col_base + t.bi * o.col_inner;
    const int row0 = o.row_outer * t.bo + o.row_inner * t.bi;
    if (!o.mn_major) {
        if (mc) {   // box {64 k, width/2 rows}: rows [rank*width/2, +width/2) of the tile, 128 B per row
            tma_load_2d_mc(map, bar, dst + mc_rank * (width / 2) * 128, col0 + kg, row0 + mn0 + mc_rank * (width / 2), 3);
        } else {
            tma_load_2d(map, bar, dst, col0 + kg, row0 + mn0);                  // box {64 k, 128 (or BN) rows}
        }
    } else {
        for (int h = h_lo; h < h_hi; ++h) {                                     // box {64 mn, 64 k} per 64-wide half
            if (mc) tma_load_2d_mc(map, bar, dst + h * (kTileBytes / 2), col0 + mn0 + 64 * h, row0 + kg, 3);
            else tma_load_2d(map, bar, dst + h * (kTileBytes / 2), col0 + mn0 + 64 * h, row0 + kg);
        }
    }
}

// A-operand tile of (t, k block kg) -> L2, same addressing as load_operand's A cases
__device__ __forceinline__ void prefetch_a(const CUtensorMap* map, const OperandMap& o, const TileCoord& t, int kg) {
    if (o.conv) {
        if (o.mn_major) return;
        const int cblocks = o.C / 64, hw = o.H * o.W;
        const int img = t.m0 / hw, y0 = (t.m0 - img * hw) / o.W;
        const int kb = kg / 64, tap = kb / cblocks, cb = kb - tap * cblocks;
        int dy, dx;
        tap_offset(tap, o.taps, dy, dx);
        tma_prefetch_4d(map, cb * 64, dx, y0 + dy, img);
        return;
    }
    const int col0 = o.col_base + t.bi * o.col_inner;
    const int row0 = o.row_outer * t.bo + o.row_inner * t.bi;
    if (!o.mn_major) {
        tma_prefetch_2d(map, col0 + kg, row0 + t.m0);
    } else {
        tma_prefetch_2d(map, col0 + t.m0, row0 + kg);
        tma_prefetch_2d(map, col0 + t.m0 + 64, row0 + kg);
    }
}

template <int BN, int MC>
__global__ void __launch_bounds__(kThreads, 1)
gemm_split_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                  const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                  const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_c_hi,
                  const __grid_constant__ CUtensorMap map_c_lo, const GemmParams p) {
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B tiles need 1024 B alignment
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int kStages = Plan<BN, MC>::kStages, kStageBytes = Plan<BN, MC>::kStageBytes, kBSlot = Plan<BN, MC>::kBSlot;
    constexpr bool kPair = MC == 4;
    unsigned char* store_bufs = smem + kStages * kStageBytes;
    constexpr int kStoreBufs = Plan<BN, MC>::kStoreBufs;
    uint64_t* bars = reinterpret_cast<uint64_t*>(store_bufs + Plan<BN, MC>::kStoreTotal);
    uint64_t* full = bars;                       // [kStages]
    uint64_t* empty = bars + kStages;            // [kStages]
    uint64_t* tmem_full = bars + 2 * kStages;    // [kAccStages]
    uint64_t* tmem_empty = tmem_full + kAccStages;
    uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(tmem_empty + kAccStages);

    constexpr int kTmemColsAlloc = kAccStages * BN;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta_rank = (MC != 1) ? (int)cluster_ctarank() : 0;
    // MC = 2: work items are tile pairs, distributed over clusters
    const int num_tiles = p.batch * p.splits * ((MC == 1) ? p.num_m : (p.num_m + 1) / 2) * p.num_n;
    const int work_first = (MC == 1) ? (int)blockIdx.x : (int)(blockIdx.x >> 1);
    const int work_stride = (MC == 1) ? (int)gridDim.x : (int)(gridDim.x >> 1);
    const int num_k = p.num_k;
    const bool a_lo_on = p.terms == 3 && !p.a_exact, b_lo_on = p.terms == 3 && !p.b_exact;   // which (hi, lo) pairs carry a lo half

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
        if (a_lo_on) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo) : "memory");
        if (b_lo_on) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_lo) : "memory");
        // MC=2: both CTAs' MMAs release a stage.  Pair: the leader's `full` collects one expect_tx arrival per CTA, its
        // `tmem_empty` the epilogue warps of both CTAs; `empty` / `tmem_full` get one multicast commit each.
        for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], kPair ? 2 : 1); mbar_init(&empty[i], MC == 2 ? 2 : 1); }
        for (int i = 0; i < kAccStages; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], kPair ? 2 * kEpiWarps : kEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM allocation is warp-collective; the same warp deallocates at the end
        if (kPair) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_ptr)),
                         "n"(kTmemColsAlloc) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_ptr)),
                         "n"(kTmemColsAlloc) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (MC != 1) cluster_sync_all();           // sibling barriers are initialised before any multicast can reach them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_ptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            // bytes THIS CTA contributes to a stage (pair: its A tile + its half of the B tile)
            const uint32_t tx = (a_lo_on ? 2 : 1) * kTileBytes + (b_lo_on ? 2 : 1) * (kPair ? BN * 64 : BN * 128);
            for (int tile = work_first; tile < num_tiles; tile += work_stride) {
                const TileCoord t = decode_tile<BN, MC>(tile, p, cta_rank);
                const bool pf = p.prefetch && tile + work_stride < num_tiles;
                const TileCoord tn = decode_tile<BN, MC>(pf ? tile + work_stride : tile, p, cta_rank);
                for (int kb = 0; kb < num_k; ++kb) {
                    const int kg = (t.s * num_k + kb) * BK;
                    if (pf && (tn.m0 != t.m0 || tn.b != t.b || tn.s != t.s)) {      // next tile's A block -> L2, one tile ahead
                        const int kgn = (tn.s * num_k + kb) * BK;
                        prefetch_a(&map_a_hi, p.a, tn, kgn);
                        if (a_lo_on) prefetch_a(&map_a_lo, p.a, tn, kgn);
                    }
                    mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char* st = smem + stage * kStageBytes;
                    uint32_t pair_bar = 0;
                    if (p.debug & 4) {
                        if (kPair) mbar_arrive_cluster(mapa_u32(smem_u32(&full[stage]), 0)); else mbar_arrive(&full[stage]);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    if (kPair) {
                        pair_bar = mapa_u32(smem_u32(&full[stage]), 0);     // the leader's barrier
                        mbar_expect_tx_cluster(pair_bar, tx);
                    } else {
                        mbar_expect_tx(&full[stage], tx);
                    }
                    const int brank = MC != 1 ? cta_rank : -1;
                    load_operand(&map_a_hi, &full[stage], st, p.a, t, t.m0, kg, BM, -1, pair_bar);
                    load_operand(&map_b_hi, &full[stage], st + 2 * kTileBytes, p.b, t, t.n0, kg, BN, brank, pair_bar);
                    if (a_lo_on) load_operand(&map_a_lo, &full[stage], st + kTileBytes, p.a, t, t.m0, kg, BM, -1, pair_bar);
                    if (b_lo_on)
                        load_operand(&map_b_lo, &full[stage], st + 2 * kTileBytes + kBSlot, p.b, t, t.n0, kg, BN, brank,
                                     pair_bar);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1 && (!kPair || cta_rank == 0)) {
        // ===================== MMA issuer (pair: the leader CTA only) =====================
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        const bool a_mn = p.a.mn_major != 0, b_mn = p.b.mn_major != 0;
        const uint32_t idesc = make_idesc(a_mn, b_mn, BN, kPair ? 2 * BM : BM);
        // descriptor start-address step (>>4) per UMMA_K: 32 B inside the swizzle row (K-major) or 16 k-rows (MN-major)
        const uint64_t a_step = a_mn ? (uint64_t)((UMMA_K * 128) >> 4) : (uint64_t)((UMMA_K * 2) >> 4);
        const uint64_t b_step = b_mn ? (uint64_t)((UMMA_K * 128) >> 4) : (uint64_t)((UMMA_K * 2) >> 4);
        for (int tile = work_first; tile < num_tiles; tile += work_stride, ++it) {
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
                    const uint64_t a_hi = make_desc(sa, a_mn), a_lo = make_desc(sa + kTileBytes, a_mn);
                    const uint64_t b_hi = make_desc(sa + 2 * kTileBytes, b_mn), b_lo = make_desc(sa + 2 * kTileBytes + kBSlot, b_mn);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        if (p.debug & 2) break;
                        const uint64_t ao = a_step * k, bo = b_step * k;
                        uint32_t accum = (kb | k) ? 1u : 0u;        // the tile's first MMA overwrites the accumulator
                        if (kPair) {
                            if (a_lo_on) { umma_pair(d_tmem, a_lo + ao, b_hi + bo, idesc, accum); accum = 1u; }
                            if (b_lo_on) { umma_pair(d_tmem, a_hi + ao, b_lo + bo, idesc, accum); accum = 1u; }
                            umma_pair(d_tmem, a_hi + ao, b_hi + bo, idesc, accum);
                        } else {
                            if (a_lo_on) { umma(d_tmem, a_lo + ao, b_hi + bo, idesc, accum); accum = 1u; }
                            if (b_lo_on) { umma(d_tmem, a_hi + ao, b_lo + bo, idesc, accum); accum = 1u; }
                            umma(d_tmem, a_hi + ao, b_hi + bo, idesc, accum);
                        }
                    }
                    if (kPair) {
                        umma_commit_pair(&empty[stage]);                        // both CTAs' producers may refill the stage
                        if (kb == num_k - 1) umma_commit_pair(&tmem_full[acc]); // both CTAs' epilogues may read their half
                    } else {
                        if (MC == 2) umma_commit_mc(&empty[stage], 3);     // the stage is shared: tell both producers
                        else umma_commit(&empty[stage]);                    // frees this smem stage when the MMAs retire
                        if (kb == num_k - 1) umma_commit(&tmem_full[acc]);  // accumulator complete
                    }
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 2) {
        // ===================== epilogue (warps 2..9) =====================
        // TMEM -> registers -> (alpha, +bias, ReLU) -> swizzled smem box -> TMA store: every global write is a full,
        // coalesced 128 B line issued by the copy engine; the staging box is double buffered per warp.
        const int q = warp & 3;                    // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;          // warps 2..5 take the left half of the tile's columns, 6..9 the right
        unsigned char* my_buf = store_bufs + (warp - 2) * kStoreBufs * kStoreBufBytes;
        int buf_sel = 0;                           // staging boxes alternate per store group
        int it = 0;
        for (int tile = work_first; tile < num_tiles; tile += work_stride, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const TileCoord t = decode_tile<BN, MC>(tile, p, cta_rank);
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const int row_in_batch = t.m0 + q * 32 + lane;
            const bool row_ok = t.valid && row_in_batch < p.M;
            const int c_row0 = (p.debug & 8) ? q * 32 : t.bo * p.c_row_outer + t.bi * p.c_row_inner + t.s * p.c_row_split + t.m0 + q * 32;
            const int c_col0 = (p.debug & 8) ? 0 : p.c_col_base + t.bi * p.c_col_inner + t.n0;
#pragma unroll 1
            for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
                uint32_t r[32];
                if (p.debug & 32) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = lane + j;
                } else {
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), r);
                }
                if (c0 == (half + 1) * (BN / 2) - 32) {   // this warp's share fully read: hand TMEM back to the MMA warp early
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (kPair) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));   // the leader's MMA warp waits
                        else mbar_arrive(&tmem_empty[acc]);
                    }
                }
                // The epilogue is issue-bound (one warp per scheduler owns a 32 x BN slab), so keep it at ~2 instructions
                // per element: bias comes in as 8 broadcast 16-byte loads, alpha/bias fold into one FFMA, ReLU is an
                // FMNMX against 0 or -inf, and the residual path is a separate warp-uniform branch.
                float v[32];
                if (p.bias) {
                    const float4* b4 = reinterpret_cast<const float4*>(p.bias + t.n0 + c0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 bb = __ldg(b4 + j);
                        v[4 * j] = bb.x; v[4 * j + 1] = bb.y; v[4 * j + 2] = bb.z; v[4 * j + 3] = bb.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0.f;
                }
                if (p.residual && row_ok) {
                    const float4* r4 = reinterpret_cast<const float4*>(p.residual + (int64_t)(c_row0 + lane) * p.ldc + c_col0 + c0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 rr = __ldg(r4 + j);
                        v[4 * j] += rr.x; v[4 * j + 1] += rr.y; v[4 * j + 2] += rr.z; v[4 * j + 3] += rr.w;
                    }
                }
                const float floor_v = p.relu ? 0.f : -3.402823466e38f;
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(fmaf(__uint_as_float(r[j]), p.alpha, v[j]), floor_v);
                if (p.mask && row_ok) {
                    // dX of a layer whose input is another layer's ReLU output: multiply by that ReLU's derivative here (the
                    // sign of the saved bf16 output) instead of in a separate pass over the gradient
                    const uint4* m4 = reinterpret_cast<const uint4*>(p.mask + (int64_t)(c_row0 + lane) * p.ldc + c_col0 + c0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 mm = __ldg(m4 + j);
                        const uint32_t w[4] = {mm.x, mm.y, mm.z, mm.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (!(w[e] & 0x00007fffu)) v[8 * j + 2 * e] = 0.f;
                            if (!(w[e] & 0x7fff0000u)) v[8 * j + 2 * e + 1] = 0.f;
                        }
                    }
                }
                if (p.colsum) {
                    // column sums of this 32 x 32 block without leaving the register file: a 5-stage butterfly in which every
                    // lane keeps the half of the columns that matches its lane bit and ships the other half (16 + 8 + 4 + 2 + 1 =
                    // 31 shuffles); lane l ends up with the sum of column l over the warp's 32 rows
                    float cs[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float lo_h = row_ok ? v[j] : 0.f, hi_h = row_ok ? v[j + 16] : 0.f;
                        const bool up = lane & 16;
                        cs[j] = (up ? hi_h : lo_h) + __shfl_xor_sync(0xffffffffu, up ? lo_h : hi_h, 16);
                    }
#define DSB_COLSUM_STAGE(W)                                                                           \
                    _Pragma("unroll") for (int j = 0; j < (W); ++j) {                                 \
                        const bool up = lane & (W);                                                   \
                        const float keep = up ? cs[j + (W)] : cs[j], send = up ? cs[j] : cs[j + (W)]; \
                        cs[j] = keep + __shfl_xor_sync(0xffffffffu, send, (W));                       \
                    }
                    DSB_COLSUM_STAGE(8) DSB_COLSUM_STAGE(4) DSB_COLSUM_STAGE(2) DSB_COLSUM_STAGE(1)
#undef DSB_COLSUM_STAGE
                    if (t.valid) atomicAdd(p.colsum + c_col0 + c0 + lane, cs[0]);
                }
                if (p.debug & 16) {
                    float acc_sink = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc_sink += v[j];
                    if (acc_sink == 1.2345e-30f && p.c_hi) p.c_hi[0] = __float2bfloat16(acc_sink);   // keeps the arithmetic alive
                    continue;
                }
                if (p.store_c) {
                    unsigned char* buf = my_buf + buf_sel * kStoreBufBytes;
                    buf_sel = (buf_sel + 1) % kStoreBufs;
                    // the TMA store that last read this staging box must have finished reading it (the sibling warp on
                    // this scheduler keeps issuing meanwhile)
                    if (lane == 0) tma_store_wait_read<kStoreBufs - 1>();
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
                    if (lane == 0 && t.valid && !(p.debug & 1)) {
                        if (p.accumulate) tma_reduce_add_2d(&map_c, buf, c_col0 + c0, c_row0);
                        else tma_store_2d(&map_c, buf, c_col0 + c0, c_row0);
                    }
                }
                if (p.c_hi) {
                    // bf16 (hi, lo) pair of the same 32 x 32 block: two 2 KiB SWIZZLE_64B boxes in the same staging slot,
                    // written by the copy engine as full lines (the consumer GEMM reads them straight back through TMA)
                    uint32_t h[16], l[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const float x0 = v[2 * e], x1 = v[2 * e + 1];
                        const __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
                        const float2 hf = __bfloat1622float2(hh);
                        const __nv_bfloat162 ll = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
                        h[e] = *reinterpret_cast<const uint32_t*>(&hh);
                        l[e] = *reinterpret_cast<const uint32_t*>(&ll);
                    }
                    unsigned char* buf = my_buf + buf_sel * kStoreBufBytes;
                    buf_sel = (buf_sel + 1) % kStoreBufs;
                    if (lane == 0) tma_store_wait_read<kStoreBufs - 1>();
                    __syncwarp();
                    const uint32_t rowbase = smem_u32(buf) + lane * 64;
                    const uint32_t sw = (uint32_t)((lane >> 1) & 3);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t addr = rowbase + ((j ^ sw) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(h[4 * j]), "r"(h[4 * j + 1]),
                                     "r"(h[4 * j + 2]), "r"(h[4 * j + 3]) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + 2048), "r"(l[4 * j]),
                                     "r"(l[4 * j + 1]), "r"(l[4 * j + 2]), "r"(l[4 * j + 3]) : "memory");
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0 && t.valid && !(p.debug & 1)) {
                        tma_store_2d(&map_c_hi, buf, c_col0 + c0, c_row0, false);       // one group for the pair
                        tma_store_2d(&map_c_lo, buf + 2048, c_col0 + c0, c_row0);
                    }
                }
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (MC != 1) cluster_sync_all();           // no CTA exits while a sibling may still multicast into / read from it
    if (warp == 1) {
        tc_fence_after();
        if (kPair)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemColsAlloc) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemColsAlloc) : "memory");
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

// 2-D row-major [rows, cols] tensor of `esize`-byte elements, box = [box_rows, box_cols], 128 B swizzle
int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int esize, int box_rows, int box_cols,
             CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { dsb::set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return DSB_ERR_CUDA; }
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * esize};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                     esize == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        dsb::set_error("cuTensorMapEncodeTiled failed: %d (rows %lld cols %lld esize %d)", (int)r, (long long)rows,
                       (long long)cols, esize);
        return DSB_ERR_CUDA;
    }
    return DSB_OK;
}

// 4-D NHWC bf16 tensor [imgs, H, W, C] (C contiguous), box = {64 c, W, box_h, 1}, 128 B swizzle, zero OOB fill
int make_map_nhwc(CUtensorMap* map, const void* base, int64_t imgs, int H, int W, int C, int box_h) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { dsb::set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return DSB_ERR_CUDA; }
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)imgs};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)W, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        dsb::set_error("cuTensorMapEncodeTiled(NHWC) failed: %d (imgs %lld H %d W %d C %d box_h %d)", (int)r,
                       (long long)imgs, H, W, C, box_h);
        return DSB_ERR_CUDA;
    }
    return DSB_OK;
}

template <int BN, int MC>
int launch_bn(const dsb_gemm_args& g, cudaStream_t stream) {
    const int batch = g.batch > 0 ? g.batch : 1, splits = g.splits > 0 ? g.splits : 1, inner = g.inner > 0 ? g.inner : 1;
    DSB_REQUIRE(g.m >= 0 && g.n > 0 && g.k > 0 && g.n % BN == 0 && g.k % (BK * splits) == 0,
                "gemm: need n %% %d == 0 and k %% (%d*splits) == 0 (m=%lld n=%d k=%d splits=%d)", BN, BK, (long long)g.m,
                g.n, g.k, splits);
    DSB_REQUIRE(batch == 1 || g.m % BM == 0, "gemm: batched problems need m %% %d == 0", BM);
    DSB_REQUIRE(splits == 1 || g.c_accumulate || g.c_row_split % BM == 0, "gemm: split-K needs c_row_split %% %d == 0", BM);
    DSB_REQUIRE(splits == 1 || (!g.bias && !g.relu && !g.residual), "gemm: split-K partial sums take no bias / ReLU / residual");
    DSB_REQUIRE(g.c_cols % 4 == 0, "gemm: C row pitch must be a 16-byte multiple");
    if (g.m == 0) return DSB_OK;
    const int num_m = (int)((g.m + BM - 1) / BM), num_n = g.n / BN;
    const int64_t tiles = (int64_t)batch * splits * num_m * num_n;
    DSB_REQUIRE(tiles < (1ll << 31), "gemm: too many tiles");
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo, mc, mc_hi, mc_lo;
    int rc;
    const void* a_lo = (g.terms == 3 && !g.a_exact) ? g.a_lo : g.a_hi;
    const void* b_lo = (g.terms == 3 && !g.b_exact) ? g.b_lo : g.b_hi;
    if (g.a_conv) {
        DSB_REQUIRE(!g.a_mn && batch == 1 && splits == 1, "gemm: conv A operand is K-major, unbatched");
        DSB_REQUIRE(g.conv_c % 64 == 0 && (g.conv_taps == 1 || g.conv_taps == 9) && g.conv_w <= 128 && 128 % g.conv_w == 0 &&
                    g.conv_h % (128 / g.conv_w) == 0, "gemm: unsupported conv geometry H=%d W=%d C=%d taps=%d", g.conv_h,
                    g.conv_w, g.conv_c, g.conv_taps);
        DSB_REQUIRE(g.k == g.conv_taps * g.conv_c && g.m == (int64_t)g.conv_imgs * g.conv_h * g.conv_w,
                    "gemm: conv A dims mismatch (m=%lld k=%d)", (long long)g.m, g.k);
        if ((rc = make_map_nhwc(&ma_hi, g.a_hi, g.conv_imgs, g.conv_h, g.conv_w, g.conv_c, 128 / g.conv_w))) return rc;
        if ((rc = make_map_nhwc(&ma_lo, a_lo, g.conv_imgs, g.conv_h, g.conv_w, g.conv_c, 128 / g.conv_w))) return rc;
    } else {
        DSB_REQUIRE(g.a_cols % 8 == 0, "gemm: A row pitch must be a 16-byte multiple");
        const int a_br = g.a_mn ? 64 : BM;
        if ((rc = make_map(&ma_hi, g.a_hi, g.a_rows, g.a_cols, 2, a_br, 64))) return rc;
        if ((rc = make_map(&ma_lo, a_lo, g.a_rows, g.a_cols, 2, a_br, 64))) return rc;
    }
    if (g.b_conv) {
        DSB_REQUIRE(g.b_mn && batch == 1, "gemm: conv B operand (weight gradient) is MN-major, unbatched");
        DSB_REQUIRE(g.conv_c % 64 == 0 && (g.conv_taps == 1 || g.conv_taps == 9) && g.conv_w <= 64 && 64 % g.conv_w == 0 &&
                    g.conv_h % (64 / g.conv_w) == 0, "gemm: unsupported conv geometry H=%d W=%d C=%d taps=%d", g.conv_h,
                    g.conv_w, g.conv_c, g.conv_taps);
        DSB_REQUIRE(g.n == g.conv_taps * g.conv_c && g.k == (int64_t)g.conv_imgs * g.conv_h * g.conv_w,
                    "gemm: conv B dims mismatch (n=%d k=%d)", g.n, g.k);
        if ((rc = make_map_nhwc(&mb_hi, g.b_hi, g.conv_imgs, g.conv_h, g.conv_w, g.conv_c, 64 / g.conv_w))) return rc;
        if ((rc = make_map_nhwc(&mb_lo, b_lo, g.conv_imgs, g.conv_h, g.conv_w, g.conv_c, 64 / g.conv_w))) return rc;
    } else {
        DSB_REQUIRE(g.b_cols % 8 == 0, "gemm: B row pitch must be a 16-byte multiple");
        const int b_br = g.b_mn ? 64 : (MC == 1 ? BN : BN / 2);     // cluster modes: each CTA loads half of the B rows
        if ((rc = make_map(&mb_hi, g.b_hi, g.b_rows, g.b_cols, 2, b_br, 64))) return rc;
        if ((rc = make_map(&mb_lo, b_lo, g.b_rows, g.b_cols, 2, b_br, 64))) return rc;
    }
    if (g.c_hi) {
        DSB_REQUIRE(g.c_cols % 8 == 0, "gemm: (hi, lo) output row pitch must be a 16-byte multiple");
        if ((rc = make_map(&mc_hi, g.c_hi, g.c_rows, g.c_cols, 2, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
        if ((rc = make_map(&mc_lo, g.c_lo, g.c_rows, g.c_cols, 2, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    }
    if (g.c) {
        if ((rc = make_map(&mc, g.c, g.c_rows, g.c_cols, 4, 32, 32))) return rc;
    } else {
        mc = mc_hi;                           // never dereferenced (store_c = 0)
    }
    if (!g.c_hi) mc_hi = mc_lo = mc;
    static bool configured = false;
    static int num_sms = 0;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_split_kernel<BN, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Plan<BN, MC>::kSmemBytes);
        if (e != cudaSuccess) { dsb::set_error("gemm smem attr: %s", cudaGetErrorString(e)); return DSB_ERR_CUDA; }
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        configured = true;
    }
    GemmParams p;
    p.bias = g.bias; p.residual = g.residual; p.mask = (const __nv_bfloat16*)g.relu_mask; p.colsum = g.colsum; p.c_hi = (__nv_bfloat16*)g.c_hi; p.c_lo = (__nv_bfloat16*)g.c_lo;
    p.alpha = g.alpha; p.M = g.m; p.ldc = g.c_cols; p.N = g.n; p.terms = g.terms; p.relu = g.relu;
    p.accumulate = g.c_accumulate; p.store_c = g.c != nullptr;
    p.a_exact = g.a_exact; p.b_exact = g.b_exact;
    static const int debug_bits = getenv("DSB_GEMM_DEBUG") ? atoi(getenv("DSB_GEMM_DEBUG")) : 0;
    p.debug = debug_bits;
    static const int no_prefetch = getenv("DSB_GEMM_NO_PREFETCH") ? atoi(getenv("DSB_GEMM_NO_PREFETCH")) : 0;   // DEV ONLY (A/B)
    p.prefetch = !no_prefetch && g.terms == 1;   // measured: -21 % time for 1-term products; 3-term ones get slower (CTA pair: FFN up 175 -> 184 us, FFN down 165 -> 208)
    p.num_m = num_m; p.num_n = num_n; p.num_k = g.k / (BK * splits);
    p.batch = batch; p.inner = inner; p.splits = splits;
    p.c_row_outer = g.c_row_outer; p.c_row_inner = g.c_row_inner; p.c_row_split = g.c_row_split;
    p.c_col_base = g.c_col_base; p.c_col_inner = g.c_col_inner;
    p.a = OperandMap{g.a_col_base, g.a_col_inner, g.a_row_outer, g.a_row_inner, g.a_mn, g.a_conv, g.conv_h, g.conv_w,
                     g.conv_c, g.conv_taps};
    p.b = OperandMap{g.b_col_base, g.b_col_inner, g.b_row_outer, g.b_row_inner, g.b_mn, g.b_conv, g.conv_h, g.conv_w,
                     g.conv_c, g.conv_taps};
    if (MC == 1) {
        const unsigned grid = (unsigned)(tiles < num_sms ? tiles : num_sms);
        gemm_split_kernel<BN, 1><<<grid, kThreads, Plan<BN, 1>::kSmemBytes, stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, mc, mc_hi, mc_lo, p);
    } else {
        // clusters of two CTAs; each cluster walks pairs of vertically adjacent tiles
        const int64_t pairs = (int64_t)batch * splits * ((num_m + 1) / 2) * num_n;
        const int clusters = (int)(pairs < num_sms / 2 ? pairs : num_sms / 2);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * clusters);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = Plan<BN, MC>::kSmemBytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_split_kernel<BN, MC>, ma_hi, ma_lo, mb_hi, mb_lo, mc, mc_hi, mc_lo, p);
        if (e != cudaSuccess) { dsb::set_error("gemm cluster launch: %s", cudaGetErrorString(e)); return DSB_ERR_CUDA; }
    }
    return dsb::check_launch("gemm_split");
}

int launch(const dsb_gemm_args& g, cudaStream_t stream) {
    DSB_REQUIRE(g.a_hi && g.b_hi && (g.c || g.c_hi), "gemm: null pointer");
    DSB_REQUIRE(g.c || (!g.c_accumulate && (g.splits <= 1)), "gemm: pair-only output takes no split-K / accumulate");
    DSB_REQUIRE(g.terms == 1 || g.terms == 3, "gemm: terms must be 1 or 3");
    DSB_REQUIRE(g.terms == 1 || ((g.a_lo || g.a_exact) && (g.b_lo || g.b_exact)), "gemm: terms=3 needs the lo halves (or a_exact / b_exact)");
    DSB_REQUIRE(!g.c_hi == !g.c_lo, "gemm: c_hi and c_lo go together");
    int bn = g.bn;
    if (!bn) {
        // wide tiles halve the A-operand traffic per flop (the kernel is L2->smem bound at 128x128: ncu shows 8.6 GB of
        // operand traffic for a 0.5 GB A matrix); use them when the problem still fills the machine
        const int64_t batch = g.batch > 0 ? g.batch : 1, splits = g.splits > 0 ? g.splits : 1;
        const int64_t tiles256 = (g.n % 256 == 0) ? batch * splits * ((g.m + BM - 1) / BM) * (g.n / 256) : 0;
        bn = tiles256 >= 148 ? 256 : (g.n % 128 == 0 ? 128 : 64);
    }
    DSB_REQUIRE(bn == 64 || bn == 128 || bn == 256, "gemm: bn must be 64, 128 or 256");
    // Cluster modes.  mc = 2 (TMA multicast of the B tile to two CTAs) halves the B traffic out of L2: measured +5 % for 1-term
    // products, nothing for 3-term ones.  mc = 4 (CTA pair, cta_group::2 MMAs, 256 x 256 tile over two SMs) cuts the operand
    // bytes per SM and the shared-memory reads per MMA by a third: measured -10..-17 % on the plain 3-term products with one
    // output form at M = 135168 (FFN up 19.0 -> 16.7 ms per step, FFN down 9.1 -> 7.5, the dX products 9.1 -> 8.0 / 7.4 -> 6.5),
    // +19 % on the batched attention scores (two k blocks per tile: the pair's barrier round trips do not amortise) and +12 %
    // on the dual-output QKV projection; implicit-GEMM convolutions have N <= 128.  DSB_GEMM_PAIR=0 (dev A/B) turns it off.
    int mcast = g.mc;
    if (!mcast) {
        const int64_t batch = g.batch > 0 ? g.batch : 1, splits = g.splits > 0 ? g.splits : 1;
        const int64_t num_m = (g.m + BM - 1) / BM;
        const int64_t pairs = batch * splits * ((num_m + 1) / 2) * (g.n / bn);
        mcast = (g.terms == 1 && bn >= 128 && num_m >= 2 && pairs >= 74) ? 2 : 1;
        static const int pair_mode = getenv("DSB_GEMM_PAIR") ? atoi(getenv("DSB_GEMM_PAIR")) : 1;
        if (pair_mode && g.terms == 3 && bn == 256 && batch == 1 && splits == 1 && pairs >= 74 && !g.a_conv && !g.b_conv &&
            !(g.c && g.c_hi))
            mcast = 4;
    }
    DSB_REQUIRE(mcast == 1 || ((mcast == 2 || mcast == 4) && bn >= 128), "gemm: mc must be 1, 2 or 4 (2 and 4 need bn >= 128)");
    if (mcast == 4) return bn == 256 ? launch_bn<256, 4>(g, stream) : launch_bn<128, 4>(g, stream);
    if (mcast == 2) return bn == 256 ? launch_bn<256, 2>(g, stream) : launch_bn<128, 2>(g, stream);
    if (bn == 256) return launch_bn<256, 1>(g, stream);
    return bn == 128 ? launch_bn<128, 1>(g, stream) : launch_bn<64, 1>(g, stream);
}

}  // namespace

extern "C" int dsb_gemm_ex(const dsb_gemm_args* args, dsb_stream_t stream) {
    DSB_REQUIRE(args, "gemm_ex: null args");
    return launch(*args, (cudaStream_t)stream);
}

extern "C" int dsb_gemm_bf16_split(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo,
                                   const float* bias, float* c, void* c_hi, void* c_lo, int64_t M, int N, int K,
                                   int terms, int relu, dsb_stream_t stream) {
    dsb_gemm_args g = {};
    g.a_hi = a_hi; g.a_lo = a_lo; g.b_hi = w_hi; g.b_lo = w_lo;
    g.a_rows = M; g.a_cols = K; g.b_rows = N; g.b_cols = K;
    g.bias = bias; g.alpha = 1.0f; g.relu = relu; g.terms = terms;
    g.c = c; g.c_rows = M; g.c_cols = N; g.c_hi = c_hi; g.c_lo = c_lo;
    g.m = M; g.n = N; g.k = K; g.batch = 1; g.inner = 1; g.splits = 1;
    return launch(g, (cudaStream_t)stream);
}
