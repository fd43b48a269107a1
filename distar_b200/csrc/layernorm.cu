// Fused (residual +) LayerNorm forward / backward, and the fused LayerNorm-LSTM cell.
//
// layernorm: y = LN(x [+ residual]) * gamma + beta  — replaces the `x + a` add and nn.LayerNorm launches of
// TransformerLayer (model/module_utils.py:130-139), ResFCBlock / ResFCBlock2 (ctools/torch_utils/network/res_block.py:
// 68-141) and the LSTM gate norms (model/lstm.py:142-143).  One warp per row, the row lives in registers
// (D = 128*VEC), y leaves as fp32 plus (optionally) the bf16 (hi, lo) pair the next tensor-core GEMM consumes, so no
// separate split pass runs.  Backward: one warp per row for dx, per-block partial sums for dgamma / dbeta.
//
// lstm cell: LayerNormLSTMCell (model/lstm.py:138-153) after the two matmuls: gates = ig + LN_h(hg_raw);
// c' = LN_c(sigmoid(f)*c + sigmoid(i)*tanh(g)); h' = sigmoid(o)*tanh(c').  One warp per batch row; ~12 ATen launches
// per (layer, timestep) become one.  (The backward of the cell is composed from these pieces in ops.py.)
#include "common.cuh"

namespace {

constexpr int kWarps = 4;

__device__ __forceinline__ void store_split4(__nv_bfloat16* hi, __nv_bfloat16* lo, int64_t off, float4 v) {
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
    const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y);
    const __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
    *reinterpret_cast<uint2*>(hi + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    *reinterpret_cast<uint2*>(lo + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
}

template <int VEC>
__global__ void __launch_bounds__(kWarps * 32)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ gamma,
              const float* __restrict__ beta, float* __restrict__ sum_out, float* __restrict__ y,
              __nv_bfloat16* __restrict__ y_hi, __nv_bfloat16* __restrict__ y_lo, float* __restrict__ stats,
              int64_t rows, float eps) {
    constexpr int D = 128 * VEC;
    const int64_t r = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int lane = threadIdx.x & 31;
    float4 v[VEC];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const int64_t off = r * D + i * 128 + lane * 4;
        float4 a = *reinterpret_cast<const float4*>(x + off);
        if (res) {
            const float4 b = *reinterpret_cast<const float4*>(res + off);
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            if (sum_out) *reinterpret_cast<float4*>(sum_out + off) = a;
        }
        v[i] = a;
        s += (a.x + a.y) + (a.z + a.w);
    }
    const float mean = dsb::warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
        q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
    const float rstd = rsqrtf(dsb::warp_sum(q) * (1.0f / D) + eps);
    if (lane == 0 && stats) { stats[2 * r] = mean; stats[2 * r + 1] = rstd; }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const int c = i * 128 + lane * 4;
        const float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
        float4 o;
        o.x = (v[i].x - mean) * rstd * g.x + b.x;
        o.y = (v[i].y - mean) * rstd * g.y + b.y;
        o.z = (v[i].z - mean) * rstd * g.z + b.z;
        o.w = (v[i].w - mean) * rstd * g.w + b.w;
        *reinterpret_cast<float4*>(y + r * D + c) = o;
        if (y_hi) store_split4(y_hi, y_lo, r * D + c, o);
    }
}

// dx = rstd * (g*gamma - mean(g*gamma) - xhat * mean(g*gamma*xhat));  partial dgamma/dbeta per block
template <int VEC>
__global__ void __launch_bounds__(kWarps * 32)
ln_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ xin, const float* __restrict__ gamma,
              const float* __restrict__ stats, float* __restrict__ gx, float* __restrict__ pgamma,
              float* __restrict__ pbeta, int64_t rows, int rows_per_block, int atomic) {
    constexpr int D = 128 * VEC;
    __shared__ float sg[kWarps][D];
    __shared__ float sb[kWarps][D];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 ag[VEC], ab[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) { ag[i] = make_float4(0, 0, 0, 0); ab[i] = make_float4(0, 0, 0, 0); }
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    for (int64_t r = r0 + warp; r < r0 + rows_per_block && r < rows; r += kWarps) {
        const float mean = stats[2 * r], rstd = stats[2 * r + 1];
        float4 g[VEC], xh[VEC];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const int c = i * 128 + lane * 4;
            const float4 go = *reinterpret_cast<const float4*>(gy + r * D + c);
            const float4 xv = *reinterpret_cast<const float4*>(xin + r * D + c);
            const float4 gm = *reinterpret_cast<const float4*>(gamma + c);
            xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
            g[i] = make_float4(go.x * gm.x, go.y * gm.y, go.z * gm.z, go.w * gm.w);
            s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
            s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
            ag[i].x += go.x * xh[i].x; ag[i].y += go.y * xh[i].y; ag[i].z += go.z * xh[i].z; ag[i].w += go.w * xh[i].w;
            ab[i].x += go.x; ab[i].y += go.y; ab[i].z += go.z; ab[i].w += go.w;
        }
        s1 = dsb::warp_sum(s1) * (1.0f / D);
        s2 = dsb::warp_sum(s2) * (1.0f / D);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const int c = i * 128 + lane * 4;
            float4 o;
            o.x = rstd * (g[i].x - s1 - xh[i].x * s2);
            o.y = rstd * (g[i].y - s1 - xh[i].y * s2);
            o.z = rstd * (g[i].z - s1 - xh[i].z * s2);
            o.w = rstd * (g[i].w - s1 - xh[i].w * s2);
            *reinterpret_cast<float4*>(gx + r * D + c) = o;
        }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        *reinterpret_cast<float4*>(&sg[warp][i * 128 + lane * 4]) = ag[i];
        *reinterpret_cast<float4*>(&sb[warp][i * 128 + lane * 4]) = ab[i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += kWarps * 32) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) { a += sg[w][c]; b += sb[w][c]; }
        if (atomic) {                 // pgamma / pbeta are the parameters' gradients themselves ([D])
            atomicAdd(pgamma + c, a);
            atomicAdd(pbeta + c, b);
        } else {
            pgamma[(int64_t)blockIdx.x * D + c] = a;
            pbeta[(int64_t)blockIdx.x * D + c] = b;
        }
    }
}

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + __expf(-x)); }

// H = hidden size = 128*HV (core LSTM: 384).  One warp per batch row.
// ig   [B, 4H]  = LN_i(x W_ih^T) (already normalised), hg [B, 4H] = raw h W_hh^T, c [B, H]
// outputs: h_out, c_out [B, H]; saved for backward: hgn [B,4H] (normalised hg), stats_h [B,2], pre_c [B,H], stats_c [B,2]
template <int HV>
__global__ void __launch_bounds__(kWarps * 32)
lstm_cell_fwd_kernel(const float* __restrict__ ig, const float* __restrict__ hg, const float* __restrict__ c_in,
                     const float* __restrict__ gh, const float* __restrict__ bh, const float* __restrict__ gc,
                     const float* __restrict__ bc, float* __restrict__ h_out, float* __restrict__ c_out,
                     float* __restrict__ gates_out, float* __restrict__ stats_h, float* __restrict__ pre_c,
                     float* __restrict__ stats_c, int B, float eps) {
    constexpr int H = 128 * HV, G = 4 * H;
    const int r = blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (r >= B) return;
    const int lane = threadIdx.x & 31;
    float4 v[4 * HV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4 * HV; ++i) {
        v[i] = *reinterpret_cast<const float4*>(hg + (int64_t)r * G + i * 128 + lane * 4);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = dsb::warp_sum(s) * (1.0f / G);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 4 * HV; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(dsb::warp_sum(q) * (1.0f / G) + eps);
    if (lane == 0) { stats_h[2 * r] = mean; stats_h[2 * r + 1] = rstd; }
    // gates = ig + LN_h(hg); keep the pre-activation gates for the backward pass
#pragma unroll
    for (int i = 0; i < 4 * HV; ++i) {
        const int c = i * 128 + lane * 4;
        const float4 g = *reinterpret_cast<const float4*>(gh + c), b = *reinterpret_cast<const float4*>(bh + c);
        const float4 a = *reinterpret_cast<const float4*>(ig + (int64_t)r * G + c);
        v[i].x = a.x + ((v[i].x - mean) * rstd * g.x + b.x);
        v[i].y = a.y + ((v[i].y - mean) * rstd * g.y + b.y);
        v[i].z = a.z + ((v[i].z - mean) * rstd * g.z + b.z);
        v[i].w = a.w + ((v[i].w - mean) * rstd * g.w + b.w);
        *reinterpret_cast<float4*>(gates_out + (int64_t)r * G + c) = v[i];
    }
    // chunk order: in, forget, cell, out (lstm.py:145)
    float4 pc[HV];
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < HV; ++i) {
        const float4 cin = *reinterpret_cast<const float4*>(c_in + (int64_t)r * H + i * 128 + lane * 4);
        const float4 gi = v[i], gf = v[HV + i], gg = v[2 * HV + i];
        pc[i].x = sigmoidf(gf.x) * cin.x + sigmoidf(gi.x) * tanhf(gg.x);
        pc[i].y = sigmoidf(gf.y) * cin.y + sigmoidf(gi.y) * tanhf(gg.y);
        pc[i].z = sigmoidf(gf.z) * cin.z + sigmoidf(gi.z) * tanhf(gg.z);
        pc[i].w = sigmoidf(gf.w) * cin.w + sigmoidf(gi.w) * tanhf(gg.w);
        *reinterpret_cast<float4*>(pre_c + (int64_t)r * H + i * 128 + lane * 4) = pc[i];
        s2 += (pc[i].x + pc[i].y) + (pc[i].z + pc[i].w);
    }
    const float mean_c = dsb::warp_sum(s2) * (1.0f / H);
    float q2 = 0.f;
#pragma unroll
    for (int i = 0; i < HV; ++i) {
        const float a = pc[i].x - mean_c, b = pc[i].y - mean_c, c = pc[i].z - mean_c, d = pc[i].w - mean_c;
        q2 += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd_c = rsqrtf(dsb::warp_sum(q2) * (1.0f / H) + eps);
    if (lane == 0) { stats_c[2 * r] = mean_c; stats_c[2 * r + 1] = rstd_c; }
#pragma unroll
    for (int i = 0; i < HV; ++i) {
        const int c = i * 128 + lane * 4;
        const float4 g = *reinterpret_cast<const float4*>(gc + c), b = *reinterpret_cast<const float4*>(bc + c);
        const float4 go = v[3 * HV + i];
        float4 cy, hy;
        cy.x = (pc[i].x - mean_c) * rstd_c * g.x + b.x;
        cy.y = (pc[i].y - mean_c) * rstd_c * g.y + b.y;
        cy.z = (pc[i].z - mean_c) * rstd_c * g.z + b.z;
        cy.w = (pc[i].w - mean_c) * rstd_c * g.w + b.w;
        hy.x = sigmoidf(go.x) * tanhf(cy.x);
        hy.y = sigmoidf(go.y) * tanhf(cy.y);
        hy.z = sigmoidf(go.z) * tanhf(cy.z);
        hy.w = sigmoidf(go.w) * tanhf(cy.w);
        *reinterpret_cast<float4*>(c_out + (int64_t)r * H + c) = cy;
        *reinterpret_cast<float4*>(h_out + (int64_t)r * H + c) = hy;
    }
}

// Backward of the fused cell.  Inputs: upstream gh [B,H] (wrt h_out), gcy [B,H] (wrt c_out), the saved pre-activation
// gates [B,4H], raw hg [B,4H] + stats_h, c_in, pre_c + stats_c.  Outputs: d_ig [B,4H] (gradient of the normalised input
// half = gradient of the gate pre-activations), d_hg [B,4H] (wrt the raw recurrent product), d_cin [B,H]; the four
// LayerNorm parameter gradients are accumulated with atomics into zero-initialised [4H] / [H] buffers.
template <int HV>
__global__ void __launch_bounds__(kWarps * 32)
lstm_cell_bwd_kernel(const float* __restrict__ gh, const float* __restrict__ gcy, const float* __restrict__ gates,
                     const float* __restrict__ hg, const float* __restrict__ stats_h, const float* __restrict__ c_in,
                     const float* __restrict__ pre_c, const float* __restrict__ stats_c,
                     const float* __restrict__ gam_h, const float* __restrict__ gam_c, const float* __restrict__ bet_c,
                     float* __restrict__ d_ig, float* __restrict__ d_hg, float* __restrict__ d_cin,
                     float* __restrict__ dgam_h, float* __restrict__ dbet_h, float* __restrict__ dgam_c,
                     float* __restrict__ dbet_c, int B) {
    constexpr int H = 128 * HV, G = 4 * H;
    const int r = blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (r >= B) return;
    const int lane = threadIdx.x & 31;
    const float mean_c = stats_c[2 * r], rstd_c = stats_c[2 * r + 1];
    // ---- through h = sigmoid(o) * tanh(cy), cy = LN_c(pc)
    float4 dcy[HV], xhc[HV], dgo[HV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < HV; ++i) {
        const int c = i * 128 + lane * 4;
        const float4 g_h = *reinterpret_cast<const float4*>(gh + (int64_t)r * H + c);
        const float4 g_c = gcy ? *reinterpret_cast<const float4*>(gcy + (int64_t)r * H + c) : make_float4(0, 0, 0, 0);
        const float4 pc = *reinterpret_cast<const float4*>(pre_c + (int64_t)r * H + c);
        const float4 gm = *reinterpret_cast<const float4*>(gam_c + c), bt = *reinterpret_cast<const float4*>(bet_c + c);
        const float4 go = *reinterpret_cast<const float4*>(gates + (int64_t)r * G + 3 * H + c);
        float xs[4] = {(pc.x - mean_c) * rstd_c, (pc.y - mean_c) * rstd_c, (pc.z - mean_c) * rstd_c, (pc.w - mean_c) * rstd_c};
        const float gms[4] = {gm.x, gm.y, gm.z, gm.w}, bts[4] = {bt.x, bt.y, bt.z, bt.w};
        const float ghs[4] = {g_h.x, g_h.y, g_h.z, g_h.w}, gcs[4] = {g_c.x, g_c.y, g_c.z, g_c.w};
        const float gos[4] = {go.x, go.y, go.z, go.w};
        float dcys[4], dgos[4], dgam[4], dbet[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float cy = xs[e] * gms[e] + bts[e];
            const float th = tanhf(cy), so = sigmoidf(gos[e]);
            dgos[e] = ghs[e] * th * so * (1.f - so);
            const float d = ghs[e] * so * (1.f - th * th) + gcs[e];
            dgam[e] = d * xs[e];
            dbet[e] = d;
            dcys[e] = d * gms[e];                    // gradient wrt xhat_c
            s1 += dcys[e];
            s2 += dcys[e] * xs[e];
        }
        dcy[i] = make_float4(dcys[0], dcys[1], dcys[2], dcys[3]);
        xhc[i] = make_float4(xs[0], xs[1], xs[2], xs[3]);
        dgo[i] = make_float4(dgos[0], dgos[1], dgos[2], dgos[3]);
#pragma unroll
        for (int e = 0; e < 4; ++e) { atomicAdd(dgam_c + c + e, dgam[e]); atomicAdd(dbet_c + c + e, dbet[e]); }
    }
    s1 = dsb::warp_sum(s1) * (1.0f / H);
    s2 = dsb::warp_sum(s2) * (1.0f / H);
    // ---- d_pc and the gate pre-activation gradients
    float4 dg[4 * HV];
#pragma unroll
    for (int i = 0; i < HV; ++i) {
        const int c = i * 128 + lane * 4;
        const float4 cin = *reinterpret_cast<const float4*>(c_in + (int64_t)r * H + c);
        const float4 gi = *reinterpret_cast<const float4*>(gates + (int64_t)r * G + c);
        const float4 gf = *reinterpret_cast<const float4*>(gates + (int64_t)r * G + H + c);
        const float4 gg = *reinterpret_cast<const float4*>(gates + (int64_t)r * G + 2 * H + c);
        const float dcs[4] = {dcy[i].x, dcy[i].y, dcy[i].z, dcy[i].w}, xs[4] = {xhc[i].x, xhc[i].y, xhc[i].z, xhc[i].w};
        const float cins[4] = {cin.x, cin.y, cin.z, cin.w}, gis[4] = {gi.x, gi.y, gi.z, gi.w};
        const float gfs[4] = {gf.x, gf.y, gf.z, gf.w}, ggs[4] = {gg.x, gg.y, gg.z, gg.w};
        float di[4], df[4], dgg[4], dci[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float dpc = rstd_c * (dcs[e] - s1 - xs[e] * s2);
            const float si = sigmoidf(gis[e]), sf = sigmoidf(gfs[e]), tg = tanhf(ggs[e]);
            di[e] = dpc * tg * si * (1.f - si);
            df[e] = dpc * cins[e] * sf * (1.f - sf);
            dgg[e] = dpc * si * (1.f - tg * tg);
            dci[e] = dpc * sf;
        }
        dg[i] = make_float4(di[0], di[1], di[2], di[3]);
        dg[HV + i] = make_float4(df[0], df[1], df[2], df[3]);
        dg[2 * HV + i] = make_float4(dgg[0], dgg[1], dgg[2], dgg[3]);
        dg[3 * HV + i] = dgo[i];
        *reinterpret_cast<float4*>(d_cin + (int64_t)r * H + c) = make_float4(dci[0], dci[1], dci[2], dci[3]);
    }
    // ---- d_ig = dg;  d_hg = LN_h backward of dg
    const float mean_h = stats_h[2 * r], rstd_h = stats_h[2 * r + 1];
    float t1 = 0.f, t2 = 0.f;
    float4 xh[4 * HV];
#pragma unroll
    for (int i = 0; i < 4 * HV; ++i) {
        const int c = i * 128 + lane * 4;
        *reinterpret_cast<float4*>(d_ig + (int64_t)r * G + c) = dg[i];
        const float4 hv = *reinterpret_cast<const float4*>(hg + (int64_t)r * G + c);
        const float4 gm = *reinterpret_cast<const float4*>(gam_h + c);
        xh[i] = make_float4((hv.x - mean_h) * rstd_h, (hv.y - mean_h) * rstd_h, (hv.z - mean_h) * rstd_h, (hv.w - mean_h) * rstd_h);
        atomicAdd(dgam_h + c + 0, dg[i].x * xh[i].x); atomicAdd(dgam_h + c + 1, dg[i].y * xh[i].y);
        atomicAdd(dgam_h + c + 2, dg[i].z * xh[i].z); atomicAdd(dgam_h + c + 3, dg[i].w * xh[i].w);
        atomicAdd(dbet_h + c + 0, dg[i].x); atomicAdd(dbet_h + c + 1, dg[i].y);
        atomicAdd(dbet_h + c + 2, dg[i].z); atomicAdd(dbet_h + c + 3, dg[i].w);
        dg[i] = make_float4(dg[i].x * gm.x, dg[i].y * gm.y, dg[i].z * gm.z, dg[i].w * gm.w);
        t1 += (dg[i].x + dg[i].y) + (dg[i].z + dg[i].w);
        t2 += (dg[i].x * xh[i].x + dg[i].y * xh[i].y) + (dg[i].z * xh[i].z + dg[i].w * xh[i].w);
    }
    t1 = dsb::warp_sum(t1) * (1.0f / G);
    t2 = dsb::warp_sum(t2) * (1.0f / G);
#pragma unroll
    for (int i = 0; i < 4 * HV; ++i) {
        const int c = i * 128 + lane * 4;
        float4 o;
        o.x = rstd_h * (dg[i].x - t1 - xh[i].x * t2);
        o.y = rstd_h * (dg[i].y - t1 - xh[i].y * t2);
        o.z = rstd_h * (dg[i].z - t1 - xh[i].z * t2);
        o.w = rstd_h * (dg[i].w - t1 - xh[i].w * t2);
        *reinterpret_cast<float4*>(d_hg + (int64_t)r * G + c) = o;
    }
}

template <int VEC>
int launch_ln_fwd(const float* x, const float* res, const float* gamma, const float* beta, float* sum_out, float* y,
                  void* y_hi, void* y_lo, float* stats, int64_t rows, float eps, cudaStream_t s) {
    const int64_t blocks = (rows + kWarps - 1) / kWarps;
    ln_fwd_kernel<VEC><<<(unsigned)blocks, kWarps * 32, 0, s>>>(x, res, gamma, beta, sum_out, y, (__nv_bfloat16*)y_hi,
                                                               (__nv_bfloat16*)y_lo, stats, rows, eps);
    return dsb::check_launch("layernorm_fwd");
}
template <int VEC>
int launch_ln_bwd(const float* gy, const float* xin, const float* gamma, const float* stats, float* gx, float* pg,
                  float* pb, int64_t rows, int rows_per_block, int blocks, int atomic, cudaStream_t s) {
    ln_bwd_kernel<VEC><<<(unsigned)blocks, kWarps * 32, 0, s>>>(gy, xin, gamma, stats, gx, pg, pb, rows, rows_per_block, atomic);
    return dsb::check_launch("layernorm_bwd");
}

}  // namespace

extern "C" int dsb_layernorm_supported(int D) { return (D == 128 || D == 256 || D == 384 || D == 512 || D == 1536) ? 1 : 0; }

extern "C" int dsb_layernorm_fwd(const float* x, const float* residual, const float* gamma, const float* beta,
                                 float* sum_out, float* y, void* y_hi, void* y_lo, float* stats, int64_t rows, int D,
                                 float eps, dsb_stream_t stream) {
    DSB_REQUIRE(x && gamma && beta && y && rows >= 0, "layernorm_fwd: bad argument");
    DSB_REQUIRE(!y_hi == !y_lo, "layernorm_fwd: y_hi and y_lo go together");
    DSB_REQUIRE(dsb_layernorm_supported(D), "layernorm_fwd: unsupported width %d", D);
    if (rows == 0) return DSB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    switch (D / 128) {
        case 1: return launch_ln_fwd<1>(x, residual, gamma, beta, sum_out, y, y_hi, y_lo, stats, rows, eps, s);
        case 2: return launch_ln_fwd<2>(x, residual, gamma, beta, sum_out, y, y_hi, y_lo, stats, rows, eps, s);
        case 3: return launch_ln_fwd<3>(x, residual, gamma, beta, sum_out, y, y_hi, y_lo, stats, rows, eps, s);
        case 4: return launch_ln_fwd<4>(x, residual, gamma, beta, sum_out, y, y_hi, y_lo, stats, rows, eps, s);
        default: return launch_ln_fwd<12>(x, residual, gamma, beta, sum_out, y, y_hi, y_lo, stats, rows, eps, s);
    }
}

extern "C" int dsb_layernorm_bwd_blocks(int64_t rows) {
    int64_t b = (rows + 63) / 64;
    if (b > 592) b = 592;
    if (b < 1) b = 1;
    return (int)b;
}

extern "C" int dsb_layernorm_bwd(const float* gy, const float* xin, const float* gamma, const float* stats, float* gx,
                                 float* pgamma, float* pbeta, int atomic, int64_t rows, int D, dsb_stream_t stream) {
    DSB_REQUIRE(gy && xin && gamma && stats && gx && pgamma && pbeta && rows > 0, "layernorm_bwd: bad argument");
    DSB_REQUIRE(dsb_layernorm_supported(D), "layernorm_bwd: unsupported width %d", D);
    const int blocks = dsb_layernorm_bwd_blocks(rows);
    const int rpb = (int)((rows + blocks - 1) / blocks);
    cudaStream_t s = (cudaStream_t)stream;
    switch (D / 128) {
        case 1: return launch_ln_bwd<1>(gy, xin, gamma, stats, gx, pgamma, pbeta, rows, rpb, blocks, atomic, s);
        case 2: return launch_ln_bwd<2>(gy, xin, gamma, stats, gx, pgamma, pbeta, rows, rpb, blocks, atomic, s);
        case 3: return launch_ln_bwd<3>(gy, xin, gamma, stats, gx, pgamma, pbeta, rows, rpb, blocks, atomic, s);
        case 4: return launch_ln_bwd<4>(gy, xin, gamma, stats, gx, pgamma, pbeta, rows, rpb, blocks, atomic, s);
        default: return launch_ln_bwd<12>(gy, xin, gamma, stats, gx, pgamma, pbeta, rows, rpb, blocks, atomic, s);
    }
}

extern "C" int dsb_lstm_cell_bwd(const float* gh, const float* gcy, const float* gates, const float* hg,
                                 const float* stats_h, const float* c_in, const float* pre_c, const float* stats_c,
                                 const float* gamma_h, const float* gamma_c, const float* beta_c, float* d_ig, float* d_hg,
                                 float* d_cin, float* dgamma_h, float* dbeta_h, float* dgamma_c, float* dbeta_c, int B,
                                 int H, dsb_stream_t stream) {
    DSB_REQUIRE(gh && gates && hg && stats_h && c_in && pre_c && stats_c && gamma_h && gamma_c && beta_c && d_ig && d_hg &&
                d_cin && dgamma_h && dbeta_h && dgamma_c && dbeta_c && B > 0, "lstm_cell_bwd: bad argument");
    DSB_REQUIRE(H == 384 || H == 128, "lstm_cell_bwd: hidden size must be 128 or 384 (got %d)", H);
    const int blocks = (B + kWarps - 1) / kWarps;
    cudaStream_t s = (cudaStream_t)stream;
    if (H == 384)
        lstm_cell_bwd_kernel<3><<<blocks, kWarps * 32, 0, s>>>(gh, gcy, gates, hg, stats_h, c_in, pre_c, stats_c, gamma_h,
                                                              gamma_c, beta_c, d_ig, d_hg, d_cin, dgamma_h, dbeta_h,
                                                              dgamma_c, dbeta_c, B);
    else
        lstm_cell_bwd_kernel<1><<<blocks, kWarps * 32, 0, s>>>(gh, gcy, gates, hg, stats_h, c_in, pre_c, stats_c, gamma_h,
                                                              gamma_c, beta_c, d_ig, d_hg, d_cin, dgamma_h, dbeta_h,
                                                              dgamma_c, dbeta_c, B);
    return dsb::check_launch("lstm_cell_bwd");
}

extern "C" int dsb_lstm_cell_fwd(const float* ig, const float* hg, const float* c_in, const float* gamma_h,
                                 const float* beta_h, const float* gamma_c, const float* beta_c, float* h_out,
                                 float* c_out, float* gates, float* stats_h, float* pre_c, float* stats_c, int B, int H,
                                 float eps, dsb_stream_t stream) {
    DSB_REQUIRE(ig && hg && c_in && gamma_h && beta_h && gamma_c && beta_c && h_out && c_out && gates && stats_h && pre_c &&
                stats_c && B > 0, "lstm_cell_fwd: bad argument");
    DSB_REQUIRE(H == 384 || H == 128, "lstm_cell_fwd: hidden size must be 128 or 384 (got %d)", H);
    const int blocks = (B + kWarps - 1) / kWarps;
    cudaStream_t s = (cudaStream_t)stream;
    if (H == 384)
        lstm_cell_fwd_kernel<3><<<blocks, kWarps * 32, 0, s>>>(ig, hg, c_in, gamma_h, beta_h, gamma_c, beta_c, h_out, c_out,
                                                              gates, stats_h, pre_c, stats_c, B, eps);
    else
        lstm_cell_fwd_kernel<1><<<blocks, kWarps * 32, 0, s>>>(ig, hg, c_in, gamma_h, beta_h, gamma_c, beta_c, h_out, c_out,
                                                              gates, stats_h, pre_c, stats_c, B, eps);
    return dsb::check_launch("lstm_cell_fwd");
}
