// Persistent LayerNorm-LSTM layer: ALL time steps of one layer in ONE launch each way.
//
// Reference: LSTMLayer.forward / LayerNormLSTMCell.forward, model/lstm.py:138-167 — a Python loop over the T+1 steps, each
// step ~14 launches (the h W_hh^T matmul, three LayerNorms, chunk, sigmoids ...).  Round 1 had it at one library fp32 GEMM
// + one fused cell kernel per (layer, step) each way: 99 + 99 launches forward, the same backward.  Here a CTA owns kRows
// batch rows for the whole sequence: h and c never leave the SM, the recurrent product is a register GEMV against W_hh
// streamed from L2 (2.4 MB, resident there), and both LayerNorms, the gates and the cell update follow in the same step.
// The input half LN_i(x W_ih^T) is one tensor-core GEMM + one LayerNorm launch over all steps (policy_net.Net.lstm); the
// W_hh gradient is one tensor-core GEMM over all (step, row) pairs (ops._LstmLayer.backward).
//
// Forward saves what the reference's autograd would: pre-activation gates, raw h W_hh^T and its LayerNorm statistics, the
// pre-LayerNorm cell state and its statistics.  Backward walks the steps in reverse inside the kernel, carries dh / dc in
// shared memory / registers and accumulates the LayerNorm parameter gradients in registers (one atomicAdd per parameter
// element and CTA at the end instead of one per step).
#include "common.cuh"

namespace {

constexpr int kRows = 2;                    // batch rows per CTA: B = 128 -> 64 CTAs

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + __expf(-x)); }

// sums of kRows values over the whole block (all threads get the result); red: [kRows][32] floats
template <int NW>
__device__ __forceinline__ void block_sum_rows(float (&v)[kRows], float* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < kRows; ++r) v[r] = dsb::warp_sum(v[r]);
    __syncthreads();                        // previous use of `red` is over
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < kRows; ++r) red[r * 32 + w] = v[r];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < NW; ++i) t += red[r * 32 + i];
        v[r] = t;
    }
}

// H = 128 * HV threads; thread u owns hidden unit u in the cell update and gate columns [4u, 4u+4) in the recurrent GEMV
template <int HV>
__global__ void __launch_bounds__(128 * HV)
lstm_seq_fwd_kernel(const float* __restrict__ ig_all, const float* __restrict__ h0, const float* __restrict__ c0,
                    const float* __restrict__ w_t /* [H, 4H] = W_hh^T */, const float* __restrict__ gam_h,
                    const float* __restrict__ bet_h, const float* __restrict__ gam_c, const float* __restrict__ bet_c,
                    float* __restrict__ hs, float* __restrict__ cs, float* __restrict__ gates_out,
                    float* __restrict__ hg_out, float* __restrict__ st_h, float* __restrict__ pre_c_out,
                    float* __restrict__ st_c, int L, int B, float eps) {
    constexpr int H = 128 * HV, G = 4 * H, NW = H / 32;
    __shared__ __align__(16) float h_s[kRows][H];
    __shared__ __align__(16) float g_s[kRows][G];
    __shared__ float red[kRows * 32];
    const int u = threadIdx.x;
    const int row0 = blockIdx.x * kRows;
    float c_prev[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const int b = min(row0 + r, B - 1);                 // a ragged last CTA recomputes row B-1 and does not store it
        h_s[r][u] = h0[(int64_t)b * H + u];
        c_prev[r] = c0[(int64_t)b * H + u];
    }
    const float4 gh4 = *reinterpret_cast<const float4*>(gam_h + 4 * u), bh4 = *reinterpret_cast<const float4*>(bet_h + 4 * u);
    const float gcu = gam_c[u], bcu = bet_c[u];
    __syncthreads();
    for (int t = 0; t < L; ++t) {
        // ---- recurrent product hg[r, 4u..4u+3] = sum_k h[r, k] W_hh[4u.., k]
        float4 acc[kRows];
#pragma unroll
        for (int r = 0; r < kRows; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* wp = reinterpret_cast<const float4*>(w_t) + u;
#pragma unroll 8
        for (int k = 0; k < H; ++k) {
            const float4 w = __ldg(wp + (int64_t)k * (G / 4));
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                const float hv = h_s[r][k];
                acc[r].x = fmaf(hv, w.x, acc[r].x); acc[r].y = fmaf(hv, w.y, acc[r].y);
                acc[r].z = fmaf(hv, w.z, acc[r].z); acc[r].w = fmaf(hv, w.w, acc[r].w);
            }
        }
        // ---- LayerNorm_h statistics over the 4H raw values of each row (two-pass, as nn.LayerNorm)
        float s[kRows];
#pragma unroll
        for (int r = 0; r < kRows; ++r) s[r] = (acc[r].x + acc[r].y) + (acc[r].z + acc[r].w);
        block_sum_rows<NW>(s, red);
        float mean[kRows], q[kRows];
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            mean[r] = s[r] * (1.0f / G);
            const float a = acc[r].x - mean[r], b = acc[r].y - mean[r], c = acc[r].z - mean[r], d = acc[r].w - mean[r];
            q[r] = (a * a + b * b) + (c * c + d * d);
        }
        block_sum_rows<NW>(q, red);
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            const int b = row0 + r;
            const float rstd = rsqrtf(q[r] * (1.0f / G) + eps);
            const bool live = b < B;
            const int64_t base = ((int64_t)t * B + (live ? b : B - 1)) * G + 4 * u;
            const float4 a = *reinterpret_cast<const float4*>(ig_all + base);
            float4 g;
            g.x = a.x + ((acc[r].x - mean[r]) * rstd * gh4.x + bh4.x);
            g.y = a.y + ((acc[r].y - mean[r]) * rstd * gh4.y + bh4.y);
            g.z = a.z + ((acc[r].z - mean[r]) * rstd * gh4.z + bh4.z);
            g.w = a.w + ((acc[r].w - mean[r]) * rstd * gh4.w + bh4.w);
            *reinterpret_cast<float4*>(&g_s[r][4 * u]) = g;
            if (live) {
                *reinterpret_cast<float4*>(hg_out + base) = acc[r];
                *reinterpret_cast<float4*>(gates_out + base) = g;
                if (u == 0) { st_h[((int64_t)t * B + b) * 2] = mean[r]; st_h[((int64_t)t * B + b) * 2 + 1] = rstd; }
            }
        }
        __syncthreads();
        // ---- cell update for hidden unit u (chunk order in, forget, cell, out: lstm.py:145)
        float pc[kRows], go[kRows];
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            const float gi = g_s[r][u], gf = g_s[r][H + u], gg = g_s[r][2 * H + u];
            go[r] = g_s[r][3 * H + u];
            pc[r] = sigmoidf(gf) * c_prev[r] + sigmoidf(gi) * tanhf(gg);
            s[r] = pc[r];
        }
        block_sum_rows<NW>(s, red);
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            mean[r] = s[r] * (1.0f / H);
            const float d = pc[r] - mean[r];
            q[r] = d * d;
        }
        block_sum_rows<NW>(q, red);            // (its leading barrier also orders the g_s reads above before the next step's writes)
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            const int b = row0 + r;
            const float rstd = rsqrtf(q[r] * (1.0f / H) + eps);
            const float cy = (pc[r] - mean[r]) * rstd * gcu + bcu;
            const float hy = sigmoidf(go[r]) * tanhf(cy);
            c_prev[r] = cy;
            h_s[r][u] = hy;
            if (b < B) {
                const int64_t o = ((int64_t)t * B + b) * H + u;
                hs[o] = hy;
                cs[o] = cy;
                pre_c_out[o] = pc[r];
                if (u == 0) { st_c[((int64_t)t * B + b) * 2] = mean[r]; st_c[((int64_t)t * B + b) * 2 + 1] = rstd; }
            }
        }
        __syncthreads();                       // h_s complete before the next step's GEMV reads it
    }
}

// Backward.  d_hg[t] feeds (a) the recurrent gradient dh[t-1] += d_hg[t] W_hh inside the kernel and (b) the W_hh gradient GEMM
// outside.  W is read in its own layout [4H, H]: the 4 thread groups of H/4 threads split the 4H rows, each thread carries
// 4 consecutive hidden columns (float4), partial sums meet in shared memory.
template <int HV>
__global__ void __launch_bounds__(128 * HV)
lstm_seq_bwd_kernel(const float* __restrict__ g_hs, const float* __restrict__ g_clast, const float* __restrict__ gates,
                    const float* __restrict__ hg_all, const float* __restrict__ st_h, const float* __restrict__ pre_c,
                    const float* __restrict__ st_c, const float* __restrict__ cs, const float* __restrict__ c0,
                    const float* __restrict__ w /* [4H, H] */, const float* __restrict__ gam_h,
                    const float* __restrict__ gam_c, const float* __restrict__ bet_c, float* __restrict__ d_ig,
                    float* __restrict__ d_hg, float* __restrict__ dh0, float* __restrict__ dc0,
                    float* __restrict__ dgam_h, float* __restrict__ dbet_h, float* __restrict__ dgam_c,
                    float* __restrict__ dbet_c, int L, int B) {
    constexpr int H = 128 * HV, G = 4 * H, NW = H / 32, Q = H / 4;    // Q threads per group
    __shared__ __align__(16) float dhg_s[kRows][G];
    __shared__ __align__(16) float part_s[4][kRows][H];
    __shared__ float red[kRows * 32];
    const int u = threadIdx.x;
    const int row0 = blockIdx.x * kRows;
    const int grp = u / Q, k4 = u - grp * Q;
    float dh_next[kRows], dc_next[kRows];
    float a_gh[4] = {0.f, 0.f, 0.f, 0.f}, a_bh[4] = {0.f, 0.f, 0.f, 0.f}, a_gc = 0.f, a_bc = 0.f;
    const float gcu = gam_c[u], bcu = bet_c[u];
    float gmh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) gmh[j] = gam_h[j * H + u];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const int b = row0 + r;
        dh_next[r] = 0.f;
        dc_next[r] = (g_clast && b < B) ? g_clast[(int64_t)b * H + u] : 0.f;
    }
    for (int t = L - 1; t >= 0; --t) {
        float dcy[kRows], xs[kRows], dgo[kRows], s1[kRows], s2[kRows];
        bool live[kRows];
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            const int b = row0 + r;
            live[r] = b < B;
            const int64_t o = ((int64_t)t * B + (live[r] ? b : B - 1)) * H + u;
            const float gh = (live[r] ? (g_hs ? g_hs[o] : 0.f) + dh_next[r] : 0.f);
            const float mean_c = st_c[((int64_t)t * B + (live[r] ? b : B - 1)) * 2], rstd_c = st_c[((int64_t)t * B + (live[r] ? b : B - 1)) * 2 + 1];
            xs[r] = (pre_c[o] - mean_c) * rstd_c;
            const float cy = xs[r] * gcu + bcu;
            const float th = tanhf(cy);
            const float so = sigmoidf(gates[((int64_t)t * B + (live[r] ? b : B - 1)) * G + 3 * H + u]);
            dgo[r] = gh * th * so * (1.f - so);
            const float d = live[r] ? gh * so * (1.f - th * th) + dc_next[r] : 0.f;
            a_gc += d * xs[r];
            a_bc += d;
            dcy[r] = d * gcu;
            s1[r] = dcy[r];
            s2[r] = dcy[r] * xs[r];
        }
        block_sum_rows<NW>(s1, red);
        block_sum_rows<NW>(s2, red);
        float dg[kRows][4], xh[kRows][4], t1[kRows], t2[kRows], rstd_h[kRows];
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            const int b = live[r] ? row0 + r : B - 1;
            const int64_t gb = ((int64_t)t * B + b) * G;
            const float rstd_c = st_c[((int64_t)t * B + b) * 2 + 1];
            const float dpc = rstd_c * (dcy[r] - s1[r] * (1.0f / H) - xs[r] * s2[r] * (1.0f / H));
            const float cin = t > 0 ? cs[((int64_t)(t - 1) * B + b) * H + u] : c0[(int64_t)b * H + u];
            const float gi = gates[gb + u], gf = gates[gb + H + u], gg = gates[gb + 2 * H + u];
            const float si = sigmoidf(gi), sf = sigmoidf(gf), tg = tanhf(gg);
            dg[r][0] = dpc * tg * si * (1.f - si);
            dg[r][1] = dpc * cin * sf * (1.f - sf);
            dg[r][2] = dpc * si * (1.f - tg * tg);
            dg[r][3] = dgo[r];
            dc_next[r] = dpc * sf;
            const float mean_h = st_h[((int64_t)t * B + b) * 2];
            rstd_h[r] = st_h[((int64_t)t * B + b) * 2 + 1];
            t1[r] = 0.f;
            t2[r] = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (!live[r]) dg[r][j] = 0.f;
                xh[r][j] = (hg_all[gb + j * H + u] - mean_h) * rstd_h[r];
                if (live[r]) d_ig[gb + j * H + u] = dg[r][j];
                a_gh[j] += dg[r][j] * xh[r][j];
                a_bh[j] += dg[r][j];
                const float dgs = dg[r][j] * gmh[j];
                dg[r][j] = dgs;
                t1[r] += dgs;
                t2[r] += dgs * xh[r][j];
            }
        }
        block_sum_rows<NW>(t1, red);
        block_sum_rows<NW>(t2, red);
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            const int b = live[r] ? row0 + r : B - 1;
            const int64_t gb = ((int64_t)t * B + b) * G;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float v = rstd_h[r] * (dg[r][j] - t1[r] * (1.0f / G) - xh[r][j] * t2[r] * (1.0f / G));
                dhg_s[r][j * H + u] = v;
                if (live[r]) d_hg[gb + j * H + u] = v;
            }
        }
        __syncthreads();
        // ---- dh[t-1][r, k] = sum_j d_hg[r, j] W_hh[j, k]: group `grp` takes rows j in [grp*H, (grp+1)*H), 4 columns per thread
        float4 acc[kRows];
#pragma unroll
        for (int r = 0; r < kRows; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* wp = reinterpret_cast<const float4*>(w) + (int64_t)grp * H * (H / 4) + k4;
#pragma unroll 8
        for (int j = 0; j < H; ++j) {
            const float4 wv = __ldg(wp + (int64_t)j * (H / 4));
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                const float dv = dhg_s[r][grp * H + j];
                acc[r].x = fmaf(dv, wv.x, acc[r].x); acc[r].y = fmaf(dv, wv.y, acc[r].y);
                acc[r].z = fmaf(dv, wv.z, acc[r].z); acc[r].w = fmaf(dv, wv.w, acc[r].w);
            }
        }
#pragma unroll
        for (int r = 0; r < kRows; ++r) *reinterpret_cast<float4*>(&part_s[grp][r][4 * k4]) = acc[r];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kRows; ++r) dh_next[r] = (part_s[0][r][u] + part_s[1][r][u]) + (part_s[2][r][u] + part_s[3][r][u]);
        // (the next iteration's first block_sum_rows barrier separates these reads from the next writes of part_s / dhg_s)
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const int b = row0 + r;
        if (b < B) {
            dh0[(int64_t)b * H + u] = dh_next[r];
            dc0[(int64_t)b * H + u] = dc_next[r];
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        atomicAdd(dgam_h + j * H + u, a_gh[j]);
        atomicAdd(dbet_h + j * H + u, a_bh[j]);
    }
    atomicAdd(dgam_c + u, a_gc);
    atomicAdd(dbet_c + u, a_bc);
}

}  // namespace

extern "C" int dsb_lstm_seq_fwd(const float* ig_all, const float* h0, const float* c0, const float* w_hh_t, const float* gamma_h,
                                const float* beta_h, const float* gamma_c, const float* beta_c, float* hs, float* cs,
                                float* gates, float* hg, float* stats_h, float* pre_c, float* stats_c, int L, int B, int H,
                                float eps, dsb_stream_t stream) {
    DSB_REQUIRE(ig_all && h0 && c0 && w_hh_t && gamma_h && beta_h && gamma_c && beta_c && hs && cs && gates && hg && stats_h &&
                pre_c && stats_c && L > 0 && B > 0, "lstm_seq_fwd: bad argument");
    DSB_REQUIRE(H == 384 || H == 128, "lstm_seq_fwd: hidden size must be 128 or 384 (got %d)", H);
    const int blocks = (B + kRows - 1) / kRows;
    cudaStream_t s = (cudaStream_t)stream;
    if (H == 384)
        lstm_seq_fwd_kernel<3><<<blocks, 384, 0, s>>>(ig_all, h0, c0, w_hh_t, gamma_h, beta_h, gamma_c, beta_c, hs, cs, gates, hg,
                                                      stats_h, pre_c, stats_c, L, B, eps);
    else
        lstm_seq_fwd_kernel<1><<<blocks, 128, 0, s>>>(ig_all, h0, c0, w_hh_t, gamma_h, beta_h, gamma_c, beta_c, hs, cs, gates, hg,
                                                      stats_h, pre_c, stats_c, L, B, eps);
    return dsb::check_launch("lstm_seq_fwd");
}

extern "C" int dsb_lstm_seq_bwd(const float* grad_hs, const float* grad_c_last, const float* gates, const float* hg,
                                const float* stats_h, const float* pre_c, const float* stats_c, const float* cs, const float* c0,
                                const float* w_hh, const float* gamma_h, const float* gamma_c, const float* beta_c, float* d_ig,
                                float* d_hg, float* dh0, float* dc0, float* dgamma_h, float* dbeta_h, float* dgamma_c,
                                float* dbeta_c, int L, int B, int H, dsb_stream_t stream) {
    DSB_REQUIRE(gates && hg && stats_h && pre_c && stats_c && cs && c0 && w_hh && gamma_h && gamma_c && beta_c && d_ig && d_hg &&
                dh0 && dc0 && dgamma_h && dbeta_h && dgamma_c && dbeta_c && L > 0 && B > 0, "lstm_seq_bwd: bad argument");
    DSB_REQUIRE(H == 384 || H == 128, "lstm_seq_bwd: hidden size must be 128 or 384 (got %d)", H);
    const int blocks = (B + kRows - 1) / kRows;
    cudaStream_t s = (cudaStream_t)stream;
    if (H == 384)
        lstm_seq_bwd_kernel<3><<<blocks, 384, 0, s>>>(grad_hs, grad_c_last, gates, hg, stats_h, pre_c, stats_c, cs, c0, w_hh,
                                                      gamma_h, gamma_c, beta_c, d_ig, d_hg, dh0, dc0, dgamma_h, dbeta_h, dgamma_c,
                                                      dbeta_c, L, B);
    else
        lstm_seq_bwd_kernel<1><<<blocks, 128, 0, s>>>(grad_hs, grad_c_last, gates, hg, stats_h, pre_c, stats_c, cs, c0, w_hh,
                                                      gamma_h, gamma_c, beta_c, d_ig, d_hg, dh0, dc0, dgamma_h, dbeta_h, dgamma_c,
                                                      dbeta_c, L, B);
    return dsb::check_launch("lstm_seq_bwd");
}
