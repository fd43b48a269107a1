// Fused V-trace / UPGO / TD(lambda) backward-in-time scans — replaces the python `for t in reversed(range(T))`
// loops of rl_training/as_rl_utils.py:157-192 (multistep_forward_view), :265-281 (upgo_returns) and :284-312
// (vtrace_advantages), which the reference re-runs per head x per field (2-8 k micro-launches per step).
// One thread per (batch column b, work item); the whole T recursion stays in registers.  Arithmetic order
// follows the reference expression by expression so results match the CPU oracle to the last bit where the
// compiler does not contract (compiled with -fmad=false for this file's kernels via explicit __fadd_rn/__fmul_rn).
#include "common.cuh"

namespace {

// items: [0, F*R)            vtrace advantages for (field f, head r)
//        [F*R, F*R + R)      upgo advantages for head r (field 0)
//        [F*R + R, F*R+R+F)  td-lambda returns for field f
__global__ void return_scan_kernel(const float* __restrict__ reward, const float* __restrict__ value,
                                   const float* __restrict__ rho, const float* __restrict__ gamma_td,
                                   float lambda_td, float* __restrict__ vtrace_adv, float* __restrict__ upgo_adv,
                                   float* __restrict__ td_return, int F, int R, int T, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int item = blockIdx.y;
    if (b >= B) return;
    const size_t TB = (size_t)T * B;
    if (item < F * R) {
        const int f = item / R, r = item % R;
        const float* rw = reward + f * TB;
        const float* v = value + (size_t)f * (T + 1) * B;
        const float* rh = rho + r * TB;
        float* out = vtrace_adv + ((size_t)f * R + r) * TB;
        // vs[T] = V[T];  vs[t] = V[t] + delta[t] + 1*1*c[t]*(vs[t+1]-V[t+1]);  adv[t] = rho*(r + 1*vs[t+1] - V[t])
        float vs_next = v[(size_t)T * B + b];
        float v_next = vs_next;
        for (int t = T - 1; t >= 0; --t) {
            const float vt = v[(size_t)t * B + b], rt = rw[(size_t)t * B + b], ct = rh[(size_t)t * B + b];
            const float delta = __fmul_rn(ct, __fsub_rn(__fadd_rn(rt, __fmul_rn(1.0f, v_next)), vt));
            out[(size_t)t * B + b] = __fmul_rn(ct, __fsub_rn(__fadd_rn(rt, __fmul_rn(1.0f, vs_next)), vt));
            const float vs_t = __fadd_rn(__fadd_rn(vt, delta), __fmul_rn(ct, __fsub_rn(vs_next, v_next)));
            vs_next = vs_t;
            v_next = vt;
        }
    } else if (item < F * R + R) {
        const int r = item - F * R;
        const float* rw = reward;
        const float* v = value;
        const float* rh = rho + r * TB;
        float* out = upgo_adv + r * TB;
        // G[T-1] = r + V[T];  G[t] = r + lam'[t]*G[t+1] + (1-lam'[t])*V[t+1], lam'[t] = [r[t+1]+V[t+2] >= V[t+1]]
        float g_next = 0.f;
        for (int t = T - 1; t >= 0; --t) {
            const float rt = rw[(size_t)t * B + b], v1 = v[(size_t)(t + 1) * B + b], vt = v[(size_t)t * B + b];
            float g;
            if (t == T - 1) {
                g = __fadd_rn(rt, __fmul_rn(1.0f, v1));
            } else {
                const float r1 = rw[(size_t)(t + 1) * B + b], v2 = v[(size_t)(t + 2) * B + b];
                const float lam = (__fadd_rn(r1, v2) >= v1) ? 1.0f : 0.0f;
                // reference: rewards + discounts*result[t+1] + (gammas - discounts)*bootstrap[t], gammas = 1
                g = __fadd_rn(__fadd_rn(rt, __fmul_rn(lam, g_next)), __fmul_rn(__fsub_rn(1.0f, lam), v1));
            }
            out[(size_t)t * B + b] = __fmul_rn(rh[(size_t)t * B + b], __fsub_rn(g, vt));
            g_next = g;
        }
    } else {
        const int f = item - F * R - R;
        const float* rw = reward + f * TB;
        const float* v = value + (size_t)f * (T + 1) * B;
        float* out = td_return + f * TB;
        const float gamma = gamma_td[f];
        const float disc = __fmul_rn(gamma, lambda_td);
        const float rest = __fsub_rn(gamma, disc);
        float g_next = 0.f;
        for (int t = T - 1; t >= 0; --t) {
            const float rt = rw[(size_t)t * B + b], v1 = v[(size_t)(t + 1) * B + b];
            float g;
            if (t == T - 1) g = __fadd_rn(rt, __fmul_rn(gamma, v1));
            else g = __fadd_rn(__fadd_rn(rt, __fmul_rn(disc, g_next)), __fmul_rn(rest, v1));
            out[(size_t)t * B + b] = g;
            g_next = g;
        }
    }
}

}  // namespace

extern "C" int dsb_return_scan(const float* reward, const float* value, const float* rho, const float* gamma_td,
                               float lambda_td, float* vtrace_adv, float* upgo_adv, float* td_return, int F, int R,
                               int T, int B, dsb_stream_t stream) {
    DSB_REQUIRE(reward && value && rho && gamma_td && vtrace_adv && upgo_adv && td_return, "return_scan: null pointer");
    DSB_REQUIRE(F > 0 && R > 0 && T > 0 && B > 0, "return_scan: bad shape");
    const int threads = 128;
    dim3 grid((B + threads - 1) / threads, F * R + R + F);
    return_scan_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(reward, value, rho, gamma_td, lambda_td,
                                                                   vtrace_adv, upgo_adv, td_return, F, R, T, B);
    return dsb::check_launch("return_scan");
}
