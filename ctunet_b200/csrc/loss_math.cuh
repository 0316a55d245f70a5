// Per-voxel arithmetic of the soft-Dice + CrossEntropy loss, shared by loss.cu (fp32 NCDHW predictions) and the fused
// head + loss kernels of head.cu (predictions recomputed from the blocked sources, never written).
//   dice_loss.forward          ctunet/utilities.py:39-50
//   CE + softmax + weighting   ctunet/pytorch/ProblemHandler.py:59-91, 228-298
#pragma once
#include "common.cuh"

namespace ctu {

template <int C>
__device__ __forceinline__ void softmax_c(const float (&x)[C], float (&p)[C], float& lse) {
    float mx = x[0];
#pragma unroll
    for (int c = 1; c < C; ++c) mx = fmaxf(mx, x[c]);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        p[c] = expf(x[c] - mx);
        sum += p[c];
    }
#pragma unroll
    for (int c = 0; c < C; ++c) p[c] /= sum;
    lse = mx + logf(sum);
}

template <int C>
__device__ __forceinline__ int first_argmax(const float (&t)[C]) {
    int bi = 0;
    float best = t[0];
#pragma unroll
    for (int c = 1; c < C; ++c)
        if (t[c] > best || (t[c] != t[c] && best == best)) {
            best = t[c];
            bi = c;
        }
    return bi;
}

// acc += { sum q*t, sum q*q, sum t*t, CE } of one voxel; q = softmax(x) when softmax_for_dice, else x
template <int C>
__device__ __forceinline__ void loss_voxel_fwd(const float (&x)[C], const float (&t)[C], int softmax_for_dice, int want_ce,
                                               float (&acc)[4]) {
    float p[C], lse = 0.f;
    if (softmax_for_dice || want_ce) softmax_c<C>(x, p, lse);
    if (want_ce) {
        const int k = first_argmax<C>(t);
        float xk = x[0];
#pragma unroll
        for (int c = 1; c < C; ++c) xk = (c == k) ? x[c] : xk;
        acc[3] += lse - xk;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float q = softmax_for_dice ? p[c] : x[c];
        acc[0] = fmaf(q, t[c], acc[0]);
        acc[1] = fmaf(q, q, acc[1]);
        acc[2] = fmaf(t[c], t[c], acc[2]);
    }
}

// Per-sample coefficients of the Dice gradient: d dice / d q_c = kt * t_c + kq * q_c with N = sum q*t + eps,
// D = sum q*q + sum t*t + eps:  -(2/B) * (t_c * D - 2 q_c N) / D^2, times the weight g_dice
__device__ __forceinline__ void dice_coefficients(const double* sums4, int nb, float g_dice, float& kt, float& kq) {
    const double eps = 0.0000001;
    const double N = sums4[0] + eps, D = sums4[1] + sums4[2] + eps;
    kt = (float)(-2.0 / nb / D) * g_dice;
    kq = (float)(4.0 / nb * N / (D * D)) * g_dice;
}

// gx = d(total) / dx of one voxel; kce = g_ce / (B * spatial) (0 when the CE term is off)
template <int C>
__device__ __forceinline__ void loss_voxel_bwd(const float (&x)[C], const float (&t)[C], int softmax_for_dice, int want_ce,
                                               float kt, float kq, float kce, float (&gx)[C]) {
    float p[C], lse, gq[C];
    if (softmax_for_dice || want_ce) softmax_c<C>(x, p, lse);
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float q = softmax_for_dice ? p[c] : x[c];
        gq[c] = kt * t[c] + kq * q;
    }
    if (softmax_for_dice) {
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) dot = fmaf(gq[c], p[c], dot);
#pragma unroll
        for (int c = 0; c < C; ++c) gx[c] = p[c] * (gq[c] - dot);
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) gx[c] = gq[c];
    }
    if (want_ce) {
        const int k = first_argmax<C>(t);
#pragma unroll
        for (int c = 0; c < C; ++c) gx[c] += kce * (p[c] - (c == k ? 1.f : 0.f));
    }
}

// ---- two-class fast path of the fused head + loss kernels (head.cu): one exponential per softmax, approximate
// reciprocal / logarithm (abs. error ~1e-7 on probabilities in [0, 1]; the exact forms above stay the ones the
// stand-alone loss kernels use behind the reference's handler interface)
__device__ __forceinline__ void softmax2_fast(float x0, float x1, float& p0, float& p1, float& lse) {
    const float d = x1 - x0;
    const float e = __expf(-fabsf(d));
    const float hi = __fdividef(1.f, 1.f + e), lo = e * hi;
    p0 = d > 0.f ? lo : hi;
    p1 = d > 0.f ? hi : lo;
    lse = fmaxf(x0, x1) + __logf(1.f + e);
}

__device__ __forceinline__ void loss_voxel_fwd2(float x0, float x1, float t0, float t1, int softmax_for_dice, int want_ce,
                                                float (&acc)[4]) {
    float p0 = x0, p1 = x1, lse = 0.f;
    if (softmax_for_dice || want_ce) softmax2_fast(x0, x1, p0, p1, lse);
    if (want_ce) acc[3] += lse - ((t1 > t0) ? x1 : x0);            // first_argmax: ties -> class 0
    const float q0 = softmax_for_dice ? p0 : x0, q1 = softmax_for_dice ? p1 : x1;
    acc[0] = fmaf(q0, t0, fmaf(q1, t1, acc[0]));
    acc[1] = fmaf(q0, q0, fmaf(q1, q1, acc[1]));
    acc[2] = fmaf(t0, t0, fmaf(t1, t1, acc[2]));
}

__device__ __forceinline__ void loss_voxel_bwd2(float x0, float x1, float t0, float t1, int softmax_for_dice, int want_ce,
                                                float kt, float kq, float kce, float& g0, float& g1) {
    float p0 = x0, p1 = x1, lse;
    if (softmax_for_dice || want_ce) softmax2_fast(x0, x1, p0, p1, lse);
    const float q0 = softmax_for_dice ? p0 : x0, q1 = softmax_for_dice ? p1 : x1;
    const float gq0 = fmaf(kt, t0, kq * q0), gq1 = fmaf(kt, t1, kq * q1);
    if (softmax_for_dice) {
        const float dot = fmaf(gq0, p0, gq1 * p1);
        g0 = p0 * (gq0 - dot);
        g1 = p1 * (gq1 - dot);
    } else {
        g0 = gq0;
        g1 = gq1;
    }
    if (want_ce) {
        const bool k1 = t1 > t0;
        g0 = fmaf(kce, p0 - (k1 ? 0.f : 1.f), g0);
        g1 = fmaf(kce, p1 - (k1 ? 1.f : 0.f), g1);
    }
}

}  // namespace ctu
