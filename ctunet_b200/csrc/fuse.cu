// Fused up-sampling stage: ConvTranspose3d(k2, s2) followed by Conv3d(k^3, "same") as ONE 3x3x3 convolution on the
// low-resolution grid (reference: ctunet/pytorch/models.py:37-38 and :427-430 -- the first two modules of every up block).
//
// Both layers are linear, so for every output phase q in {0,1}^3 of a low-res voxel u
//     y[2u + q][co] = sum_{delta in {-1,0,1}^3} sum_ci Wn[(q, co)][ci][delta] * x[u + delta][ci]
//     Wn[(q, co)][ci][delta] = sum_{kk : floor((q + kk - pad) / 2) = delta} sum_cm W3[co][cm][kk] * WT[ci][cm][(q + kk - pad) mod 2]
// (per dimension).  The transposed convolution's bias rides on an extra all-ones input channel (index cin), zero outside
// the volume like every other channel, which reproduces the zero padding of the (never materialised) 8x larger
// intermediate exactly.  The kernels below build Wn from (WT, bT, W3) and push a gradient dWn back to (dWT, dbT, dW3);
// the convolution itself is the ordinary conv3d path with 8*ceil8(cout) "phase-major" output channels.
// CPU restatement for the tests: oracle/upfuse_oracle.py.
#include "common.cuh"

namespace ctu {

struct FuseDims {
    int cin, cm, cout, cop, k, pad, k3;
};

// per-dimension decomposition of the high-res tap kk seen from output phase q
__device__ __forceinline__ void tap_of(int q, int kk, int pad, int& delta, int& p) {
    const int o = q + kk - pad;          // high-res offset relative to 2u
    delta = (o >= 0) ? (o >> 1) : -((1 - o) >> 1);
    p = o - 2 * delta;
}
__device__ __forceinline__ float wta(const float* __restrict__ wt, const float* __restrict__ bt, const FuseDims& g, int ci,
                                     int cm, int p3) {
    if (ci < g.cin) return wt[((long long)ci * g.cm + cm) * 8 + p3];
    return bt != nullptr ? bt[cm] : 0.f;
}

constexpr int FT = 16;   // tile edge of the small GEMMs below

// Wn[(q*cop + co)][ci][d3]: block = (ci tile, co tile, (q*27 + d3)*k + kd); every thread owns one (co, ci) entry and
// adds its kd slice with one atomic (wn zeroed by the caller).  These are tiny GEMMs bound by load latency, so the work
// is spread over as many blocks as possible instead of looping inside a few.
__global__ void __launch_bounds__(FT* FT) upfuse_compose_kernel(const float* __restrict__ wt, const float* __restrict__ bt,
                                                                const float* __restrict__ w3, float* __restrict__ wn,
                                                                FuseDims g) {
    __shared__ float As[FT][FT + 1];   // W3[co][cm]
    __shared__ float Bs[FT][FT + 1];   // WTa[ci][cm]
    const int tx = threadIdx.x % FT, ty = threadIdx.x / FT;
    const int kd = blockIdx.z % g.k, qz = blockIdx.z / g.k;
    const int q = qz / 27, d3 = qz % 27;
    const int qd = q >> 2, qh = (q >> 1) & 1, qw = q & 1;
    const int dd = d3 / 9 - 1, dh = (d3 / 3) % 3 - 1, dw = d3 % 3 - 1;
    const int ci0 = blockIdx.x * FT, co0 = blockIdx.y * FT;
    float acc = 0.f;
    {
        int e, pd;
        tap_of(qd, kd, g.pad, e, pd);
        if (e != dd) return;
        for (int kh = 0; kh < g.k; ++kh) {
            int ph;
            tap_of(qh, kh, g.pad, e, ph);
            if (e != dh) continue;
            for (int kw = 0; kw < g.k; ++kw) {
                int pw;
                tap_of(qw, kw, g.pad, e, pw);
                if (e != dw) continue;
                const int tap = (kd * g.k + kh) * g.k + kw, p3 = pd * 4 + ph * 2 + pw;
                for (int c0 = 0; c0 < g.cm; c0 += FT) {
                    const int co = co0 + ty, ci = ci0 + ty, cmx = c0 + tx;
                    As[ty][tx] = (co < g.cout && cmx < g.cm) ? w3[((long long)co * g.cm + cmx) * g.k3 + tap] : 0.f;
                    Bs[ty][tx] = (ci <= g.cin && cmx < g.cm) ? wta(wt, bt, g, ci, cmx, p3) : 0.f;
                    __syncthreads();
#pragma unroll
                    for (int j = 0; j < FT; ++j) acc = fmaf(As[ty][j], Bs[tx][j], acc);
                    __syncthreads();
                }
            }
        }
    }
    const int co = co0 + ty, ci = ci0 + tx;
    if (co < g.cout && ci <= g.cin && acc != 0.f) atomicAdd(wn + (((long long)q * g.cop + co) * (g.cin + 1) + ci) * 27 + d3, acc);
}

// dWTa[ci][cm][p3] = sum_{(q, kk) with phase p3} sum_co dWn[(q, co)][ci][delta(q, kk)] * W3[co][cm][kk]
// block = (cm tile, ci tile, p3 + 8 * (qd, kd, qh, kh)): a 16x16 tile of (ci, cm), one thread per entry; shared-memory
// tiles over co; one atomic per entry (dwt zeroed by the caller).  Row ci == cin is the bias: dbT[cm] += its value.
__global__ void __launch_bounds__(FT* FT) upfuse_dwt_kernel(const float* __restrict__ dwn, const float* __restrict__ w3,
                                                            float* __restrict__ dwt, float* __restrict__ dbt, FuseDims g,
                                                            int rows) {
    __shared__ float As[FT][FT + 1];   // dWn[q][co][ci][d3]  as [ci][co]
    __shared__ float Bs[FT][FT + 1];   // W3[co][cm][tap]     as [co][cm]
    const int tx = threadIdx.x % FT, ty = threadIdx.x / FT;
    const int p3 = blockIdx.z & 7;
    int sp = blockIdx.z >> 3;
    const int kh = sp % g.k; sp /= g.k;
    const int qh = sp & 1; sp >>= 1;
    const int kd = sp % g.k, qd = sp / g.k;
    const int pd = p3 >> 2, ph = (p3 >> 1) & 1, pw = p3 & 1;
    const int cm0 = blockIdx.x * FT, ci0 = blockIdx.y * FT;
    float acc = 0.f;
    int dd, dh, p;
    tap_of(qd, kd, g.pad, dd, p);
    if (p != pd) return;
    tap_of(qh, kh, g.pad, dh, p);
    if (p != ph) return;
    for (int qw = 0; qw < 2; ++qw)
        for (int kw = 0; kw < g.k; ++kw) {
            int dw;
            tap_of(qw, kw, g.pad, dw, p);
            if (p != pw) continue;
            const int q = qd * 4 + qh * 2 + qw, d3 = ((dd + 1) * 3 + dh + 1) * 3 + dw + 1;
            const int tap = (kd * g.k + kh) * g.k + kw;
            for (int c0 = 0; c0 < g.cout; c0 += FT) {
                const int ci = ci0 + ty, co_a = c0 + tx;     // As[ty = ci][tx = co]
                As[ty][tx] = (ci < rows && co_a < g.cout)
                                 ? dwn[(((long long)q * g.cop + co_a) * (g.cin + 1) + ci) * 27 + d3] : 0.f;
                const int co_b = c0 + ty, cm = cm0 + tx;     // Bs[ty = co][tx = cm]
                Bs[ty][tx] = (co_b < g.cout && cm < g.cm) ? w3[((long long)co_b * g.cm + cm) * g.k3 + tap] : 0.f;
                __syncthreads();
#pragma unroll
                for (int j = 0; j < FT; ++j) acc = fmaf(As[ty][j], Bs[j][tx], acc);
                __syncthreads();
            }
        }
    const int ci = ci0 + ty, cm = cm0 + tx;
    if (cm < g.cm && acc != 0.f) {
        if (ci < g.cin)
            atomicAdd(dwt + ((long long)ci * g.cm + cm) * 8 + p3, acc);
        else if (ci == g.cin && ci < rows)
            atomicAdd(dbt + cm, acc);
    }
}

// dW3[co][cm][kk] = sum_q sum_{ci <= cin} dWn[(q, co)][ci][delta(q, kk)] * WTa[ci][cm][p(q, kk)]
// block = (cm tile, co tile, tap * 8 + q): a 16x16 tile of (co, cm); shared-memory tiles over ci; one atomic per entry
// (dw3 zeroed by the caller).
__global__ void __launch_bounds__(FT* FT) upfuse_dw3_kernel(const float* __restrict__ dwn, const float* __restrict__ wt,
                                                            const float* __restrict__ bt, float* __restrict__ dw3,
                                                            FuseDims g) {
    __shared__ float As[FT][FT + 1];   // dWn[q][co][ci][d3] as [co][ci]
    __shared__ float Bs[FT][FT + 1];   // WTa[ci][cm][p3]    as [ci][cm]
    const int tx = threadIdx.x % FT, ty = threadIdx.x / FT;
    const int tap = blockIdx.z >> 3, q = blockIdx.z & 7;
    const int kd = tap / (g.k * g.k), kh = (tap / g.k) % g.k, kw = tap % g.k;
    const int cm0 = blockIdx.x * FT, co0 = blockIdx.y * FT;
    float acc = 0.f;
    {
        int dd, dh, dw, pd, ph, pw;
        tap_of(q >> 2, kd, g.pad, dd, pd);
        tap_of((q >> 1) & 1, kh, g.pad, dh, ph);
        tap_of(q & 1, kw, g.pad, dw, pw);
        const int d3 = ((dd + 1) * 3 + dh + 1) * 3 + dw + 1, p3 = pd * 4 + ph * 2 + pw;
        for (int c0 = 0; c0 <= g.cin; c0 += FT) {
            const int co = co0 + ty, ci_a = c0 + tx;          // As[ty = co][tx = ci]
            As[ty][tx] = (co < g.cout && ci_a <= g.cin) ? dwn[(((long long)q * g.cop + co) * (g.cin + 1) + ci_a) * 27 + d3] : 0.f;
            const int ci_b = c0 + ty, cm = cm0 + tx;          // Bs[ty = ci][tx = cm]
            Bs[ty][tx] = (ci_b <= g.cin && cm < g.cm) ? wta(wt, bt, g, ci_b, cm, p3) : 0.f;
            __syncthreads();
#pragma unroll
            for (int j = 0; j < FT; ++j) acc = fmaf(As[ty][j], Bs[j][tx], acc);
            __syncthreads();
        }
    }
    const int co = co0 + ty, cm = cm0 + tx;
    if (co < g.cout && cm < g.cm && acc != 0.f) atomicAdd(dw3 + ((long long)co * g.cm + cm) * g.k3 + tap, acc);
}

// b3n[(q*cop + co)] = b3[co] (pad lanes 0); db3[co] = sum_q dbn[(q*cop + co)]
__global__ void upfuse_bias_kernel(const float* __restrict__ b3, float* __restrict__ b3n, const float* __restrict__ dbn,
                                   float* __restrict__ db3, int cout, int cop) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (b3n != nullptr && i < 8 * cop) {
        const int co = i % cop;
        b3n[i] = co < cout ? b3[co] : 0.f;
    }
    if (db3 != nullptr && i < cout) {
        float s = 0.f;
        for (int q = 0; q < 8; ++q) s += dbn[q * cop + i];
        db3[i] = s;
    }
}

static bool fuse_dims(FuseDims& g, int cin, int cout, int k) {
    if (cin < 1 || cout < 1 || (k != 3 && k != 5)) return false;
    g.cin = cin; g.cm = cin; g.cout = cout; g.cop = (cout + 7) / 8 * 8; g.k = k; g.pad = k / 2; g.k3 = k * k * k;
    return true;
}

}  // namespace ctu

using namespace ctu;

extern "C" {

int ctu_upfuse_cout(int cout) { return 8 * ((cout + 7) / 8 * 8); }

int ctu_upfuse_compose(const float* wt, const float* bt, const float* w3, const float* b3, float* wn, float* b3n, int cin,
                       int cout, int k, ctu_stream stream) {
    FuseDims g;
    CTU_REQUIRE(wt && w3 && wn && fuse_dims(g, cin, cout, k) && ((b3 == nullptr) == (b3n == nullptr)),
                "ctu_upfuse_compose: bad arguments");
    cudaError_t e = cudaMemsetAsync(wn, 0, sizeof(float) * 8 * g.cop * (cin + 1) * 27, (cudaStream_t)stream);
    if (e != cudaSuccess) {
        set_error("ctu_upfuse_compose: memset: %s", cudaGetErrorString(e));
        return (int)e;
    }
    dim3 grid(cdiv(cin + 1, FT), cdiv(cout, FT), 8 * 27 * k);
    upfuse_compose_kernel<<<grid, FT * FT, 0, (cudaStream_t)stream>>>(wt, bt, w3, wn, g);
    int rc = check_launch("ctu_upfuse_compose");
    if (rc == CTU_OK && b3 != nullptr) {
        upfuse_bias_kernel<<<cdiv(8 * g.cop, 128), 128, 0, (cudaStream_t)stream>>>(b3, b3n, nullptr, nullptr, cout, g.cop);
        rc = check_launch("ctu_upfuse_compose(bias)");
    }
    return rc;
}

int ctu_upfuse_decompose(const float* dwn, const float* dbn, const float* wt, const float* bt, const float* w3, float* dwt,
                         float* dbt, float* dw3, float* db3, int cin, int cout, int k, ctu_stream stream) {
    FuseDims g;
    CTU_REQUIRE(dwn && wt && w3 && dwt && dw3 && fuse_dims(g, cin, cout, k) && ((dbn == nullptr) == (db3 == nullptr)) &&
                    ((bt == nullptr) == (dbt == nullptr)),
                "ctu_upfuse_decompose: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(dwt, 0, sizeof(float) * cin * cin * 8, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(dw3, 0, sizeof(float) * cout * cin * g.k3, st);
    if (e == cudaSuccess && dbt != nullptr) e = cudaMemsetAsync(dbt, 0, sizeof(float) * cin, st);
    if (e != cudaSuccess) {
        set_error("ctu_upfuse_decompose: memset: %s", cudaGetErrorString(e));
        return (int)e;
    }
    const int rows = cin + (dbt != nullptr ? 1 : 0);
    upfuse_dwt_kernel<<<dim3(cdiv(g.cm, FT), cdiv(rows, FT), 8 * 4 * k * k), FT * FT, 0, st>>>(dwn, w3, dwt, dbt, g, rows);
    int rc = check_launch("ctu_upfuse_decompose(dwt)");
    if (rc != CTU_OK) return rc;
    upfuse_dw3_kernel<<<dim3(cdiv(g.cm, FT), cdiv(cout, FT), g.k3 * 8), FT * FT, 0, st>>>(dwn, wt, bt, dw3, g);
    rc = check_launch("ctu_upfuse_decompose(dw3)");
    if (rc == CTU_OK && db3 != nullptr) {
        upfuse_bias_kernel<<<cdiv(cout, 128), 128, 0, st>>>(nullptr, nullptr, dbn, db3, cout, g.cop);
        rc = check_launch("ctu_upfuse_decompose(db3)");
    }
    return rc;
}

}  // extern "C"
