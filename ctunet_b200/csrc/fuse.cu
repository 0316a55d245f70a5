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

// The GEMM operands are gathered with strides of k^3 (W3), 8 (WT) and 27 (dWn) floats in their native layouts; the tiles
// below read TRANSPOSED copies in a caller-provided workspace instead, so every tile row is one contiguous segment:
//   W3T [tap][co][cm], WTT [p3][ci <= cin][cm] (row cin = the bias), dWnT [q][d3][co][ci <= cin]
struct FuseWs {
    float* w3t;
    float* wtt;
    float* dwnt;
};
static inline long long ws_w3t(const FuseDims& g) { return (long long)g.k3 * g.cout * g.cm; }
static inline long long ws_wtt(const FuseDims& g) { return 8LL * (g.cin + 1) * g.cm; }
static inline long long ws_dwnt(const FuseDims& g) { return 8LL * 27 * g.cout * (g.cin + 1); }

__global__ void upfuse_transpose_kernel(const float* __restrict__ w3, const float* __restrict__ wt, const float* __restrict__ bt,
                                        const float* __restrict__ dwn, FuseWs ws, FuseDims g, long long n3, long long nt,
                                        long long nd) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n3) {                                  // W3T[tap][co][cm] = W3[co][cm][tap]
        const int cm = (int)(i % g.cm);
        const int co = (int)((i / g.cm) % g.cout);
        const int tap = (int)(i / ((long long)g.cm * g.cout));
        ws.w3t[i] = w3[((long long)co * g.cm + cm) * g.k3 + tap];
    } else if (i < n3 + nt) {                      // WTT[p3][ci][cm] = WTa[ci][cm][p3]
        const long long j = i - n3;
        const int cm = (int)(j % g.cm);
        const int ci = (int)((j / g.cm) % (g.cin + 1));
        const int p3 = (int)(j / ((long long)g.cm * (g.cin + 1)));
        ws.wtt[j] = wta(wt, bt, g, ci, cm, p3);
    } else if (i < n3 + nt + nd) {                 // dWnT[q][d3][co][ci] = dWn[(q, co)][ci][d3]
        const long long j = i - n3 - nt;
        const int ci = (int)(j % (g.cin + 1));
        const int co = (int)((j / (g.cin + 1)) % g.cout);
        const int d3 = (int)((j / ((long long)(g.cin + 1) * g.cout)) % 27);
        const int q = (int)(j / ((long long)(g.cin + 1) * g.cout * 27));
        ws.dwnt[j] = dwn[(((long long)q * g.cop + co) * (g.cin + 1) + ci) * 27 + d3];
    }
}

// Wn[(q*cop + co)][ci][d3]: block = (ci tile, co tile, (q*27 + d3)*k + kd); every thread owns 2 x 2 (co, ci) entries and
// adds its kd slice with atomics (wn zeroed by the caller).  These are small GEMMs bound by load latency, so the work is
// spread over as many blocks as possible instead of looping inside a few.
constexpr int GT = 32;   // output tile edge of the GEMMs below: 256 threads, 2 x 2 outputs each, K step FT = 16
// shared tiles are S[k][index]: the FMA loop reads S[j][ty (+16)] (broadcast) and S[j][tx (+16)] (conflict-free)
// operand contiguous along the reduced index: element (r, k) at src[r * ld + k]
__device__ __forceinline__ void tile_kc(float (*S)[GT + 1], const float* __restrict__ src, long long ld, int r0, int nr, int k0,
                                        int nk, int tx, int ty) {
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int r = ty + FT * a;
        S[tx][r] = (r0 + r < nr && k0 + tx < nk) ? src[(long long)(r0 + r) * ld + k0 + tx] : 0.f;
    }
}
// operand contiguous along the kept index: element (k, c) at src[k * ld + c]
__device__ __forceinline__ void tile_nc(float (*S)[GT + 1], const float* __restrict__ src, long long ld, int k0, int nk, int c0,
                                        int nc, int tx, int ty) {
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int c = tx + FT * a;
        S[ty][c] = (k0 + ty < nk && c0 + c < nc) ? src[(long long)(k0 + ty) * ld + c0 + c] : 0.f;
    }
}
__device__ __forceinline__ void tile_fma(const float (*A)[GT + 1], const float (*B)[GT + 1], float (&acc)[2][2], int tx, int ty) {
#pragma unroll
    for (int j = 0; j < FT; ++j) {
        const float a0 = A[j][ty], a1 = A[j][ty + FT], b0 = B[j][tx], b1 = B[j][tx + FT];
        acc[0][0] = fmaf(a0, b0, acc[0][0]);
        acc[0][1] = fmaf(a0, b1, acc[0][1]);
        acc[1][0] = fmaf(a1, b0, acc[1][0]);
        acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
}

__global__ void __launch_bounds__(FT* FT) upfuse_compose_kernel(FuseWs ws, float* __restrict__ wn, FuseDims g) {
    __shared__ float As[FT][GT + 1];   // W3T[tap][co][cm]  as [cm][co]
    __shared__ float Bs[FT][GT + 1];   // WTT[p3][ci][cm]   as [cm][ci]
    const int tx = threadIdx.x % FT, ty = threadIdx.x / FT;
    const int kd = blockIdx.z % g.k, qz = blockIdx.z / g.k;
    const int q = qz / 27, d3 = qz % 27;
    const int qd = q >> 2, qh = (q >> 1) & 1, qw = q & 1;
    const int dd = d3 / 9 - 1, dh = (d3 / 3) % 3 - 1, dw = d3 % 3 - 1;
    const int ci0 = blockIdx.x * GT, co0 = blockIdx.y * GT;
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
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
            const float* a_src = ws.w3t + (long long)tap * g.cout * g.cm;
            const float* b_src = ws.wtt + (long long)p3 * (g.cin + 1) * g.cm;
            for (int c0 = 0; c0 < g.cm; c0 += FT) {
                tile_kc(As, a_src, g.cm, co0, g.cout, c0, g.cm, tx, ty);
                tile_kc(Bs, b_src, g.cm, ci0, g.cin + 1, c0, g.cm, tx, ty);
                __syncthreads();
                tile_fma(As, Bs, acc, tx, ty);
                __syncthreads();
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int co = co0 + ty + FT * a, ci = ci0 + tx + FT * b;
            if (co < g.cout && ci <= g.cin && acc[a][b] != 0.f)
                atomicAdd(wn + (((long long)q * g.cop + co) * (g.cin + 1) + ci) * 27 + d3, acc[a][b]);
        }
}

// dWTa[ci][cm][p3] = sum_{(q, kk) with phase p3} sum_co dWn[(q, co)][ci][delta(q, kk)] * W3[co][cm][kk]
// block = (cm tile, ci tile, p3 + 8 * (qd, kd, qh, kh)): a 32x32 tile of (ci, cm); shared-memory tiles over co; one atomic
// per entry (dwt zeroed by the caller).  Row ci == cin is the bias: dbT[cm] += its value.
__global__ void __launch_bounds__(FT* FT) upfuse_dwt_kernel(FuseWs ws, float* __restrict__ dwt, float* __restrict__ dbt,
                                                            FuseDims g, int rows) {
    __shared__ float As[FT][GT + 1];   // dWnT[q][d3][co][ci]  as [co][ci]
    __shared__ float Bs[FT][GT + 1];   // W3T[tap][co][cm]     as [co][cm]
    const int tx = threadIdx.x % FT, ty = threadIdx.x / FT;
    const int p3 = blockIdx.z & 7;
    int sp = blockIdx.z >> 3;
    const int kh = sp % g.k; sp /= g.k;
    const int qh = sp & 1; sp >>= 1;
    const int kd = sp % g.k, qd = sp / g.k;
    const int pd = p3 >> 2, ph = (p3 >> 1) & 1, pw = p3 & 1;
    const int cm0 = blockIdx.x * GT, ci0 = blockIdx.y * GT;
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
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
            const float* a_src = ws.dwnt + ((long long)q * 27 + d3) * g.cout * (g.cin + 1);
            const float* b_src = ws.w3t + (long long)tap * g.cout * g.cm;
            for (int c0 = 0; c0 < g.cout; c0 += FT) {
                tile_nc(As, a_src, g.cin + 1, c0, g.cout, ci0, rows, tx, ty);
                tile_nc(Bs, b_src, g.cm, c0, g.cout, cm0, g.cm, tx, ty);
                __syncthreads();
                tile_fma(As, Bs, acc, tx, ty);
                __syncthreads();
            }
        }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int ci = ci0 + ty + FT * a, cm = cm0 + tx + FT * b;
            if (cm < g.cm && acc[a][b] != 0.f) {
                if (ci < g.cin)
                    atomicAdd(dwt + ((long long)ci * g.cm + cm) * 8 + p3, acc[a][b]);
                else if (ci == g.cin && ci < rows)
                    atomicAdd(dbt + cm, acc[a][b]);
            }
        }
}

// dW3[co][cm][kk] = sum_q sum_{ci <= cin} dWn[(q, co)][ci][delta(q, kk)] * WTa[ci][cm][p(q, kk)]
// block = (cm tile, co tile, tap * 8 + q): a 32x32 tile of (co, cm); shared-memory tiles over ci; one atomic per entry
// (dw3 zeroed by the caller).
__global__ void __launch_bounds__(FT* FT) upfuse_dw3_kernel(FuseWs ws, float* __restrict__ dw3, FuseDims g) {
    __shared__ float As[FT][GT + 1];   // dWnT[q][d3][co][ci] as [ci][co]
    __shared__ float Bs[FT][GT + 1];   // WTT[p3][ci][cm]     as [ci][cm]
    const int tx = threadIdx.x % FT, ty = threadIdx.x / FT;
    const int tap = blockIdx.z >> 3, q = blockIdx.z & 7;
    const int kd = tap / (g.k * g.k), kh = (tap / g.k) % g.k, kw = tap % g.k;
    const int cm0 = blockIdx.x * GT, co0 = blockIdx.y * GT;
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    int dd, dh, dw, pd, ph, pw;
    tap_of(q >> 2, kd, g.pad, dd, pd);
    tap_of((q >> 1) & 1, kh, g.pad, dh, ph);
    tap_of(q & 1, kw, g.pad, dw, pw);
    const int d3 = ((dd + 1) * 3 + dh + 1) * 3 + dw + 1, p3 = pd * 4 + ph * 2 + pw;
    const float* a_src = ws.dwnt + ((long long)q * 27 + d3) * g.cout * (g.cin + 1);
    const float* b_src = ws.wtt + (long long)p3 * (g.cin + 1) * g.cm;
    for (int c0 = 0; c0 <= g.cin; c0 += FT) {
        tile_kc(As, a_src, g.cin + 1, co0, g.cout, c0, g.cin + 1, tx, ty);
        tile_nc(Bs, b_src, g.cm, c0, g.cin + 1, cm0, g.cm, tx, ty);
        __syncthreads();
        tile_fma(As, Bs, acc, tx, ty);
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int co = co0 + ty + FT * a, cm = cm0 + tx + FT * b;
            if (co < g.cout && cm < g.cm && acc[a][b] != 0.f) atomicAdd(dw3 + ((long long)co * g.cm + cm) * g.k3 + tap, acc[a][b]);
        }
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

long long ctu_upfuse_workspace_floats(int cin, int cout, int k) {
    FuseDims g;
    if (!fuse_dims(g, cin, cout, k)) return -1;
    return ws_w3t(g) + ws_wtt(g) + ws_dwnt(g);
}

static FuseWs make_ws(float* workspace, const FuseDims& g) {
    FuseWs ws;
    ws.w3t = workspace;
    ws.wtt = workspace + ws_w3t(g);
    ws.dwnt = ws.wtt + ws_wtt(g);
    return ws;
}

int ctu_upfuse_compose(const float* wt, const float* bt, const float* w3, const float* b3, float* wn, float* b3n, int cin,
                       int cout, int k, float* workspace, ctu_stream stream) {
    FuseDims g;
    CTU_REQUIRE(wt && w3 && wn && workspace && fuse_dims(g, cin, cout, k) && ((b3 == nullptr) == (b3n == nullptr)),
                "ctu_upfuse_compose: bad arguments");
    const FuseWs ws = make_ws(workspace, g);
    {
        const long long n3 = ws_w3t(g), nt = ws_wtt(g);
        upfuse_transpose_kernel<<<cdiv(n3 + nt, 256), 256, 0, (cudaStream_t)stream>>>(w3, wt, bt, nullptr, ws, g, n3, nt, 0);
    }
    cudaError_t e = cudaMemsetAsync(wn, 0, sizeof(float) * 8 * g.cop * (cin + 1) * 27, (cudaStream_t)stream);
    if (e != cudaSuccess) {
        set_error("ctu_upfuse_compose: memset: %s", cudaGetErrorString(e));
        return (int)e;
    }
    dim3 grid(cdiv(cin + 1, GT), cdiv(cout, GT), 8 * 27 * k);
    upfuse_compose_kernel<<<grid, FT * FT, 0, (cudaStream_t)stream>>>(ws, wn, g);
    int rc = check_launch("ctu_upfuse_compose");
    if (rc == CTU_OK && b3 != nullptr) {
        upfuse_bias_kernel<<<cdiv(8 * g.cop, 128), 128, 0, (cudaStream_t)stream>>>(b3, b3n, nullptr, nullptr, cout, g.cop);
        rc = check_launch("ctu_upfuse_compose(bias)");
    }
    return rc;
}

int ctu_upfuse_decompose(const float* dwn, const float* dbn, const float* wt, const float* bt, const float* w3, float* dwt,
                         float* dbt, float* dw3, float* db3, int cin, int cout, int k, float* workspace, ctu_stream stream) {
    FuseDims g;
    CTU_REQUIRE(dwn && wt && w3 && dwt && dw3 && workspace && fuse_dims(g, cin, cout, k) &&
                    ((dbn == nullptr) == (db3 == nullptr)) && ((bt == nullptr) == (dbt == nullptr)),
                "ctu_upfuse_decompose: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const FuseWs ws = make_ws(workspace, g);
    {
        const long long n3 = ws_w3t(g), nt = ws_wtt(g), nd = ws_dwnt(g);
        upfuse_transpose_kernel<<<cdiv(n3 + nt + nd, 256), 256, 0, st>>>(w3, wt, bt, dwn, ws, g, n3, nt, nd);
    }
    cudaError_t e = cudaMemsetAsync(dwt, 0, sizeof(float) * cin * cin * 8, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(dw3, 0, sizeof(float) * cout * cin * g.k3, st);
    if (e == cudaSuccess && dbt != nullptr) e = cudaMemsetAsync(dbt, 0, sizeof(float) * cin, st);
    if (e != cudaSuccess) {
        set_error("ctu_upfuse_decompose: memset: %s", cudaGetErrorString(e));
        return (int)e;
    }
    const int rows = cin + (dbt != nullptr ? 1 : 0);
    upfuse_dwt_kernel<<<dim3(cdiv(g.cm, GT), cdiv(rows, GT), 8 * 4 * k * k), FT * FT, 0, st>>>(ws, dwt, dbt, g, rows);
    int rc = check_launch("ctu_upfuse_decompose(dwt)");
    if (rc != CTU_OK) return rc;
    upfuse_dw3_kernel<<<dim3(cdiv(g.cm, GT), cdiv(cout, GT), g.k3 * 8), FT * FT, 0, st>>>(ws, dw3, g);
    rc = check_launch("ctu_upfuse_decompose(dw3)");
    if (rc == CTU_OK && db3 != nullptr) {
        upfuse_bias_kernel<<<cdiv(cout, 128), 128, 0, st>>>(nullptr, nullptr, dbn, db3, cout, g.cop);
        rc = check_launch("ctu_upfuse_decompose(db3)");
    }
    return rc;
}

}  // extern "C"
