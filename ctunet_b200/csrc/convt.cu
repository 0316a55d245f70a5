// ConvTranspose3d, kernel 2, stride 2, with bias (ctunet/pytorch/models.py:37, :427) on the
// blocked layout.  Non-overlapping: output voxel (2z+a, 2y+b, 2x+c) depends on input voxel
// (z, y, x) only, so the op is 8 independent 1x1x1 convolutions plus a pixel shuffle.
//   y[co, 2v+abc] = bias[co] + sum_ci x[ci, v] * W[ci, co, abc]
#include "common.cuh"

namespace ctu {

constexpr int kCtThreads = 256;

struct ConvTParams {
    const void* src[CTU_MAX_SRC];
    int src_cb[CTU_MAX_SRC];
    int nsrc;
    int cb_total;
    const float* wp;     // fprop: [cob][cib][abc][ci][co]; dgrad: [cbs][cob][abc][co][ci]
    const float* bias;
    void* y;             // fprop: output; dgrad: dx
    const void* dy;
    float* dwp;
    float* dbias;
    int cout, cob_n;
    int n, d, h, w;      // INPUT dims
    int vox_per_cta;
};

template <typename T>
__global__ void __launch_bounds__(kCtThreads) convt2_fprop_kernel(ConvTParams p) {
    extern __shared__ __align__(16) float wsm[];   // [cb_total][8 abc][8 ci][8 co]
    const int cob = blockIdx.y, n = blockIdx.z, tid = threadIdx.x;
    const float* wg = p.wp + (long long)cob * p.cb_total * 512;
    for (int i = tid; i < p.cb_total * 128; i += kCtThreads)
        reinterpret_cast<float4*>(wsm)[i] = __ldg(reinterpret_cast<const float4*>(wg) + i);
    __syncthreads();
    const long long plane = (long long)p.d * p.h * p.w;
    const long long v = (long long)blockIdx.x * kCtThreads + tid;
    if (v >= plane) return;
    const int x = (int)(v % p.w);
    const int y = (int)((v / p.w) % p.h);
    const int z = (int)(v / ((long long)p.w * p.h));

    float acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[a][c] = 0.f;
    int cib = 0;
    for (int s = 0; s < p.nsrc; ++s) {
        const T* sp = reinterpret_cast<const T*>(p.src[s]);
        for (int b = 0; b < p.src_cb[s]; ++b, ++cib) {
            V8 xv = Vec8<T>::load(sp + (((long long)n * p.src_cb[s] + b) * plane + v) * 8);
            const float* wb = wsm + cib * 512;
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int ci = 0; ci < 8; ++ci) {
                    float4 w0 = *reinterpret_cast<const float4*>(wb + a * 64 + ci * 8);
                    float4 w1 = *reinterpret_cast<const float4*>(wb + a * 64 + ci * 8 + 4);
                    float xc = xv.v[ci];
                    acc[a][0] = fmaf(xc, w0.x, acc[a][0]);
                    acc[a][1] = fmaf(xc, w0.y, acc[a][1]);
                    acc[a][2] = fmaf(xc, w0.z, acc[a][2]);
                    acc[a][3] = fmaf(xc, w0.w, acc[a][3]);
                    acc[a][4] = fmaf(xc, w1.x, acc[a][4]);
                    acc[a][5] = fmaf(xc, w1.y, acc[a][5]);
                    acc[a][6] = fmaf(xc, w1.z, acc[a][6]);
                    acc[a][7] = fmaf(xc, w1.w, acc[a][7]);
                }
        }
    }
    float bv[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        int ch = cob * 8 + c;
        bv[c] = (p.bias != nullptr && ch < p.cout) ? __ldg(p.bias + ch) : 0.f;
    }
    const int oh = 2 * p.h, ow = 2 * p.w;
    const long long oplane = plane * 8;
    T* yb = reinterpret_cast<T*>(p.y) + ((long long)n * p.cob_n + cob) * oplane * 8;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int az = a >> 2, ay = (a >> 1) & 1, ax = a & 1;
        V8 o;
#pragma unroll
        for (int c = 0; c < 8; ++c) o.v[c] = acc[a][c] + bv[c];
        Vec8<T>::store(yb + (((long long)(2 * z + az) * oh + (2 * y + ay)) * ow + (2 * x + ax)) * 8, o);
    }
}

// dx[ci, v] = sum_{co, abc} dy[co, 2v+abc] * W[ci, co, abc]  (one source, blocks cbs = gridDim.y)
template <typename T>
__global__ void __launch_bounds__(kCtThreads) convt2_dgrad_kernel(ConvTParams p) {
    extern __shared__ __align__(16) float wsm[];   // [cob_n][8 abc][8 co][8 ci]
    const int cbs_i = blockIdx.y, n = blockIdx.z, tid = threadIdx.x;
    const float* wg = p.wp + (long long)cbs_i * p.cob_n * 512;
    for (int i = tid; i < p.cob_n * 128; i += kCtThreads)
        reinterpret_cast<float4*>(wsm)[i] = __ldg(reinterpret_cast<const float4*>(wg) + i);
    __syncthreads();
    const long long plane = (long long)p.d * p.h * p.w;
    const long long v = (long long)blockIdx.x * kCtThreads + tid;
    if (v >= plane) return;
    const int x = (int)(v % p.w);
    const int y = (int)((v / p.w) % p.h);
    const int z = (int)(v / ((long long)p.w * p.h));
    const int oh = 2 * p.h, ow = 2 * p.w;
    const long long oplane = plane * 8;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    const T* dyp = reinterpret_cast<const T*>(p.dy);
    for (int cob = 0; cob < p.cob_n; ++cob) {
        const T* db = dyp + ((long long)n * p.cob_n + cob) * oplane * 8;
        const float* wb = wsm + cob * 512;
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int az = a >> 2, ay = (a >> 1) & 1, ax = a & 1;
            V8 g = Vec8<T>::load(db + (((long long)(2 * z + az) * oh + (2 * y + ay)) * ow + (2 * x + ax)) * 8);
#pragma unroll
            for (int co = 0; co < 8; ++co) {
                float4 w0 = *reinterpret_cast<const float4*>(wb + a * 64 + co * 8);
                float4 w1 = *reinterpret_cast<const float4*>(wb + a * 64 + co * 8 + 4);
                float gc = g.v[co];
                acc[0] = fmaf(gc, w0.x, acc[0]);
                acc[1] = fmaf(gc, w0.y, acc[1]);
                acc[2] = fmaf(gc, w0.z, acc[2]);
                acc[3] = fmaf(gc, w0.w, acc[3]);
                acc[4] = fmaf(gc, w1.x, acc[4]);
                acc[5] = fmaf(gc, w1.y, acc[5]);
                acc[6] = fmaf(gc, w1.z, acc[6]);
                acc[7] = fmaf(gc, w1.w, acc[7]);
            }
        }
    }
    V8 o;
#pragma unroll
    for (int c = 0; c < 8; ++c) o.v[c] = acc[c];
    Vec8<T>::store(reinterpret_cast<T*>(p.y) + (((long long)n * gridDim.y + cbs_i) * plane + v) * 8, o);
}

// dW[ci, co, abc] = sum_v x[ci, v] * dy[co, 2v+abc];  dbias[co] = sum dy[co, .]
// One CTA = a chunk of input voxels x one (cib, cob) pair; warp w takes voxels w, w+8, ...;
// lane = (abc, ci-quad, co-quad) owns a 4x4 block of the pair's 8x8x8 outputs.
template <typename T>
__global__ void __launch_bounds__(kCtThreads) convt2_wgrad_kernel(ConvTParams p) {
    __shared__ float red[8][32][17];
    const int pair = blockIdx.y, n = blockIdx.z, tid = threadIdx.x;
    const int cib = pair % p.cb_total, cob = pair / p.cb_total;
    const int warp = tid >> 5, lane = tid & 31;
    const int a = lane >> 2, ciq = (lane >> 1) & 1, coq = lane & 1;
    const int az = a >> 2, ay = (a >> 1) & 1, ax = a & 1;
    const long long plane = (long long)p.d * p.h * p.w;
    const int oh = 2 * p.h, ow = 2 * p.w;
    const long long oplane = plane * 8;
    int s = 0, b = cib;
    while (b >= p.src_cb[s]) { b -= p.src_cb[s]; ++s; }
    const T* xb = reinterpret_cast<const T*>(p.src[s]) + ((long long)n * p.src_cb[s] + b) * plane * 8;
    const T* db = reinterpret_cast<const T*>(p.dy) + ((long long)n * p.cob_n + cob) * oplane * 8;

    float acc[4][4], bs[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        bs[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    }
    const long long v0 = (long long)blockIdx.x * p.vox_per_cta;
    long long v1 = v0 + p.vox_per_cta;
    if (v1 > plane) v1 = plane;
    for (long long v = v0 + warp; v < v1; v += 8) {
        const int x = (int)(v % p.w);
        const int y = (int)((v / p.w) % p.h);
        const int z = (int)(v / ((long long)p.w * p.h));
        const T* xp = xb + v * 8 + ciq * 4;
        const T* dp = db + (((long long)(2 * z + az) * oh + (2 * y + ay)) * ow + (2 * x + ax)) * 8 + coq * 4;
        float xa[4], da[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            xa[i] = (float)xp[i];
            da[i] = (float)dp[i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            bs[i] += da[i];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], da[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) red[warp][lane][i * 4 + j] = acc[i][j];
    red[warp][lane][16] = 0.f;
    __syncthreads();
    // 512 outputs of the pair: thread t sums over the 8 warps for outputs t and t+256
    float* out = p.dwp + ((long long)cob * p.cb_total + cib) * 512;
    for (int o = tid; o < 512; o += kCtThreads) {
        const int ln = o >> 4, e = o & 15;           // lane that owns it, element within its 4x4
        float sum = 0.f;
#pragma unroll
        for (int wq = 0; wq < 8; ++wq) sum += red[wq][ln][e];
        const int la = ln >> 2, lciq = (ln >> 1) & 1, lcoq = ln & 1;
        const int ci = lciq * 4 + (e >> 2), co = lcoq * 4 + (e & 3);
        if (sum != 0.f) atomicAdd(out + la * 64 + ci * 8 + co, sum);
    }
    if (p.dbias != nullptr && cib == 0) {
        __syncthreads();
        // lanes with ciq == 0 hold the dy sums of (abc, coq); reduce over abc and warps
#pragma unroll
        for (int i = 0; i < 4; ++i) red[warp][lane][i] = (ciq == 0) ? bs[i] : 0.f;
        __syncthreads();
        if (tid < 8) {
            const int co = tid, q = co >> 2, i = co & 3;
            float sum = 0.f;
            for (int wq = 0; wq < 8; ++wq)
                for (int ln = 0; ln < 32; ++ln)
                    if ((ln & 1) == q && ((ln >> 1) & 1) == 0) sum += red[wq][ln][i];
            const int ch = cob * 8 + co;
            if (ch < p.cout && sum != 0.f) atomicAdd(p.dbias + ch, sum);
        }
    }
}

static void fill_sources(ConvTParams& p, const SrcMap& m, const void* const* h_srcs, int nsrc) {
    for (int i = 0; i < CTU_MAX_SRC; ++i) {
        p.src[i] = (h_srcs && i < nsrc) ? h_srcs[i] : nullptr;
        p.src_cb[i] = i < nsrc ? (m.ch[i] + 7) / 8 : 0;
    }
    p.nsrc = nsrc;
    p.cb_total = m.cb_total;
}

template <typename KernT>
static int set_smem(KernT kern, size_t smem, const char* what) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("%s: smem %zu: %s", what, smem, cudaGetErrorString(e));
            return (int)e;
        }
    }
    return CTU_OK;
}

}  // namespace ctu

using namespace ctu;

extern "C" {

int ctu_convt2_fprop(int dtype, const void* const* h_srcs, const int* h_src_channels, int nsrc, const float* wp,
                     const float* bias, void* y, int cout, int n, int d, int h, int w, ctu_stream stream) {
    SrcMap m;
    int rc = make_srcmap(m, nsrc, h_src_channels);
    if (rc != CTU_OK) return rc;
    CTU_REQUIRE(h_srcs && wp && y && cout > 0 && n > 0 && d > 0 && h > 0 && w > 0, "ctu_convt2_fprop: bad arguments");
    ConvTParams p = {};
    fill_sources(p, m, h_srcs, nsrc);
    p.wp = wp; p.bias = bias; p.y = y; p.cout = cout; p.cob_n = (cout + 7) / 8;
    p.n = n; p.d = d; p.h = h; p.w = w;
    const long long plane = (long long)d * h * w;
    const size_t smem = (size_t)m.cb_total * 512 * sizeof(float);
    dim3 grid(cdiv(plane, kCtThreads), p.cob_n, n);
    CTU_DISPATCH_DTYPE(dtype, {
        rc = set_smem(convt2_fprop_kernel<T>, smem, "ctu_convt2_fprop");
        if (rc != CTU_OK) return rc;
        convt2_fprop_kernel<T><<<grid, kCtThreads, smem, (cudaStream_t)stream>>>(p);
    });
    return check_launch("ctu_convt2_fprop");
}

int ctu_convt2_dgrad(int dtype, const void* dy, const float* wpd, void* dx, int cout, int src_channels, int n, int d,
                     int h, int w, ctu_stream stream) {
    CTU_REQUIRE(dy && wpd && dx && cout > 0 && src_channels > 0 && n > 0 && d > 0 && h > 0 && w > 0,
                "ctu_convt2_dgrad: bad arguments");
    ConvTParams p = {};
    p.wp = wpd; p.y = dx; p.dy = dy; p.cout = cout; p.cob_n = (cout + 7) / 8;
    p.n = n; p.d = d; p.h = h; p.w = w;
    const long long plane = (long long)d * h * w;
    const size_t smem = (size_t)p.cob_n * 512 * sizeof(float);
    dim3 grid(cdiv(plane, kCtThreads), (src_channels + 7) / 8, n);
    int rc;
    CTU_DISPATCH_DTYPE(dtype, {
        rc = set_smem(convt2_dgrad_kernel<T>, smem, "ctu_convt2_dgrad");
        if (rc != CTU_OK) return rc;
        convt2_dgrad_kernel<T><<<grid, kCtThreads, smem, (cudaStream_t)stream>>>(p);
    });
    return check_launch("ctu_convt2_dgrad");
}

int ctu_convt2_wgrad(int dtype, const void* const* h_srcs, const int* h_src_channels, int nsrc, const void* dy,
                     float* dwp, float* dbias, int cout, int n, int d, int h, int w, ctu_stream stream) {
    SrcMap m;
    int rc = make_srcmap(m, nsrc, h_src_channels);
    if (rc != CTU_OK) return rc;
    CTU_REQUIRE(h_srcs && dy && dwp && cout > 0 && n > 0 && d > 0 && h > 0 && w > 0, "ctu_convt2_wgrad: bad arguments");
    ConvTParams p = {};
    fill_sources(p, m, h_srcs, nsrc);
    p.dy = dy; p.dwp = dwp; p.dbias = dbias; p.cout = cout; p.cob_n = (cout + 7) / 8;
    p.n = n; p.d = d; p.h = h; p.w = w;
    const long long plane = (long long)d * h * w;
    p.vox_per_cta = 16384;
    const long long nfl = (long long)p.cob_n * m.cb_total * 512;
    cudaError_t e = cudaMemsetAsync(dwp, 0, nfl * sizeof(float), (cudaStream_t)stream);
    if (e == cudaSuccess && dbias) e = cudaMemsetAsync(dbias, 0, cout * sizeof(float), (cudaStream_t)stream);
    if (e != cudaSuccess) {
        set_error("ctu_convt2_wgrad: memset: %s", cudaGetErrorString(e));
        return (int)e;
    }
    dim3 grid(cdiv(plane, p.vox_per_cta), m.cb_total * p.cob_n, n);
    CTU_DISPATCH_DTYPE(dtype, (convt2_wgrad_kernel<T><<<grid, kCtThreads, 0, (cudaStream_t)stream>>>(p)));
    return check_launch("ctu_convt2_wgrad");
}

}  // extern "C"
