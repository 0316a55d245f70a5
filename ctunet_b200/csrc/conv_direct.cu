// CUDA-core direct 3-D convolution (k in {1,3,5}, stride 1, "same" zero padding) on the blocked
// layout: forward / data-gradient (same kernel, different packed weights) and weight-gradient.
// This is the fp32 "accumulate-check" path of north_star and the fallback for shapes the tcgen05
// implicit-GEMM kernels (conv_tc.cu) do not cover.
// Reference semantics: nn.Conv3d at ctunet/pytorch/models.py:26,29,38,41,71,76,403,407,430,434.
#include "common.cuh"

namespace ctu {

constexpr int kThreads = 256;
constexpr int RD = 4;  // output voxels per thread along d

struct ConvParams {
    const void* src[CTU_MAX_SRC];
    int src_cb[CTU_MAX_SRC];
    int nsrc;
    int cb_total;
    const float* wp;
    const float* bias;
    void* y;
    int cout, cob_n;
    int n, d, h, w;
    int tw, th, dg;           // thread tile: tw*th*dg <= 256, tile depth td = dg*RD
    int tiles_w, tiles_h, tiles_d;
};

template <typename T, int K>
__global__ void __launch_bounds__(kThreads) conv3d_direct_kernel(ConvParams p) {
    constexpr int PAD = K / 2;
    constexpr int TAPS = K * K * K;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* wsm = reinterpret_cast<float*>(smem_raw);                       // [TAPS][8 ci][8 co]
    T* tile = reinterpret_cast<T*>(smem_raw + sizeof(float) * TAPS * 64);  // [hd][hh][hw][8]

    const int td = p.dg * RD;
    const int hw_ = p.tw + K - 1, hh_ = p.th + K - 1, hd_ = td + K - 1;
    int t = blockIdx.x;
    const int twi = t % p.tiles_w; t /= p.tiles_w;
    const int thi = t % p.tiles_h; t /= p.tiles_h;
    const int tdi = t;
    const int x0 = twi * p.tw, y0 = thi * p.th, z0 = tdi * td;
    const int cob = blockIdx.y, n = blockIdx.z;

    const int tid = threadIdx.x;
    const int lw = tid % p.tw;
    const int r = tid / p.tw;
    const int lh = r % p.th;
    const int dgi = r / p.th;
    const bool active = dgi < p.dg;       // small volumes: the tile is capped, surplus threads only help loading

    float acc[RD][8];
#pragma unroll
    for (int j = 0; j < RD; ++j)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[j][c] = 0.f;

    const long long plane = (long long)p.d * p.h * p.w;
    int cib = 0;
    for (int s = 0; s < p.nsrc; ++s) {
        const T* sp = reinterpret_cast<const T*>(p.src[s]);
        for (int b = 0; b < p.src_cb[s]; ++b, ++cib) {
            __syncthreads();
            // weights of (cob, cib): TAPS*64 floats
            const float* wg = p.wp + ((long long)cob * p.cb_total + cib) * TAPS * 64;
            for (int i = tid; i < TAPS * 16; i += kThreads)
                reinterpret_cast<float4*>(wsm)[i] = __ldg(reinterpret_cast<const float4*>(wg) + i);
            // halo tile of this input block, zero outside the volume
            const T* base = sp + ((long long)n * p.src_cb[s] + b) * plane * 8;
            const int nvox = hd_ * hh_ * hw_;
            for (int i = tid; i < nvox; i += kThreads) {
                int xx = i % hw_;
                int rr = i / hw_;
                int yy = rr % hh_;
                int zz = rr / hh_;
                int gx = x0 + xx - PAD, gy = y0 + yy - PAD, gz = z0 + zz - PAD;
                V8 v;
                if (gx >= 0 && gx < p.w && gy >= 0 && gy < p.h && gz >= 0 && gz < p.d) {
                    v = Vec8<T>::load(base + (((long long)gz * p.h + gy) * p.w + gx) * 8);
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c) v.v[c] = 0.f;
                }
                Vec8<T>::store(tile + (long long)i * 8, v);
            }
            __syncthreads();
            if (!active) continue;
            for (int kd = 0; kd < K; ++kd)
                for (int kh = 0; kh < K; ++kh)
#pragma unroll
                    for (int kw = 0; kw < K; ++kw) {
                        const float* wt = wsm + ((kd * K + kh) * K + kw) * 64;
                        V8 xv[RD];
#pragma unroll
                        for (int j = 0; j < RD; ++j) {
                            int zz = dgi * RD + j + kd, yy = lh + kh, xx = lw + kw;
                            xv[j] = Vec8<T>::load(tile + ((long long)(zz * hh_ + yy) * hw_ + xx) * 8);
                        }
#pragma unroll
                        for (int ci = 0; ci < 8; ++ci) {
                            float4 w0 = *reinterpret_cast<const float4*>(wt + ci * 8);
                            float4 w1 = *reinterpret_cast<const float4*>(wt + ci * 8 + 4);
#pragma unroll
                            for (int j = 0; j < RD; ++j) {
                                float xc = xv[j].v[ci];
                                acc[j][0] = fmaf(xc, w0.x, acc[j][0]);
                                acc[j][1] = fmaf(xc, w0.y, acc[j][1]);
                                acc[j][2] = fmaf(xc, w0.z, acc[j][2]);
                                acc[j][3] = fmaf(xc, w0.w, acc[j][3]);
                                acc[j][4] = fmaf(xc, w1.x, acc[j][4]);
                                acc[j][5] = fmaf(xc, w1.y, acc[j][5]);
                                acc[j][6] = fmaf(xc, w1.z, acc[j][6]);
                                acc[j][7] = fmaf(xc, w1.w, acc[j][7]);
                            }
                        }
                    }
        }
    }
    // epilogue: bias, store
    float bv[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        int ch = cob * 8 + c;
        bv[c] = (p.bias != nullptr && ch < p.cout) ? __ldg(p.bias + ch) : 0.f;
    }
    T* yb = reinterpret_cast<T*>(p.y) + ((long long)n * p.cob_n + cob) * plane * 8;
    const int gx = x0 + lw, gy = y0 + lh;
    if (active && gx < p.w && gy < p.h) {
#pragma unroll
        for (int j = 0; j < RD; ++j) {
            int gz = z0 + dgi * RD + j;
            if (gz < p.d) {
                V8 o;
#pragma unroll
                for (int c = 0; c < 8; ++c) o.v[c] = acc[j][c] + bv[c];
                Vec8<T>::store(yb + (((long long)gz * p.h + gy) * p.w + gx) * 8, o);
            }
        }
    }
}

// ------------------------------------------------------------------------------------ wgrad
struct WgradParams {
    const void* src[CTU_MAX_SRC];
    int src_cb[CTU_MAX_SRC];
    int nsrc;
    int cb_total;
    const void* dy;
    float* dwp;
    float* dbias;
    int cout, cob_n;
    int n, d, h, w;
    int tw, th, td;
    int tiles_w, tiles_h, tiles_d;
};

// One CTA = one spatial tile x one (input block, output block) pair.  Work items are
// (tap, ci-quad, co-quad, voxel slice): 16 accumulators each, reduced into dwp with float atomics.
template <typename T, int K>
__global__ void __launch_bounds__(kThreads) conv3d_wgrad_kernel(WgradParams p) {
    constexpr int PAD = K / 2;
    constexpr int TAPS = K * K * K;
    constexpr int UNITS = TAPS * 4;
    constexpr int SLICES = (kThreads / UNITS) > 0 ? (kThreads / UNITS) : 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int hw_ = p.tw + K - 1, hh_ = p.th + K - 1, hd_ = p.td + K - 1;
    const int nvox_t = p.tw * p.th * p.td;
    T* xt = reinterpret_cast<T*>(smem_raw);                     // halo tile [hd][hh][hw][8]
    T* dyt = xt + (long long)hd_ * hh_ * hw_ * 8;                 // [td][th][tw][8]

    int t = blockIdx.x;
    const int twi = t % p.tiles_w; t /= p.tiles_w;
    const int thi = t % p.tiles_h; t /= p.tiles_h;
    const int tdi = t;
    const int x0 = twi * p.tw, y0 = thi * p.th, z0 = tdi * p.td;
    const int pair = blockIdx.y;
    const int cib = pair % p.cb_total, cob = pair / p.cb_total;
    const int n = blockIdx.z;
    const int tid = threadIdx.x;
    const long long plane = (long long)p.d * p.h * p.w;

    int s = 0, b = cib;
    while (b >= p.src_cb[s]) { b -= p.src_cb[s]; ++s; }
    const T* xb = reinterpret_cast<const T*>(p.src[s]) + ((long long)n * p.src_cb[s] + b) * plane * 8;
    const T* dyb = reinterpret_cast<const T*>(p.dy) + ((long long)n * p.cob_n + cob) * plane * 8;

    const int nhalo = hd_ * hh_ * hw_;
    for (int i = tid; i < nhalo; i += kThreads) {
        int xx = i % hw_;
        int rr = i / hw_;
        int yy = rr % hh_;
        int zz = rr / hh_;
        int gx = x0 + xx - PAD, gy = y0 + yy - PAD, gz = z0 + zz - PAD;
        V8 v;
        if (gx >= 0 && gx < p.w && gy >= 0 && gy < p.h && gz >= 0 && gz < p.d) {
            v = Vec8<T>::load(xb + (((long long)gz * p.h + gy) * p.w + gx) * 8);
        } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) v.v[c] = 0.f;
        }
        Vec8<T>::store(xt + (long long)i * 8, v);
    }
    for (int i = tid; i < nvox_t; i += kThreads) {
        int xx = i % p.tw;
        int rr = i / p.tw;
        int yy = rr % p.th;
        int zz = rr / p.th;
        int gx = x0 + xx, gy = y0 + yy, gz = z0 + zz;
        V8 v;
        if (gx < p.w && gy < p.h && gz < p.d) {
            v = Vec8<T>::load(dyb + (((long long)gz * p.h + gy) * p.w + gx) * 8);
        } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) v.v[c] = 0.f;
        }
        Vec8<T>::store(dyt + (long long)i * 8, v);
    }
    __syncthreads();

    // bias gradient: sum of dy over voxels (once per output block: the cib == 0 CTAs do it)
    if (p.dbias != nullptr && cib == 0) {
        float bs[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) bs[c] = 0.f;
        for (int i = tid; i < nvox_t; i += kThreads) {
            V8 v = Vec8<T>::load(dyt + (long long)i * 8);
#pragma unroll
            for (int c = 0; c < 8; ++c) bs[c] += v.v[c];
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float sum = warp_sum(bs[c]);
            int ch = cob * 8 + c;
            if ((tid & 31) == 0 && ch < p.cout && sum != 0.f) atomicAdd(p.dbias + ch, sum);
        }
    }

    float* out = p.dwp + ((long long)cob * p.cb_total + cib) * TAPS * 64;
    for (int item = tid; item < UNITS * SLICES; item += kThreads) {
        const int unit = item % UNITS, slice = item / UNITS;
        const int tap = unit >> 2, ciq = (unit >> 1) & 1, coq = unit & 1;
        const int kd = tap / (K * K), kh = (tap / K) % K, kw = tap % K;
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
        for (int v = slice; v < nvox_t; v += SLICES) {
            int xx = v % p.tw;
            int rr = v / p.tw;
            int yy = rr % p.th;
            int zz = rr / p.th;
            const T* xp = xt + ((long long)((zz + kd) * hh_ + (yy + kh)) * hw_ + (xx + kw)) * 8 + ciq * 4;
            const T* dp = dyt + (long long)v * 8 + coq * 4;
            float xa[4], da[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                xa[a] = (float)xp[a];
                da[a] = (float)dp[a];
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(xa[a], da[c], acc[a][c]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (acc[a][c] != 0.f) atomicAdd(out + tap * 64 + (ciq * 4 + a) * 8 + coq * 4 + c, acc[a][c]);
    }
}

static void pick_tile(int w, int h, int d, int& tw, int& th, int& dg) {
    tw = 32;
    while (tw > 1 && tw / 2 >= w) tw /= 2;      // smallest power of two >= min(w, 32)
    int rest = kThreads / tw;
    th = 8;
    if (rest < th) th = rest;
    while (th > 1 && th / 2 >= h) th /= 2;
    while (th * 2 <= rest && th < h && th < 16 && tw < 32) th *= 2;
    dg = rest / th;
    const int need = (d + RD - 1) / RD;          // never tile deeper than the volume
    if (dg > need) dg = need;
}

template <typename T>
static int launch_fprop(const ConvParams& p, int k, cudaStream_t stream) {
    const int td = p.dg * RD;
    const size_t tile_b = (size_t)(td + k - 1) * (p.th + k - 1) * (p.tw + k - 1) * 8 * sizeof(T);
    const size_t smem = (size_t)k * k * k * 64 * sizeof(float) + tile_b;
    dim3 grid(p.tiles_w * p.tiles_h * p.tiles_d, p.cob_n, p.n);
    auto go = [&](auto kern) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("conv3d_fprop: smem %zu: %s", smem, cudaGetErrorString(e));
            return (int)e;
        }
        kern<<<grid, kThreads, smem, stream>>>(p);
        return check_launch("ctu_conv3d_fprop(direct)");
    };
    if (k == 1) return go(conv3d_direct_kernel<T, 1>);
    if (k == 3) return go(conv3d_direct_kernel<T, 3>);
    if (k == 5) return go(conv3d_direct_kernel<T, 5>);
    set_error("conv3d: kernel size %d unsupported (1, 3, 5)", k);
    return CTU_ERR_UNSUPPORTED;
}

template <typename T>
static int launch_wgrad(const WgradParams& p, int k, cudaStream_t stream) {
    const size_t smem = ((size_t)(p.td + k - 1) * (p.th + k - 1) * (p.tw + k - 1) + (size_t)p.td * p.th * p.tw) * 8 * sizeof(T);
    dim3 grid(p.tiles_w * p.tiles_h * p.tiles_d, p.cb_total * p.cob_n, p.n);
    auto go = [&](auto kern) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("conv3d_wgrad: smem %zu: %s", smem, cudaGetErrorString(e));
            return (int)e;
        }
        kern<<<grid, kThreads, smem, stream>>>(p);
        return check_launch("ctu_conv3d_wgrad(direct)");
    };
    if (k == 1) return go(conv3d_wgrad_kernel<T, 1>);
    if (k == 3) return go(conv3d_wgrad_kernel<T, 3>);
    if (k == 5) return go(conv3d_wgrad_kernel<T, 5>);
    set_error("conv3d wgrad: kernel size %d unsupported (1, 3, 5)", k);
    return CTU_ERR_UNSUPPORTED;
}

// implemented in conv_tc.cu
int conv3d_fprop_tc(const void* const* h_srcs, const int* h_src_channels, int nsrc, const float* wp, const float* bias,
                    void* y, double* stats, int stat_cout, int cout, int k, int n, int d, int h, int w, cudaStream_t stream,
                    const void* bn_y = nullptr, const float* bn_ss = nullptr, int bn_pm = 0, int prezeroed = 0);
int conv3d_wgrad_tc(const void* const* h_srcs, const int* h_src_channels, int nsrc, const void* dy, float* dwp,
                    float* dbias, int phase_cout, int cout, int k, int n, int d, int h, int w, cudaStream_t stream);
// implemented in conv_wide.cu
int conv3d_wgrad_small(const void* x, int cin, const void* dy, float* dwp, float* dbias, int cout, int k, int n, int d, int h,
                       int w, cudaStream_t stream);
int conv3d_fprop_wide(const void* x, int cin, const void* wimg, const float* bias, void* y, int cout, int k, int n, int d,
                      int h, int w, cudaStream_t stream);

}  // namespace ctu

using namespace ctu;

extern "C" {

int ctu_conv3d_fprop(int dtype, const void* const* h_srcs, const int* h_src_channels, int nsrc, const void* wp,
                     const float* bias, void* y, double* bn_sums, int stat_cout, int cout, int k, int n, int d, int h,
                     int w, int use_tensor_path, ctu_stream stream) {
    SrcMap m;
    int rc = make_srcmap(m, nsrc, h_src_channels);
    if (rc != CTU_OK) return rc;
    const int prezeroed = use_tensor_path & CTU_ACCUM_PREZEROED;
    use_tensor_path &= ~CTU_ACCUM_PREZEROED;
    CTU_REQUIRE(h_srcs && wp && y && cout > 0 && n > 0 && d > 0 && h > 0 && w > 0, "ctu_conv3d_fprop: bad arguments");
    if (stat_cout <= 0) stat_cout = cout;
    const int cob_nat = (stat_cout + 7) / 8;
    CTU_REQUIRE(stat_cout == cout || (cout % 8 == 0 && (cout / 8) % cob_nat == 0),
                "ctu_conv3d_fprop: stat_cout=%d does not divide the %d output blocks", stat_cout, cout / 8);
    for (int i = 0; i < nsrc; ++i) CTU_REQUIRE(h_srcs[i] != nullptr, "ctu_conv3d_fprop: null source %d", i);
    if (use_tensor_path) {
        if (dtype != CTU_BF16) {
            set_error("ctu_conv3d_fprop: the tensor path is bf16 only");
            return CTU_ERR_UNSUPPORTED;
        }
        if (use_tensor_path == 2) {      // weight-streaming kernel for wide, low-resolution layers (conv_wide.cu)
            CTU_REQUIRE(nsrc == 1, "ctu_conv3d_fprop: the wide tensor path takes one source");
            rc = conv3d_fprop_wide(h_srcs[0], h_src_channels[0], wp, bias, y, cout, k, n, d, h, w, (cudaStream_t)stream);
            if (rc == CTU_OK && bn_sums != nullptr)
                rc = ctu_bn_stats(dtype, y, stat_cout, (((cout + 7) / 8) / cob_nat) | prezeroed, n, (long long)d * h * w, bn_sums, stream);
            return rc;
        }
        return conv3d_fprop_tc(h_srcs, h_src_channels, nsrc, (const float*)wp, bias, y, bn_sums, stat_cout, cout, k, n, d,
                               h, w, (cudaStream_t)stream, nullptr, nullptr, 0, prezeroed);
    }
    ConvParams p;
    for (int i = 0; i < CTU_MAX_SRC; ++i) {
        p.src[i] = i < nsrc ? h_srcs[i] : nullptr;
        p.src_cb[i] = i < nsrc ? (m.ch[i] + 7) / 8 : 0;
    }
    p.nsrc = nsrc; p.cb_total = m.cb_total; p.wp = (const float*)wp; p.bias = bias; p.y = y;
    p.cout = cout; p.cob_n = (cout + 7) / 8; p.n = n; p.d = d; p.h = h; p.w = w;
    pick_tile(w, h, d, p.tw, p.th, p.dg);
    p.tiles_w = cdiv(w, p.tw); p.tiles_h = cdiv(h, p.th); p.tiles_d = cdiv(d, p.dg * RD);
    CTU_DISPATCH_DTYPE(dtype, rc = launch_fprop<T>(p, k, (cudaStream_t)stream));
    if (rc == CTU_OK && bn_sums != nullptr)
        rc = ctu_bn_stats(dtype, y, stat_cout, (((cout + 7) / 8) / cob_nat) | prezeroed, n, (long long)d * h * w, bn_sums, stream);
    return rc;
}

int ctu_conv3d_dgrad_bnred(const void* const* h_srcs, const int* h_src_channels, int nsrc, const void* wimg, void* dx,
                           int cout, int k, int n, int d, int h, int w, const void* bn_y, const float* bn_ss,
                           double* bn_sums2, int bn_y_phase_major, ctu_stream stream) {
    SrcMap m;
    int rc = make_srcmap(m, nsrc, h_src_channels);
    if (rc != CTU_OK) return rc;
    CTU_REQUIRE(h_srcs && wimg && dx && bn_y && bn_ss && bn_sums2 && cout > 0 && n > 0 && d > 0 && h > 0 && w > 0,
                "ctu_conv3d_dgrad_bnred: bad arguments");
    CTU_REQUIRE(!bn_y_phase_major || (d % 2 == 0 && h % 2 == 0 && w % 2 == 0), "ctu_conv3d_dgrad_bnred: phase-major y needs even dims");
    for (int i = 0; i < nsrc; ++i) CTU_REQUIRE(h_srcs[i] != nullptr, "ctu_conv3d_dgrad_bnred: null source %d", i);
    return conv3d_fprop_tc(h_srcs, h_src_channels, nsrc, (const float*)wimg, nullptr, dx, bn_sums2, cout, cout, k, n, d, h, w,
                           (cudaStream_t)stream, bn_y, bn_ss, bn_y_phase_major);
}

int ctu_conv3d_wgrad(int dtype, const void* const* h_srcs, const int* h_src_channels, int nsrc, const void* dy,
                     float* dwp, float* dbias, int phase_cout, int cout, int k, int n, int d, int h, int w,
                     int use_tensor_path, ctu_stream stream) {
    SrcMap m;
    int rc = make_srcmap(m, nsrc, h_src_channels);
    if (rc != CTU_OK) return rc;
    CTU_REQUIRE(h_srcs && dy && dwp && cout > 0 && n > 0 && d > 0 && h > 0 && w > 0, "ctu_conv3d_wgrad: bad arguments");
    CTU_REQUIRE(k == 1 || k == 3 || k == 5, "ctu_conv3d_wgrad: kernel size %d unsupported", k);
    const long long nfl = (long long)((cout + 7) / 8) * m.cb_total * k * k * k * 64;
    const bool prezeroed = (use_tensor_path & CTU_ACCUM_PREZEROED) != 0;      // dwp (not dbias) was zeroed by the caller
    use_tensor_path &= ~CTU_ACCUM_PREZEROED;
    cudaError_t e = prezeroed ? cudaSuccess : cudaMemsetAsync(dwp, 0, nfl * sizeof(float), (cudaStream_t)stream);
    if (e == cudaSuccess && dbias) e = cudaMemsetAsync(dbias, 0, cout * sizeof(float), (cudaStream_t)stream);
    if (e != cudaSuccess) {
        set_error("ctu_conv3d_wgrad: memset: %s", cudaGetErrorString(e));
        return (int)e;
    }
    if (use_tensor_path) {
        if (dtype != CTU_BF16) {
            set_error("ctu_conv3d_wgrad: the tensor path is bf16 only");
            return CTU_ERR_UNSUPPORTED;
        }
        if (use_tensor_path == 2) {      // tap-stationary kernel for 8 x 8 plane tiles (conv_wide.cu)
            CTU_REQUIRE(nsrc == 1, "ctu_conv3d_wgrad: the small-grid tensor path takes one source");
            return conv3d_wgrad_small(h_srcs[0], h_src_channels[0], dy, dwp, dbias, cout, k, n, d, h, w, (cudaStream_t)stream);
        }
        return conv3d_wgrad_tc(h_srcs, h_src_channels, nsrc, dy, dwp, dbias, phase_cout, cout, k, n, d, h, w,
                               (cudaStream_t)stream);
    }
    WgradParams p;
    for (int i = 0; i < CTU_MAX_SRC; ++i) {
        p.src[i] = i < nsrc ? h_srcs[i] : nullptr;
        p.src_cb[i] = i < nsrc ? (m.ch[i] + 7) / 8 : 0;
    }
    p.nsrc = nsrc; p.cb_total = m.cb_total; p.dy = dy; p.dwp = dwp; p.dbias = dbias;
    p.cout = cout; p.cob_n = (cout + 7) / 8; p.n = n; p.d = d; p.h = h; p.w = w;
    p.tw = w < 16 ? w : 16; p.th = h < 8 ? h : 8; p.td = d < 8 ? d : 8;
    p.tiles_w = cdiv(w, p.tw); p.tiles_h = cdiv(h, p.th); p.tiles_d = cdiv(d, p.td);
    CTU_DISPATCH_DTYPE(dtype, return launch_wgrad<T>(p, k, (cudaStream_t)stream));
    return CTU_OK;
}

}  // extern "C"
