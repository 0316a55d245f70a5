// Integer / byte glue kernels: virtual-craniectomy masking and CT preprocessing.
//   random_blank_patch / shape_3d   ctunet/pytorch/transforms.py:241-300, ctunet/utilities.py:127-178
//   HU windowing, thresholding, resampling: no reference implementation;
//   oracle/unet_oracle.py defines the semantics these kernels are bit-compared against.
#include "common.cuh"

namespace ctu {

constexpr int kVoxPerBlock = 4096;

__global__ void count_blocks_kernel(const unsigned char* __restrict__ img, long long nvox,
                                    long long* __restrict__ block_counts, long long* __restrict__ total) {
    const long long base = (long long)blockIdx.x * kVoxPerBlock;
    int cnt = 0;
    for (int i = threadIdx.x; i < kVoxPerBlock; i += blockDim.x) {
        long long v = base + i;
        if (v < nvox && img[v] > 0) ++cnt;
    }
    __shared__ int red[8];
    int s = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        if (block_counts) block_counts[blockIdx.x] = t;
        if (total && t) atomicAdd((unsigned long long*)total, (unsigned long long)t);
    }
}

// np.argwhere(img > 0)[k] in C order: walk the per-block counts, then the block's voxels.
__global__ void kth_nonzero_kernel(const unsigned char* __restrict__ img, int d, int h, int w, long long k,
                                   const long long* __restrict__ block_counts, int nblocks, int* __restrict__ coords) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const long long nvox = (long long)d * h * w;
    long long seen = 0;
    int blk = 0;
    for (; blk < nblocks; ++blk) {
        if (seen + block_counts[blk] > k) break;
        seen += block_counts[blk];
    }
    coords[0] = coords[1] = coords[2] = -1;
    if (blk == nblocks) return;
    for (long long v = (long long)blk * kVoxPerBlock; v < nvox; ++v) {
        if (img[v] > 0) {
            if (seen == k) {
                coords[2] = (int)(v % w);
                coords[1] = (int)((v / w) % h);
                coords[0] = (int)(v / ((long long)w * h));
                return;
            }
            ++seen;
        }
    }
}

__global__ void flap_mask_kernel(const unsigned char* __restrict__ img, unsigned char* __restrict__ masked,
                                 unsigned char* __restrict__ extracted, int d, int h, int w,
                                 const int* __restrict__ center, double size, int shape, long long nvox) {
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvox) return;
    const int x = (int)(v % w);
    const int y = (int)((v / w) % h);
    const int z = (int)(v / ((long long)w * h));
    const double dz = (double)(z - center[0]), dy = (double)(y - center[1]), dx = (double)(x - center[2]);
    double dist;
    if (shape == 0) dist = sqrt(dz * dz + dy * dy + dx * dx);          // np.linalg.norm(ord=2), float64
    else dist = fmax(fabs(dz), fmax(fabs(dy), fabs(dx)));              // ord=inf
    const bool inside = dist <= size;                                   // utilities.py:177
    const bool on = img[v] != 0;
    masked[v] = (on && !inside) ? 1 : 0;                                // transforms.py:287
    extracted[v] = (on && inside) ? 1 : 0;                              // transforms.py:294
}

__global__ void hu_window_kernel(const short* __restrict__ hu, float* __restrict__ out, long long nvox, float lo,
                                 float hi) {
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvox) return;
    float x = (float)hu[v];
    x = fminf(fmaxf(x, lo), hi);
    out[v] = __fdiv_rn(__fsub_rn(x, lo), __fsub_rn(hi, lo));
}

__global__ void hu_threshold_kernel(const short* __restrict__ hu, unsigned char* __restrict__ out, long long nvox,
                                    int thr) {
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvox) return;
    out[v] = hu[v] >= thr ? 1 : 0;
}

__device__ __forceinline__ int nearest_src(int dst, int in_size, float scale) {
    int i = (int)floorf(__fmul_rn((float)dst, scale));
    return i < in_size - 1 ? i : in_size - 1;
}

template <typename V>
__global__ void resample_nearest_kernel(const V* __restrict__ src, V* __restrict__ dst, int sd, int sh, int sw, int dd,
                                        int dh, int dw, float fz, float fy, float fx, long long total) {
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= total) return;
    const int x = (int)(v % dw);
    const int y = (int)((v / dw) % dh);
    const int z = (int)(v / ((long long)dw * dh));
    const int iz = nearest_src(z, sd, fz), iy = nearest_src(y, sh, fy), ix = nearest_src(x, sw, fx);
    dst[v] = src[((long long)iz * sh + iy) * sw + ix];
}

__global__ void nearest_index_kernel(int* __restrict__ idx, int out_size, int in_size, float scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < out_size) idx[i] = nearest_src(i, in_size, scale);
}

// F.interpolate(mode='trilinear', align_corners=False): src = scale*(dst+0.5)-0.5 clamped at 0
__device__ __forceinline__ void lin_coord(int dst, int in_size, float scale, int& i0, int& i1, float& l0, float& l1) {
    float s = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
    if (s < 0.f) s = 0.f;
    i0 = (int)s;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = __fsub_rn(s, (float)i0);
    l0 = __fsub_rn(1.f, l1);
}

__global__ void resample_trilinear_kernel(const float* __restrict__ src, float* __restrict__ dst, int sd, int sh, int sw,
                                          int dd, int dh, int dw, float fz, float fy, float fx, long long total) {
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= total) return;
    const int x = (int)(v % dw);
    const int y = (int)((v / dw) % dh);
    const int z = (int)(v / ((long long)dw * dh));
    int z0, z1, y0, y1, x0, x1;
    float lz0, lz1, ly0, ly1, lx0, lx1;
    lin_coord(z, sd, fz, z0, z1, lz0, lz1);
    lin_coord(y, sh, fy, y0, y1, ly0, ly1);
    lin_coord(x, sw, fx, x0, x1, lx0, lx1);
    auto at = [&](int zz, int yy, int xx) { return __ldg(src + ((long long)zz * sh + yy) * sw + xx); };
    const float a = ly0 * (lx0 * at(z0, y0, x0) + lx1 * at(z0, y0, x1)) + ly1 * (lx0 * at(z0, y1, x0) + lx1 * at(z0, y1, x1));
    const float b = ly0 * (lx0 * at(z1, y0, x0) + lx1 * at(z1, y0, x1)) + ly1 * (lx0 * at(z1, y1, x0) + lx1 * at(z1, y1, x1));
    dst[v] = lz0 * a + lz1 * b;
}

}  // namespace ctu

using namespace ctu;

extern "C" {

int ctu_count_nonzero_u8(const unsigned char* img, long long nvox, long long* count, ctu_stream stream) {
    CTU_REQUIRE(img && count && nvox > 0, "ctu_count_nonzero_u8: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(long long), st);
    if (e != cudaSuccess) {
        set_error("ctu_count_nonzero_u8: memset: %s", cudaGetErrorString(e));
        return (int)e;
    }
    count_blocks_kernel<<<cdiv(nvox, kVoxPerBlock), 256, 0, st>>>(img, nvox, nullptr, count);
    return check_launch("ctu_count_nonzero_u8");
}

int ctu_kth_nonzero_u8(const unsigned char* img, int d, int h, int w, long long k, long long* block_counts, int* coords,
                       ctu_stream stream) {
    CTU_REQUIRE(img && block_counts && coords && d > 0 && h > 0 && w > 0 && k >= 0, "ctu_kth_nonzero_u8: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const long long nvox = (long long)d * h * w;
    const int nblocks = cdiv(nvox, kVoxPerBlock);
    count_blocks_kernel<<<nblocks, 256, 0, st>>>(img, nvox, block_counts, nullptr);
    kth_nonzero_kernel<<<1, 32, 0, st>>>(img, d, h, w, k, block_counts, nblocks, coords);
    return check_launch("ctu_kth_nonzero_u8");
}

int ctu_flap_mask_u8(const unsigned char* img, unsigned char* masked, unsigned char* extracted, int d, int h, int w,
                     const int* center, double size, int shape, ctu_stream stream) {
    CTU_REQUIRE(img && masked && extracted && center && d > 0 && h > 0 && w > 0 && (shape == 0 || shape == 1),
                "ctu_flap_mask_u8: bad arguments (shape 0 sphere / 1 box; 'flap' needs raster_geometry: parity unpinned)");
    const long long nvox = (long long)d * h * w;
    flap_mask_kernel<<<cdiv(nvox, 256), 256, 0, (cudaStream_t)stream>>>(img, masked, extracted, d, h, w, center, size, shape, nvox);
    return check_launch("ctu_flap_mask_u8");
}

int ctu_hu_window(const short* hu, float* out, long long nvox, float lo, float hi, ctu_stream stream) {
    CTU_REQUIRE(hu && out && nvox > 0 && hi > lo, "ctu_hu_window: bad arguments");
    hu_window_kernel<<<cdiv(nvox, 256), 256, 0, (cudaStream_t)stream>>>(hu, out, nvox, lo, hi);
    return check_launch("ctu_hu_window");
}

int ctu_hu_threshold(const short* hu, unsigned char* out, long long nvox, int thr, ctu_stream stream) {
    CTU_REQUIRE(hu && out && nvox > 0, "ctu_hu_threshold: bad arguments");
    hu_threshold_kernel<<<cdiv(nvox, 256), 256, 0, (cudaStream_t)stream>>>(hu, out, nvox, thr);
    return check_launch("ctu_hu_threshold");
}

static inline float nscale(int in_size, int out_size) { return (float)in_size / (float)out_size; }

int ctu_resample_nearest_f32(const float* src, float* dst, int sd, int sh, int sw, int dd, int dh, int dw,
                             ctu_stream stream) {
    CTU_REQUIRE(src && dst && sd > 0 && sh > 0 && sw > 0 && dd > 0 && dh > 0 && dw > 0, "ctu_resample_nearest_f32: bad arguments");
    const long long total = (long long)dd * dh * dw;
    resample_nearest_kernel<float><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, sd, sh, sw, dd, dh, dw, nscale(sd, dd), nscale(sh, dh), nscale(sw, dw), total);
    return check_launch("ctu_resample_nearest_f32");
}

int ctu_resample_nearest_u8(const unsigned char* src, unsigned char* dst, int sd, int sh, int sw, int dd, int dh, int dw,
                            ctu_stream stream) {
    CTU_REQUIRE(src && dst && sd > 0 && sh > 0 && sw > 0 && dd > 0 && dh > 0 && dw > 0, "ctu_resample_nearest_u8: bad arguments");
    const long long total = (long long)dd * dh * dw;
    resample_nearest_kernel<unsigned char><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, sd, sh, sw, dd, dh, dw, nscale(sd, dd), nscale(sh, dh), nscale(sw, dw), total);
    return check_launch("ctu_resample_nearest_u8");
}

int ctu_resample_nearest_index(int* idx, int out_size, int in_size, ctu_stream stream) {
    CTU_REQUIRE(idx && out_size > 0 && in_size > 0, "ctu_resample_nearest_index: bad arguments");
    nearest_index_kernel<<<cdiv(out_size, 256), 256, 0, (cudaStream_t)stream>>>(idx, out_size, in_size, nscale(in_size, out_size));
    return check_launch("ctu_resample_nearest_index");
}

int ctu_resample_trilinear_f32(const float* src, float* dst, int sd, int sh, int sw, int dd, int dh, int dw,
                               ctu_stream stream) {
    CTU_REQUIRE(src && dst && sd > 0 && sh > 0 && sw > 0 && dd > 0 && dh > 0 && dw > 0, "ctu_resample_trilinear_f32: bad arguments");
    const long long total = (long long)dd * dh * dw;
    resample_trilinear_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, sd, sh, sw, dd, dh, dw, nscale(sd, dd), nscale(sh, dh), nscale(sw, dw), total);
    return check_launch("ctu_resample_trilinear_f32");
}

}  // extern "C"
