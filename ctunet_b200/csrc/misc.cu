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

// shape_3d(shape="flap") (utilities.py:145-166): union of two cylinders along d (raster_geometry.cylinder, axis 0) and
// a cube (raster_geometry.cube).  raster_geometry is an un-vendored, unpinned dependency of the reference and is not
// installed here, so its published algorithm is RESTATED -- parity unpinned: relative positions become absolute grid
// origins x0 = round((dim - 1) * rel) (round half to even), a cylinder is {(y-y0)^2 + (x-x0)^2 <= radius^2 and
// |z-z0| <= height/2}, a cube is {|c - c0| <= side/2 on every axis}, all in float64.
__global__ void flap_shape_mask_kernel(const unsigned char* __restrict__ img, unsigned char* __restrict__ masked,
                                       unsigned char* __restrict__ extracted, int d, int h, int w,
                                       const int* __restrict__ center, double size, double c_diam, long long nvox) {
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvox) return;
    const int x = (int)(v % w);
    const int y = (int)((v / w) % h);
    const int z = (int)(v / ((long long)w * h));
    const double c0 = (double)center[0], c1 = (double)center[1], c2 = (double)center[2];
    // center_relative, z_edge_1, z_edge_2 (utilities.py:148-158) -> absolute origins
    const double z0 = rint((d - 1) * (c0 / d));
    const double ey = rint((h - 1) * ((c1 - size / 2) / h));
    const double ex1 = rint((w - 1) * ((c2 - size / 2) / w)), ex2 = rint((w - 1) * ((c2 + size / 2) / w));
    const double qy = rint((h - 1) * (c1 / h)), qx = rint((w - 1) * (c2 / w));
    const double dz = z - z0;
    const bool in_height = fabs(dz) <= size / 2;
    const double r2 = c_diam * c_diam, dy = y - ey;
    const bool cyl1 = in_height && (dy * dy + (x - ex1) * (x - ex1) <= r2);
    const bool cyl2 = in_height && (dy * dy + (x - ex2) * (x - ex2) <= r2);
    const bool cub = in_height && fabs(y - qy) <= size / 2 && fabs(x - qx) <= size / 2;
    const bool inside = cyl1 || cyl2 || cub;                            // mask; shape_np = 1 - mask (utilities.py:165-166)
    const bool on = img[v] != 0;
    masked[v] = (on && !inside) ? 1 : 0;
    extracted[v] = (on && inside) ? 1 : 0;
}

// datasets.py:195-235 on the device: image channel 0 = the broken skull as float, channel 1 = the atlas
// (load_atlas_and_append_at_axis, datasets.py:30-47); targets = one_hot(label, 2) as float32 (datasets.py:209-214).
// 16 voxels per thread: one 16-byte load per mask, 16-byte stores.
__global__ void encode_flaprec_kernel(const unsigned char* __restrict__ broken, const unsigned char* __restrict__ full,
                                      const unsigned char* __restrict__ flap, const float* __restrict__ atlas,
                                      float* __restrict__ image, float* __restrict__ skull_t, float* __restrict__ flap_t,
                                      int cin, long long spatial) {
    const int b = blockIdx.y;
    const long long v0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (v0 >= spatial) return;
    const uint4 ub = *reinterpret_cast<const uint4*>(broken + b * spatial + v0);
    const uint4 uf = *reinterpret_cast<const uint4*>(full + b * spatial + v0);
    const uint4 ul = *reinterpret_cast<const uint4*>(flap + b * spatial + v0);
    const uint32_t wb[4] = {ub.x, ub.y, ub.z, ub.w}, wf[4] = {uf.x, uf.y, uf.z, uf.w}, wl[4] = {ul.x, ul.y, ul.z, ul.w};
    float* img0 = image + (long long)b * cin * spatial + v0;
    float* sk0 = skull_t + (long long)b * 2 * spatial + v0;
    float* fl0 = flap_t + (long long)b * 2 * spatial + v0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float vb[4], vf[4], vl[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            vb[j] = (float)((wb[i] >> (8 * j)) & 0xffu);                 // tensor.float(): the value itself
            vf[j] = ((wf[i] >> (8 * j)) & 0xffu) ? 1.f : 0.f;            // one_hot class of a {0,1} label
            vl[j] = ((wl[i] >> (8 * j)) & 0xffu) ? 1.f : 0.f;
        }
        *reinterpret_cast<float4*>(img0 + 4 * i) = make_float4(vb[0], vb[1], vb[2], vb[3]);
        *reinterpret_cast<float4*>(sk0 + 4 * i) = make_float4(1.f - vf[0], 1.f - vf[1], 1.f - vf[2], 1.f - vf[3]);
        *reinterpret_cast<float4*>(sk0 + spatial + 4 * i) = make_float4(vf[0], vf[1], vf[2], vf[3]);
        *reinterpret_cast<float4*>(fl0 + 4 * i) = make_float4(1.f - vl[0], 1.f - vl[1], 1.f - vl[2], 1.f - vl[3]);
        *reinterpret_cast<float4*>(fl0 + spatial + 4 * i) = make_float4(vl[0], vl[1], vl[2], vl[3]);
        if (cin > 1) *reinterpret_cast<float4*>(img0 + spatial + 4 * i) = *reinterpret_cast<const float4*>(atlas + v0 + 4 * i);
    }
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
                     const int* center, double size, int shape, double c_diam, ctu_stream stream) {
    CTU_REQUIRE(img && masked && extracted && center && d > 0 && h > 0 && w > 0 && shape >= 0 && shape <= 2,
                "ctu_flap_mask_u8: bad arguments (shape 0 sphere / 1 box / 2 flap)");
    const long long nvox = (long long)d * h * w;
    if (shape == 2)
        flap_shape_mask_kernel<<<cdiv(nvox, 256), 256, 0, (cudaStream_t)stream>>>(img, masked, extracted, d, h, w, center, size, c_diam, nvox);
    else
        flap_mask_kernel<<<cdiv(nvox, 256), 256, 0, (cudaStream_t)stream>>>(img, masked, extracted, d, h, w, center, size, shape, nvox);
    return check_launch("ctu_flap_mask_u8");
}

int ctu_encode_flaprec_u8(const unsigned char* broken, const unsigned char* full, const unsigned char* flap,
                          const float* atlas, float* image, float* skull_target, float* flap_target, int batch,
                          int in_channels, long long spatial, ctu_stream stream) {
    CTU_REQUIRE(broken && full && flap && image && skull_target && flap_target && batch > 0 && spatial > 0,
                "ctu_encode_flaprec_u8: bad arguments");
    CTU_REQUIRE(in_channels == 1 || (in_channels == 2 && atlas), "ctu_encode_flaprec_u8: 1 input channel, or 2 with an atlas");
    CTU_REQUIRE(spatial % 16 == 0, "ctu_encode_flaprec_u8: the volume size must be a multiple of 16 voxels");
    dim3 grid(cdiv(spatial / 16, 256), batch);
    encode_flaprec_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(broken, full, flap, atlas, image, skull_target, flap_target,
                                                                  in_channels, spatial);
    return check_launch("ctu_encode_flaprec_u8");
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
