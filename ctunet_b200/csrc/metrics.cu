// Reporting metrics of the loss handlers on the device: hard-label Dice coefficient and Hausdorff distance.
//   utils.dice_coeff   ctunet/utilities.py:53-60   (monai.metrics.compute_meandice on one_hot(argmax(pred)))
//   utils.hausdorff    ctunet/utilities.py:63-70   (monai.metrics.compute_hausdorff_distance, nan/inf -> max(shape))
//   call sites         ctunet/pytorch/ProblemHandler.py:84-88, 277-295  (every stock .ini switches them on)
// monai is an undeclared, unpinned dependency of the reference (not installable here): its published algorithm is
// restated (oracle/ref_stubs/monai/metrics.py is the CPU statement the tests compare with) -> PARITY UNPINNED.
//
// Hausdorff = integer work: edge voxels (mask minus its 6-neighbourhood erosion, outside = background), an EXACT squared
// Euclidean distance transform to the other surface (three separable min-plus passes over int32 squared distances), a max
// over the edge voxels, one sqrt in double at the end.  All HBM/L2-resident byte and int32 traffic; nothing syncs.
#include "common.cuh"

namespace ctu {

constexpr int kInfD2 = 1 << 29;
constexpr int kMetThreads = 256;

// ---------------------------------------------------------------- masks + Dice counts
// item = (b, cls-1): pmask = [argmax_c pred == cls], tmask = [target[cls] == 1]; counts[item] = {sum t*m, sum t, sum m}
template <int C>
__global__ void __launch_bounds__(kMetThreads) seg_masks_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                                long long spatial, unsigned char* __restrict__ pmask,
                                                                unsigned char* __restrict__ tmask, double* __restrict__ counts) {
    const int b = blockIdx.y;
    const float* pb = pred + (long long)b * C * spatial;
    const float* tb = target + (long long)b * C * spatial;
    float inter[C - 1], ysum[C - 1];
    int psum[C - 1];
#pragma unroll
    for (int c = 0; c < C - 1; ++c) inter[c] = ysum[c] = 0.f, psum[c] = 0;
    for (long long v = (long long)blockIdx.x * kMetThreads + threadIdx.x; v < spatial; v += (long long)gridDim.x * kMetThreads) {
        int am = 0;
        float best = __ldg(pb + v);
#pragma unroll
        for (int c = 1; c < C; ++c) {
            const float x = __ldg(pb + c * spatial + v);
            if (x > best || (x != x && best == best)) best = x, am = c;       // torch.argmax: first max, NaN wins
        }
#pragma unroll
        for (int c = 1; c < C; ++c) {
            const float t = __ldg(tb + c * spatial + v);
            const int m = am == c;
            inter[c - 1] += m ? t : 0.f;
            ysum[c - 1] += t;
            psum[c - 1] += m;
            if (pmask) {
                pmask[((long long)b * (C - 1) + (c - 1)) * spatial + v] = (unsigned char)m;
                tmask[((long long)b * (C - 1) + (c - 1)) * spatial + v] = (unsigned char)(t == 1.0f);
            }
        }
    }
    __shared__ double red[kMetThreads / 32][3 * (C - 1)];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < C - 1; ++c) {
        const double a = warp_sum((double)inter[c]), y = warp_sum((double)ysum[c]), p = warp_sum((double)psum[c]);
        if (lane == 0) red[wid][3 * c] = a, red[wid][3 * c + 1] = y, red[wid][3 * c + 2] = p;
    }
    __syncthreads();
    if (threadIdx.x < 3 * (C - 1)) {
        double s = 0;
        for (int w = 0; w < kMetThreads / 32; ++w) s += red[w][threadIdx.x];
        atomicAdd(counts + ((long long)b * (C - 1)) * 3 + threadIdx.x, s);
    }
}

__global__ void dice_coeff_finalize_kernel(const double* __restrict__ counts, int items, float* __restrict__ out) {
    // monai: f = 2*I / (y_o + y_pred_o) in float32, NaN where y_o == 0; the reference takes torch.mean over [B, C-1]
    float acc = 0.f;
    for (int i = 0; i < items; ++i) {
        const float I = (float)counts[3 * i], yo = (float)counts[3 * i + 1], po = (float)counts[3 * i + 2];
        acc += yo > 0.f ? (2.0f * I) / (yo + po) : __int_as_float(0x7fc00000);
    }
    out[0] = acc / (float)items;
}

// ---------------------------------------------------------------- edges
// edge = mask AND NOT (all six face neighbours set); voxels outside the volume are background (scipy binary_erosion,
// border_value = 0).  Sets flags[set] when the set has at least one edge voxel.
__global__ void __launch_bounds__(kMetThreads) mask_edges_kernel(const unsigned char* __restrict__ mask,
                                                                 unsigned char* __restrict__ edge, int d, int h, int w,
                                                                 int* __restrict__ flags) {
    const long long spatial = (long long)d * h * w;
    const int set = blockIdx.y;
    const unsigned char* m = mask + (long long)set * spatial;
    unsigned char* e = edge + (long long)set * spatial;
    int any = 0;
    for (long long v = (long long)blockIdx.x * kMetThreads + threadIdx.x; v < spatial; v += (long long)gridDim.x * kMetThreads) {
        const int x = (int)(v % w), y = (int)((v / w) % h), z = (int)(v / ((long long)w * h));
        unsigned char r = 0;
        if (m[v]) {
            const bool inner = x > 0 && x < w - 1 && y > 0 && y < h - 1 && z > 0 && z < d - 1 && m[v - 1] && m[v + 1] &&
                               m[v - w] && m[v + w] && m[v - (long long)w * h] && m[v + (long long)w * h];
            r = inner ? 0 : 1;
        }
        e[v] = r;
        any |= r;
    }
    any = __syncthreads_or(any);
    if (threadIdx.x == 0 && any) atomicOr(flags + set, 1);
}

// ---------------------------------------------------------------- squared EDT, pass 1 (along w)
// g[z][y][x] = min over edge voxels x' of the row of (x - x')^2, kInfD2 if the row has none.
__global__ void edt_row_kernel(const unsigned char* __restrict__ edge, int* __restrict__ g, int w, long long rows) {
    extern __shared__ unsigned char srow[];
    const int rpb = blockDim.y;
    const long long row = (long long)blockIdx.x * rpb + threadIdx.y;
    unsigned char* r = srow + threadIdx.y * w;
    if (row < rows)
        for (int x = threadIdx.x; x < w; x += blockDim.x) r[x] = edge[row * w + x];
    __syncthreads();
    if (row >= rows) return;
    for (int x = threadIdx.x; x < w; x += blockDim.x) {
        int best = kInfD2;
        // nearest set voxel left and right of x: walk outwards, stop at the first hit
        for (int k = 0; k < w; ++k) {
            const int a = x - k, b = x + k;
            if ((a >= 0 && r[a]) || (b < w && r[b])) {
                best = k * k;
                break;
            }
            if (a < 0 && b >= w) break;
        }
        g[row * w + x] = best;
    }
}

// ---------------------------------------------------------------- squared EDT, passes 2 and 3 (min-plus along L)
// Volume viewed as [outer][L][inner]: out[o][l][i] = min_l' (in[o][l'][i] + (l - l')^2).  A block owns (o, a tile of 32
// inner positions): the whole column tile sits in shared memory.  FINAL: instead of storing, reduce the max over the
// voxels of the PARTNER edge set (directed Hausdorff distance, squared) into maxd2[set].
template <bool FINAL>
__global__ void edt_axis_kernel(const int* __restrict__ in, int* __restrict__ out, int L, long long inner, long long outer,
                                long long set_stride, const unsigned char* __restrict__ edge, int* __restrict__ maxd2) {
    extern __shared__ int col[];      // [L][32]
    const int set = blockIdx.z;
    const long long tiles = (inner + 31) / 32;
    const long long o = blockIdx.x / tiles;
    const long long i0 = (blockIdx.x % tiles) * 32;
    const long long base = (long long)set * set_stride + o * L * inner;
    const int lx = threadIdx.x;                    // inner lane
    const bool ok = i0 + lx < inner;
    for (int l = threadIdx.y; l < L; l += blockDim.y) col[l * 32 + lx] = ok ? in[base + (long long)l * inner + i0 + lx] : kInfD2;
    __syncthreads();
    int local_max = -1;
    for (int l = threadIdx.y; l < L; l += blockDim.y) {
        int best = col[l * 32 + lx];
        // candidates further away than sqrt(best) cannot win: shrink the window as best improves
        for (int k = 1; k < L; ++k) {
            const int kk = k * k;
            if (kk >= best) break;
            const int a = l - k, b = l + k;
            if (a >= 0) best = min(best, col[a * 32 + lx] + kk);
            if (b < L) best = min(best, col[b * 32 + lx] + kk);
            if (a < 0 && b >= L) break;
        }
        if (!ok) continue;
        const long long idx = o * L * inner + (long long)l * inner + i0 + lx;
        if (FINAL) {
            const unsigned char* pe = edge + (long long)(set ^ 1) * set_stride;      // the partner set's edges
            if (pe[idx]) local_max = max(local_max, best);
        } else {
            out[(long long)set * set_stride + idx] = best;
        }
    }
    if (FINAL) {
        for (int off = 16; off > 0; off >>= 1) local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, off));
        if (lx == 0 && local_max >= 0) atomicMax(maxd2 + set, local_max);
    }
}

// sets come in pairs (2p = prediction, 2p+1 = target).  maxd2[2p] = max over target-edge voxels of d2 to the prediction
// surface, maxd2[2p+1] the other direction.  Empty surface on either side -> NaN / inf in monai -> inf_alt (utilities.py:69).
__global__ void hausdorff_finalize_kernel(const int* __restrict__ maxd2, const int* __restrict__ flags, int pairs,
                                          double inf_alt, double* __restrict__ out) {
    double acc = 0;
    for (int p = 0; p < pairs; ++p) {
        double hd = inf_alt;
        if (flags[2 * p] && flags[2 * p + 1]) hd = sqrt((double)max(maxd2[2 * p], maxd2[2 * p + 1]));
        acc += hd;
    }
    out[0] = acc / pairs;
}

template <int C>
static int launch_masks(const float* pred, const float* target, int b, long long spatial, unsigned char* pm, unsigned char* tm,
                        double* counts, cudaStream_t st) {
    const int blocks = (int)min((long long)2048, (spatial + kMetThreads - 1) / kMetThreads);
    seg_masks_kernel<C><<<dim3(blocks, b), kMetThreads, 0, st>>>(pred, target, spatial, pm, tm, counts);
    return check_launch("seg_masks_kernel");
}

}  // namespace ctu

using namespace ctu;

extern "C" int ctu_dice_coeff(const float* pred, const float* target, int b, int c, long long spatial, double* counts,
                              float* out, ctu_stream stream) {
    CTU_REQUIRE(pred && target && counts && out, "ctu_dice_coeff: null pointer");
    CTU_REQUIRE(b >= 1 && spatial >= 1 && c >= 2 && c <= 4, "ctu_dice_coeff: b=%d c=%d (2..4 channels)", b, c);
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(counts, 0, sizeof(double) * 3 * b * (c - 1), st);
    int rc = c == 2 ? launch_masks<2>(pred, target, b, spatial, nullptr, nullptr, counts, st)
           : c == 3 ? launch_masks<3>(pred, target, b, spatial, nullptr, nullptr, counts, st)
                    : launch_masks<4>(pred, target, b, spatial, nullptr, nullptr, counts, st);
    if (rc) return rc;
    dice_coeff_finalize_kernel<<<1, 1, 0, st>>>(counts, b * (c - 1), out);
    return check_launch("dice_coeff_finalize_kernel");
}

extern "C" long long ctu_hausdorff_workspace_bytes(int b, int c, int d, int h, int w) {
    if (b < 1 || c < 2 || d < 1 || h < 1 || w < 1) return -1;
    const long long sets = 2LL * b * (c - 1), spatial = (long long)d * h * w;
    // masks + edges (1 B each) + two int32 distance planes per set, + counts / flags / maxima
    return sets * spatial * (2 + 8) + 8192 + 64LL * sets;
}

extern "C" int ctu_hausdorff(const float* pred, const float* target, int b, int c, int d, int h, int w, double inf_alt,
                             void* workspace, long long workspace_bytes, double* out, ctu_stream stream) {
    CTU_REQUIRE(pred && target && workspace && out, "ctu_hausdorff: null pointer");
    CTU_REQUIRE(b >= 1 && c >= 2 && c <= 4, "ctu_hausdorff: b=%d c=%d (2..4 channels)", b, c);
    CTU_REQUIRE(d >= 1 && h >= 1 && w >= 1 && d <= 1024 && h <= 1024 && w <= 1024, "ctu_hausdorff: dims %dx%dx%d (<= 1024)", d, h, w);
    CTU_REQUIRE(workspace_bytes >= ctu_hausdorff_workspace_bytes(b, c, d, h, w), "ctu_hausdorff: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const long long spatial = (long long)d * h * w;
    const int pairs = b * (c - 1), sets = 2 * pairs;
    // workspace carving.  Set order: interleave (prediction, target) per pair so that partner(set) = set ^ 1.
    char* p = (char*)workspace;
    double* counts = (double*)p;              p += 2048;
    int* flags = (int*)p;                     p += 1024;
    int* maxd2 = (int*)p;                     p += 1024;
    p += 64LL * sets - 0;                     // slack (keeps the planes 64-byte aligned for any set count)
    p = (char*)(((uintptr_t)p + 255) & ~(uintptr_t)255);
    unsigned char* masks = (unsigned char*)p; p += sets * spatial;
    unsigned char* edges = (unsigned char*)p; p += sets * spatial;
    p = (char*)(((uintptr_t)p + 255) & ~(uintptr_t)255);
    int* g0 = (int*)p;                        p += sets * spatial * 4;
    int* g1 = (int*)p;
    CTU_REQUIRE(3 * pairs * sizeof(double) <= 2048 && sets * sizeof(int) <= 1024, "ctu_hausdorff: too many (batch, class) pairs");
    cudaMemsetAsync(workspace, 0, 4096, st);
    // masks: kernel writes [b][c-1][spatial] for predictions and targets separately; interleave by using strided bases:
    // prediction mask of pair i at set 2i, target mask at set 2i+1  ->  write through two temporaries = the g1 plane
    unsigned char* pm = (unsigned char*)g1;
    unsigned char* tm = pm + (long long)pairs * spatial;
    int rc = c == 2 ? launch_masks<2>(pred, target, b, spatial, pm, tm, counts, st)
           : c == 3 ? launch_masks<3>(pred, target, b, spatial, pm, tm, counts, st)
                    : launch_masks<4>(pred, target, b, spatial, pm, tm, counts, st);
    if (rc) return rc;
    cudaMemcpy2DAsync(masks, 2 * spatial, pm, spatial, spatial, pairs, cudaMemcpyDeviceToDevice, st);
    cudaMemcpy2DAsync(masks + spatial, 2 * spatial, tm, spatial, spatial, pairs, cudaMemcpyDeviceToDevice, st);
    const int blocks = (int)min((long long)1024, (spatial + kMetThreads - 1) / kMetThreads);
    mask_edges_kernel<<<dim3(blocks, sets), kMetThreads, 0, st>>>(masks, edges, d, h, w, flags);
    if ((rc = check_launch("mask_edges_kernel"))) return rc;
    {   // pass 1 along w (all sets at once: rows = sets * d * h)
        const long long rows = (long long)sets * d * h;
        const int tx = w >= 128 ? 128 : (w >= 64 ? 64 : 32), ty = 256 / tx;
        edt_row_kernel<<<(unsigned)((rows + ty - 1) / ty), dim3(tx, ty), (size_t)ty * w, st>>>(edges, g0, w, rows);
        if ((rc = check_launch("edt_row_kernel"))) return rc;
    }
    {   // pass 2 along h: [outer = d][L = h][inner = w]
        const long long tiles = (w + 31) / 32;
        const int ty = 8;
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(edt_axis_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 * 32 * 4);
            cudaFuncSetAttribute(edt_axis_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 * 32 * 4);
            attr = true;
        }
        edt_axis_kernel<false><<<dim3((unsigned)(d * tiles), 1, sets), dim3(32, ty), (size_t)h * 32 * 4, st>>>(
            g0, g1, h, w, d, spatial, nullptr, nullptr);
        if ((rc = check_launch("edt_axis_kernel<h>"))) return rc;
        // pass 3 along d: [outer = 1][L = d][inner = h*w], fused with the max over the partner's edge voxels
        const long long inner = (long long)h * w, tiles3 = (inner + 31) / 32;
        edt_axis_kernel<true><<<dim3((unsigned)tiles3, 1, sets), dim3(32, ty), (size_t)d * 32 * 4, st>>>(
            g1, nullptr, d, inner, 1, spatial, edges, maxd2);
        if ((rc = check_launch("edt_axis_kernel<d>"))) return rc;
    }
    hausdorff_finalize_kernel<<<1, 1, 0, st>>>(maxd2, flags, pairs, inf_alt, out);
    return check_launch("hausdorff_finalize_kernel");
}
