// tcgen05 / TMEM / TMA implicit-GEMM 3-D convolution for sm_100a (bf16 in, fp32 accumulate).
//
// Forward and data-gradient of nn.Conv3d k in {3,5}, stride 1, "same" padding
// (ctunet/pytorch/models.py:26,29,38,41,403,407,430,434) on the blocked layout [N][Cb][D][H][W][8].
//
// Mapping (DESIGN.md "conv_tc"):
//   * GEMM M = 128 output voxels = 16 (h) x 8 (w) of one d-plane; N = Cout padded to 16; K = taps x Cin.
//   * One voxel x 8 channels is 16 B, so 8 consecutive w-voxels of one channel block ARE a no-swizzle
//     K-major UMMA core matrix (8 rows x 16 B).  A halo plane [Cb][18][18][8] (k=3) staged by TMA serves
//     every tap by descriptor arithmetic only: start address = tap shift, SBO = row pitch (next h), LBO =
//     distance to the second 8-wide K chunk (the next channel block, or the next tap).  No im2col copy.
//   * The TMA box is taken from a 4-D view [N*Cb][D][H][W*8]; out-of-range coordinates are zero-filled by the
//     hardware, which is exactly the convolution's zero padding (also across the d border).
//   * A CTA walks a d-chunk of one (n, 16x16 h-w tile): ring of K+1 planes in shared memory, so every input
//     plane is fetched once per tile; two TMEM accumulator stages overlap MMA with the epilogue.
//   * Warp roles: warp 0 TMA producer, warp 1 MMA issuer (one elected lane), warps 2-5 epilogue
//     (tcgen05.ld -> bias -> bf16 -> 16-byte stores, plus per-channel sum / sum-of-squares for the following
//     BatchNorm, so the statistics pass over y disappears).
#include "common.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>

namespace ctu {

constexpr int TC_THREADS = 192;
constexpr int TC_TH = 16, TC_TW = 16;   // output tile (h, w) per plane = 2 MMA tiles of 16x8
constexpr int TC_WB = TC_TW / 8;

// ------------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a CUDA error, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, no-swizzle shared-memory matrix descriptor (sm_100 version bit set).
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// ------------------------------------------------------------------------------------------------ schedule
// MMAs of one kd-plane: first every (tap, channel-block pair) -- second K chunk = the next channel block --
// then, when Cb is odd, the last channel block's taps pairwise -- second K chunk = the next tap.
__host__ __device__ inline int tc_mmas_per_kd(int k, int cb) {
    const int k2 = k * k;
    return k2 * (cb / 2) + ((cb & 1) ? (k2 + 1) / 2 : 0);
}
// chunk c (0/1) of MMA m: returns false for the zero-weight dummy half of an odd tail
__host__ __device__ inline bool tc_chunk(int k, int cb, int m, int c, int& blk, int& tap2d) {
    const int k2 = k * k, pairs = cb / 2;
    if (m < k2 * pairs) {
        tap2d = m / pairs;
        blk = 2 * (m % pairs) + c;
        return true;
    }
    const int j = m - k2 * pairs;
    blk = cb - 1;
    tap2d = 2 * j + c;
    if (tap2d >= k2) {
        tap2d = k2 - 1;
        return false;
    }
    return true;
}

struct TcParams {
    const __nv_bfloat16* wimg;   // [K kd][mmas_per_kd][N/8][2][8][8] bf16 (the UMMA B tiles, in order of use)
    const float* bias;
    __nv_bfloat16* y;
    double* stats;               // nullable: [2][cpad_out] sum, sum of squares (of the bf16-rounded outputs)
    int cb, cob_n, npad, cout;
    int n, d, h, w;
    int tiles_h, tiles_w, dchunks, dc;
    int total_items;
    uint32_t wimg_bytes, plane_bytes, slot_bytes;   // plane_bytes: one channel block of one plane, padded to 128
    uint32_t tmem_cols;
};

// fp32 packed weights [cob][cib][tap][ci][co] -> the bf16 B-tile image above
__global__ void tc_pack_wimg_kernel(const float* __restrict__ wp, __nv_bfloat16* __restrict__ wimg, int k, int cb,
                                    int cob_n, int npad, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int e = (int)(i & 7), r = (int)((i >> 3) & 7), c = (int)((i >> 6) & 1);
    long long q = i >> 7;
    const int ngroups = npad / 8;
    const int g = (int)(q % ngroups);
    q /= ngroups;
    const int nm = tc_mmas_per_kd(k, cb);
    const int m = (int)(q % nm);
    const int kd = (int)(q / nm);
    int blk, tap2d;
    float v = 0.f;
    if (tc_chunk(k, cb, m, c, blk, tap2d) && g < cob_n) {
        const int taps = k * k * k, tap = kd * k * k + tap2d;
        v = wp[(((long long)g * cb + blk) * taps + tap) * 64 + e * 8 + r];   // e = input lane (K), r = output lane (N)
    }
    wimg[i] = __float2bfloat16_rn(v);
}

// Epilogue of one 8-channel output block of one voxel row: TMEM -> (+bias) -> bf16 -> 16-byte store (+ statistics).
template <bool CAN_STATS>
__device__ __forceinline__ void tc_emit(uint32_t taddr, int ob, const TcParams& p, __nv_bfloat16* ybase, long long plane,
                                        bool inb, bool want_stats, float (&s1)[8], float (&s2)[8]) {
    float v[8];
    tmem_ld8(taddr + ob * 8, v);
    V8 o;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int ch = ob * 8 + c;
        float x = v[c];
        if (p.bias != nullptr && ch < p.cout) x += __ldg(p.bias + ch);
        o.v[c] = round_to<__nv_bfloat16>(x);
    }
    if (inb) {
        Vec8<__nv_bfloat16>::store(ybase + (long long)ob * plane * 8, o);
        if (CAN_STATS && want_stats) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                s1[c] += o.v[c];
                s2[c] = fmaf(o.v[c], o.v[c], s2[c]);
            }
        }
    }
}

template <int K>
__global__ void __launch_bounds__(TC_THREADS) conv3d_tc_kernel(const __grid_constant__ CUtensorMap tmap, TcParams p) {
    constexpr int PAD = K / 2;
    constexpr int NS = K + 1;                      // plane ring
    constexpr int HH = TC_TH + K - 1, WW = TC_TW + K - 1;
    constexpr uint32_t ROW = WW * 16;              // bytes per halo row of one channel block
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_w = s_base;                                            // weights image
    const uint32_t s_planes = s_base + ((p.wimg_bytes + 1023u) & ~1023u);   // NS slots
    const uint32_t s_tab = s_planes + NS * p.slot_bytes;                    // per-MMA A descriptor low words
    const int nm = tc_mmas_per_kd(K, p.cb);
    uint32_t* tab = reinterpret_cast<uint32_t*>(smem + (s_tab - s_base));
    const uint32_t s_bar = s_tab + ((nm * 4u + 15u) & ~15u);
    // barriers: plane_full[NS], plane_empty[NS], acc_full[2], acc_empty[2], w_full
    const uint32_t b_full = s_bar, b_empty = s_bar + 8 * NS, b_afull = s_bar + 16 * NS, b_aempty = b_afull + 16,
                   b_w = b_afull + 32;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (b_w + 8 - s_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; ++i) {
            mbar_init(b_full + 8 * i, 1);
            mbar_init(b_empty + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(b_afull + 8 * i, 1);
            mbar_init(b_aempty + 8 * i, 4);
        }
        mbar_init(b_w, 1);
        fence_barrier_init();
    }
    // A-descriptor table: low word = (offset within a plane slot) >> 4 | (LBO >> 4) << 16
    for (int m = threadIdx.x; m < nm; m += TC_THREADS) {
        int blk0, t0, blk1, t1;
        tc_chunk(K, p.cb, m, 0, blk0, t0);
        const bool real1 = tc_chunk(K, p.cb, m, 1, blk1, t1);
        const uint32_t off0 = blk0 * p.plane_bytes + (t0 / K) * ROW + (t0 % K) * 16;
        const uint32_t off1 = blk1 * p.plane_bytes + (t1 / K) * ROW + (t1 % K) * 16;
        const uint32_t lbo = real1 ? off1 - off0 : 0u;
        tab[m] = (off0 >> 4) | ((lbo >> 4) << 16);
    }
    if (warp == 1) {   // TMEM allocation (this warp also frees it)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int items_per_n = p.tiles_h * p.tiles_w * p.dchunks;

    if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            mbar_expect_tx(b_w, p.wimg_bytes);
            // weights in <= 64 KB pieces (bulk copy size field)
            for (uint32_t off = 0; off < p.wimg_bytes; off += 32768u) {
                const uint32_t sz = p.wimg_bytes - off < 32768u ? p.wimg_bytes - off : 32768u;
                bulk_load_1d(s_w + off, reinterpret_cast<const unsigned char*>(p.wimg) + off, sz, b_w);
            }
            uint32_t it = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                const int n = item / items_per_n;
                int r = item % items_per_n;
                const int dci = r % p.dchunks; r /= p.dchunks;
                const int twi = r % p.tiles_w, thi = r / p.tiles_w;
                const int z0 = dci * p.dc, h0 = thi * TC_TH, w0 = twi * TC_TW;
                const int nd = (p.d - z0) < p.dc ? (p.d - z0) : p.dc;
                for (int pl = 0; pl < nd + K - 1; ++pl, ++it) {
                    const uint32_t slot = it % NS, ph = (it / NS) & 1;
                    mbar_wait(b_empty + 8 * slot, ph ^ 1);
                    mbar_expect_tx(b_full + 8 * slot, (uint32_t)p.cb * HH * WW * 16);
                    for (int b = 0; b < p.cb; ++b)
                        tma_load_4d(s_planes + slot * p.slot_bytes + b * p.plane_bytes, &tmap, (w0 - PAD) * 8, h0 - PAD,
                                    z0 + pl - PAD, n * p.cb + b, b_full + 8 * slot);
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=bf16, both K-major, N at [17,23), M=128 at [24,29)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.npad >> 3) << 17) | (8u << 24);
            const uint32_t a_hi = (ROW >> 4) | (1u << 14);           // SBO = row pitch (next h), version 1
            const uint32_t b_hi = (16u) | (1u << 14);                 // B: SBO = 256 B between n-groups
            const uint32_t b_lbo = 8u << 16;                          // B: LBO = 128 B between the two K chunks
            const uint32_t btile16 = (uint32_t)p.npad * 2;            // one B tile = npad*32 bytes
            mbar_wait(b_w, 0);
            uint32_t base = 0, waited = 0, acc_it = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                const int r0 = item % items_per_n;
                const int dci = r0 % p.dchunks;
                const int z0 = dci * p.dc;
                const int nd = (p.d - z0) < p.dc ? (p.d - z0) : p.dc;
                for (int j = 0; j < nd; ++j, ++acc_it) {
                    while (waited < base + j + K) {
                        mbar_wait(b_full + 8 * (waited % NS), (waited / NS) & 1);
                        ++waited;
                    }
                    const uint32_t stage = acc_it & 1;
                    mbar_wait(b_aempty + 8 * stage, ((acc_it >> 1) & 1) ^ 1);
                    tc_fence_after();
#pragma unroll 1
                    for (int t = 0; t < TC_WB; ++t) {
                        const uint32_t d_tmem = tmem_base + (stage * TC_WB + t) * p.npad;
                        uint32_t first = 0;
#pragma unroll 1
                        for (int kd = 0; kd < K; ++kd) {
                            const uint32_t slot = (base + j + kd) % NS;
                            const uint32_t a16 = (s_planes + slot * p.slot_bytes + t * 128u) >> 4;
                            uint32_t b16 = (s_w >> 4) + (uint32_t)(kd * nm) * btile16;
#pragma unroll 1
                            for (int m = 0; m < nm; ++m, b16 += btile16) {
                                const uint64_t ad = ((uint64_t)a_hi << 32) | (uint64_t)(tab[m] + a16);
                                const uint64_t bd = ((uint64_t)b_hi << 32) | (uint64_t)(b16 | b_lbo);
                                umma_bf16(d_tmem, ad, bd, idesc, first);
                                first = 1;
                            }
                        }
                    }
                    umma_commit(b_afull + 8 * stage);
                    umma_commit(b_empty + 8 * ((base + j) % NS));   // plane j is not needed by later outputs
                }
                for (int q = 0; q < K - 1; ++q) umma_commit(b_empty + 8 * ((base + nd + q) % NS));
                base += nd + K - 1;
            }
        }
    } else {
        // ===================================================================== epilogue (warps 2..5)
        const int quarter = warp & 3;                  // TMEM lane quarter this warp may read
        const int row = quarter * 32 + lane;           // row of the 128-row MMA tile
        const int hh = row >> 3, wl = row & 7;
        float sA[8], qA[8], sB[8], qB[8];              // per-thread BatchNorm statistics of output blocks 0 and 1
        const bool want_stats = p.stats != nullptr;
        const long long plane = (long long)p.d * p.h * p.w;
        uint32_t acc_it = 0;   // statistics are fused only for cob_n <= 2 (the host splits the rest off)
#pragma unroll
        for (int i = 0; i < 8; ++i) sA[i] = qA[i] = sB[i] = qB[i] = 0.f;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            const int n = item / items_per_n;
            int r = item % items_per_n;
            const int dci = r % p.dchunks; r /= p.dchunks;
            const int twi = r % p.tiles_w, thi = r / p.tiles_w;
            const int z0 = dci * p.dc, h0 = thi * TC_TH, w0 = twi * TC_TW;
            const int nd = (p.d - z0) < p.dc ? (p.d - z0) : p.dc;
            for (int j = 0; j < nd; ++j, ++acc_it) {
                const uint32_t stage = acc_it & 1;
                mbar_wait(b_afull + 8 * stage, (acc_it >> 1) & 1);
                tc_fence_after();
                const int gz = z0 + j, gy = h0 + hh;
#pragma unroll 1
                for (int t = 0; t < TC_WB; ++t) {
                    const int gx = w0 + t * 8 + wl;
                    const bool inb = gy < p.h && gx < p.w;
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (stage * TC_WB + t) * p.npad;
                    __nv_bfloat16* ybase = p.y + (((long long)n * p.cob_n) * plane + ((long long)gz * p.h + gy) * p.w + gx) * 8;
                    tc_emit<true>(taddr, 0, p, ybase, plane, inb, want_stats, sA, qA);
                    if (p.cob_n > 1) tc_emit<true>(taddr, 1, p, ybase, plane, inb, want_stats, sB, qB);
                    for (int ob = 2; ob < p.cob_n; ++ob) tc_emit<false>(taddr, ob, p, ybase, plane, inb, false, sA, qA);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(b_aempty + 8 * stage);
            }
        }
        if (want_stats) {
            const int cpad = p.cob_n * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float a1 = warp_sum(sA[i]), a2 = warp_sum(qA[i]);
                const float b1 = warp_sum(sB[i]), b2 = warp_sum(qB[i]);
                if (lane == 0) {
                    if (i < p.cout) {
                        atomicAdd(p.stats + i, (double)a1);
                        atomicAdd(p.stats + cpad + i, (double)a2);
                    }
                    if (8 + i < p.cout) {
                        atomicAdd(p.stats + 8 + i, (double)b1);
                        atomicAdd(p.stats + cpad + 8 + i, (double)b2);
                    }
                }
            }
        }
    }
    // ------------------------------------------------------------------------- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    }
    return fn;
}

struct TcGeom {
    int cb, cob_n, npad, nm;
    uint32_t wimg_bytes, plane_bytes, slot_bytes, tmem_cols;
    size_t smem;
};

static bool tc_geometry(int k, int cin, int cout, int h, int w, TcGeom& g) {
    if (k != 3 && k != 5) return false;
    if (h % TC_TH || w % TC_TW) return false;
    g.cb = (cin + 7) / 8;
    g.cob_n = (cout + 7) / 8;
    g.npad = (cout + 15) / 16 * 16;
    if (g.npad > 128) return false;
    g.nm = tc_mmas_per_kd(k, g.cb);
    g.wimg_bytes = (uint32_t)k * g.nm * g.npad * 32;
    const int hh = TC_TH + k - 1, ww = TC_TW + k - 1;
    g.plane_bytes = ((uint32_t)hh * ww * 16 + 127u) & ~127u;
    g.slot_bytes = g.plane_bytes * g.cb;
    uint32_t cols = 2 * TC_WB * g.npad;
    g.tmem_cols = 32;
    while (g.tmem_cols < cols) g.tmem_cols *= 2;
    if (g.tmem_cols > 512) return false;
    g.smem = ((g.wimg_bytes + 1023u) & ~1023u) + (size_t)(k + 1) * g.slot_bytes + ((g.nm * 4 + 15) & ~15) + 8 * (2 * (k + 1) + 5) + 16 + 1024;
    return g.smem <= 220 * 1024;
}

int conv3d_fprop_tc(const void* const* h_srcs, const int* h_src_channels, int nsrc, const float* wp, const float* bias,
                    void* y, double* stats, int cout, int k, int n, int d, int h, int w, cudaStream_t stream) {
    if (nsrc != 1) {
        set_error("conv3d tensor path: single source only");
        return CTU_ERR_UNSUPPORTED;
    }
    TcGeom g;
    if (!tc_geometry(k, h_src_channels[0], cout, h, w, g)) {
        set_error("conv3d tensor path: shape k=%d cin=%d cout=%d %dx%dx%d not covered", k, h_src_channels[0], cout, d, h, w);
        return CTU_ERR_UNSUPPORTED;
    }
    auto encode = get_encode();
    if (!encode) {
        set_error("cuTensorMapEncodeTiled entry point not found");
        return CTU_ERR_UNSUPPORTED;
    }
    CUtensorMap tmap;
    const cuuint64_t gdim[4] = {(cuuint64_t)w * 8, (cuuint64_t)h, (cuuint64_t)d, (cuuint64_t)n * g.cb};
    const cuuint64_t gstr[3] = {(cuuint64_t)w * 16, (cuuint64_t)h * w * 16, (cuuint64_t)d * h * w * 16};
    const cuuint32_t box[4] = {(cuuint32_t)(TC_TW + k - 1) * 8, (cuuint32_t)(TC_TH + k - 1), 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(h_srcs[0]), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d)", (int)cr);
        return CTU_ERR_INVALID;
    }
    TcParams p = {};
    p.wimg = reinterpret_cast<const __nv_bfloat16*>(wp);
    p.bias = bias;
    p.y = reinterpret_cast<__nv_bfloat16*>(y);
    const bool fuse_stats = stats != nullptr && g.cob_n <= 2;
    p.stats = fuse_stats ? stats : nullptr;
    p.cb = g.cb; p.cob_n = g.cob_n; p.npad = g.npad; p.cout = cout;
    p.n = n; p.d = d; p.h = h; p.w = w;
    p.tiles_h = h / TC_TH; p.tiles_w = w / TC_TW;
    // d-chunk: enough work items to balance 148 SMs x resident CTAs, but at least 8 planes per chunk
    const int tiles = n * p.tiles_h * p.tiles_w;
    int dc = d;
    while (dc > 8 && (long long)tiles * ((d + dc - 1) / dc) < 148 * 6) dc = (dc + 1) / 2;
    p.dc = dc;
    p.dchunks = (d + dc - 1) / dc;
    p.total_items = tiles * p.dchunks;
    p.wimg_bytes = g.wimg_bytes; p.plane_bytes = g.plane_bytes; p.slot_bytes = g.slot_bytes; p.tmem_cols = g.tmem_cols;
    if (fuse_stats) {
        cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(double) * 2 * g.cob_n * 8, stream);
        if (e != cudaSuccess) {
            set_error("conv3d tensor path: memset: %s", cudaGetErrorString(e));
            return (int)e;
        }
    }
    int ctas_per_sm = (int)((227 * 1024) / (g.smem + 1024));
    if (ctas_per_sm > 512 / (int)g.tmem_cols) ctas_per_sm = 512 / (int)g.tmem_cols;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm > 3) ctas_per_sm = 3;
    int grid = 148 * ctas_per_sm;
    if (grid > p.total_items) grid = p.total_items;
    auto go = [&](auto kern) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
        if (e != cudaSuccess) {
            set_error("conv3d tensor path: smem %zu: %s", g.smem, cudaGetErrorString(e));
            return (int)e;
        }
        kern<<<grid, TC_THREADS, g.smem, stream>>>(tmap, p);
        return check_launch("ctu_conv3d_fprop(tcgen05)");
    };
    int rc = k == 3 ? go(conv3d_tc_kernel<3>) : go(conv3d_tc_kernel<5>);
    if (rc == CTU_OK && stats != nullptr && !fuse_stats)   // wide layers (low resolution): separate statistics pass
        rc = ctu_bn_stats(CTU_BF16, y, cout, n, (long long)d * h * w, stats, stream);
    return rc;
}

int conv3d_wgrad_tc(const void* const*, const int*, int, const void*, float*, float*, int, int, int, int, int, int,
                    cudaStream_t) {
    set_error("wgrad tensor path not built");
    return CTU_ERR_UNSUPPORTED;
}

}  // namespace ctu

using namespace ctu;

extern "C" {

int ctu_has_tensor_path(void) { return 1; }

int ctu_conv_tc_supported(int k, int cin, int cout, int d, int h, int w) {
    (void)d;
    TcGeom g;
    return tc_geometry(k, cin, cout, h, w, g) ? 1 : 0;
}

long long ctu_conv_tc_wimg_bytes(int k, int cin, int cout) {
    TcGeom g;
    if (!tc_geometry(k, cin, cout, TC_TH, TC_TW, g)) return -1;
    return (long long)g.wimg_bytes;
}

int ctu_conv_tc_pack_weight(const float* wp, void* wimg, int k, int cin, int cout, ctu_stream stream) {
    TcGeom g;
    CTU_REQUIRE(wp && wimg && tc_geometry(k, cin, cout, TC_TH, TC_TW, g), "ctu_conv_tc_pack_weight: bad arguments");
    const long long total = (long long)g.wimg_bytes / 2;
    tc_pack_wimg_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(wp, reinterpret_cast<__nv_bfloat16*>(wimg), k, g.cb, g.cob_n, g.npad, total);
    return check_launch("ctu_conv_tc_pack_weight");
}

}  // extern "C"
