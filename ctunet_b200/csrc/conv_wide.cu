// tcgen05 / TMEM / TMA implicit-GEMM 3-D convolution for WIDE, LOW-RESOLUTION layers (sm_100a, bf16 in, fp32 accumulate).
//
// Forward and data gradient of nn.Conv3d k in {3,5}, stride 1, "same" padding, for the layers conv_tc.cu does not
// cover or covers badly: many channels on a small grid -- the 16^3 and 8^3 levels of the networks
// (ctunet/pytorch/models.py:71,76 center block; :403,407,430,434,483,487 with 56..128 channels of the legacy 5^3
// family).  conv_tc.cu keeps ALL taps' weights in shared memory and carries the kd taps in the MMA N dimension, which
// only works while Cin*Cout*k^2 is small; here the weights are STREAMED tap by tap instead:
//
//   * GEMM per (tap, pair of input channel blocks): M = one 16(h) x 8(w) tile (M = 128; two tiles per 16x16 plane
//     tile) or one 8 x 8 plane (M = 64, the 8^3 level), N = NC output channels (16..128), K = 16 input channels.
//     The A operand is the TMA-staged halo plane addressed by descriptor arithmetic exactly as in conv_tc.cu
//     (8 w-voxels x 8 channels = one K-major core matrix; SBO = halo row pitch; LBO = the next channel block).
//   * A CTA owns (one output d-plane tile, NC output channels): its TMEM accumulators live across the whole
//     reduction (channel group, kd, kh, kw); input planes (<= 8 channel blocks per ring slot) and per-tap weight
//     tiles (pairs x NC x 32 B, contiguous in the pre-packed image, one cp.async.bulk each) arrive through two rings.
//     kd taps that fall outside the volume are skipped by producer and issuer alike.
//   * Warp roles: warp 0 producer, one MMA issuer warp per tile, four epilogue warps per tile
//     (tcgen05.ld -> bias -> bf16 -> 16-byte stores).  BatchNorm statistics of these (small) outputs are taken by
//     ctu_bn_stats afterwards.
//   * An odd channel-block count is closed with a dummy K chunk: zero weights against a re-read of the same block.
#include "tc_ptx.cuh"

namespace ctu {

constexpr int WD_MAX_CGB = 8;     // input channel blocks per plane-ring slot
constexpr int WD_NISS_SMALL = 4;  // issuer warps of the single M = 64 tile (8 x 8 planes)

struct WdParams {
    const __nv_bfloat16* wimg;   // [ngroup][pair-unit order of use][NC/8][2][8][8] bf16 (wd_pack_wimg_kernel)
    const float* bias;
    __nv_bfloat16* y;
    int cb, cgb, ncg;            // input blocks, blocks per group (even unless ncg == 1), groups
    int cob_n, cout, nc;         // output blocks, channels, channels per CTA (MMA N)
    int n, d, h, w, tiles_h, tiles_w, total_items;
    uint32_t plane_bytes, pslot_bytes, wslot_bytes, tmem_cols, np, nw;
    unsigned long long group_bytes;   // weight image of one output-channel group
};

// pairs of channel blocks (MMAs per tap) of group cg
__host__ __device__ inline int wd_group_blocks(int cb, int cgb, int cg) { return (cb - cg * cgb) < cgb ? (cb - cg * cgb) : cgb; }

// fp32 packed weights [cob][cib][tap][ci][co] -> per output group g the B tiles in order of use:
// MMA index m = cgoff(cg)*K^3 + (kd*K^2 + tap2d)*pairs(cg) + pr ; tile = [NC/8 n-groups][2 K chunks][8 n][8 k]
__global__ void wd_pack_wimg_kernel(const float* __restrict__ wp, __nv_bfloat16* __restrict__ wimg, int k, int cb, int cgb,
                                    int ncg, int cob_n, int nc, long long per_group, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int g = (int)(i / per_group);
    long long q = i % per_group;
    const int e = (int)(q & 7), r = (int)((q >> 3) & 7), c = (int)((q >> 6) & 1);
    q >>= 7;
    const int j = (int)(q % (nc / 8));
    const long long m = q / (nc / 8);
    const int k2 = k * k, k3 = k2 * k, pf = (cgb + 1) / 2;
    int cg = (int)(m / ((long long)pf * k3));
    if (cg > ncg - 1) cg = ncg - 1;
    const int ml = (int)(m - (long long)cg * pf * k3);
    const int gb = wd_group_blocks(cb, cgb, cg), pairs = (gb + 1) / 2;
    const int kdtap = ml / pairs, pr = ml % pairs;
    const int blk = cg * cgb + 2 * pr + c;
    const int cob = (g * nc) / 8 + j;
    float v = 0.f;
    if (2 * pr + c < gb && cob < cob_n && kdtap < k3)
        v = wp[(((long long)cob * cb + blk) * k3 + kdtap) * 64 + e * 8 + r];   // e = input lane (K), r = output lane (N)
    wimg[i] = __float2bfloat16_rn(v);
}

// NISS issuer warps per tile: the taps are dealt round-robin, every issuer accumulates into its own TMEM columns and the
// epilogue adds them up (one issuing thread sustains one MMA per ~76 cycles; an M = 64 MMA is worth ~24).
template <int K, int TH, int NTILE, int NISS>
__global__ void __launch_bounds__(32 * (1 + NTILE * NISS + 4 * NTILE)) conv3d_wide_kernel(const __grid_constant__ CUtensorMap xmap,
                                                                                          WdParams p) {
    constexpr int PAD = K / 2, K2 = K * K, K3 = K2 * K;
    constexpr int TWID = 8 * NTILE;
    constexpr int HH = TH + K - 1, WW = TWID + K - 1;
    constexpr uint32_t ROW = WW * 16;
    constexpr int M = TH * 8;                       // 128 or 64
    static_assert(M == 128 || M == 64, "tile");
    const uint32_t NP = p.np, NW = p.nw;
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_planes = s_base;
    const uint32_t s_w = s_planes + NP * p.pslot_bytes;
    const uint32_t s_bar = s_w + NW * p.wslot_bytes;
    const uint32_t b_pfull = s_bar, b_pempty = s_bar + 8 * TC_MAX_SLOTS, b_wfull = s_bar + 16 * TC_MAX_SLOTS,
                   b_wempty = s_bar + 24 * TC_MAX_SLOTS, b_afull = s_bar + 32 * TC_MAX_SLOTS, b_aempty = b_afull + 8 * NTILE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (b_aempty + 8 * NTILE - s_base));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < NP; ++i) {
            mbar_init(b_pfull + 8 * i, 1);
            mbar_init(b_pempty + 8 * i, NTILE * NISS);   // every issuer warp releases the plane
        }
        for (uint32_t i = 0; i < NW; ++i) {
            mbar_init(b_wfull + 8 * i, 1);
            mbar_init(b_wempty + 8 * i, NTILE);          // a tap is consumed by one issuer per tile
        }
        for (int i = 0; i < NTILE; ++i) {
            mbar_init(b_afull + 8 * i, NISS);
            mbar_init(b_aempty + 8 * i, 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int g = blockIdx.y;                        // output-channel group
    const int tiles_per_plane = p.tiles_h * p.tiles_w;
    const uint32_t mma_bytes = (uint32_t)p.nc * 32u; // one B tile
    const int pf = (p.cgb + 1) / 2;                  // pairs of a full group

    if (warp == 0) {
        // ===================================================================== producer
        if (lane == 0) {
            const unsigned char* wg = reinterpret_cast<const unsigned char*>(p.wimg) + (size_t)g * p.group_bytes;
            Ring pr = {0, 0}, wr = {0, 0};
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                int r = item;
                const int twi = r % p.tiles_w; r /= p.tiles_w;
                const int thi = r % p.tiles_h; r /= p.tiles_h;
                const int z = r % p.d, n = r / p.d;
                const int h0 = thi * TH, w0 = twi * TWID;
                for (int cg = 0; cg < p.ncg; ++cg) {
                    const int gb = wd_group_blocks(p.cb, p.cgb, cg), pairs = (gb + 1) / 2;
                    for (int kd = 0; kd < K; ++kd) {
                        const int zi = z + kd - PAD;
                        if (zi < 0 || zi >= p.d) continue;
                        mbar_wait(b_pempty + 8 * pr.slot, pr.phase ^ 1);
                        mbar_expect_tx(b_pfull + 8 * pr.slot, (uint32_t)gb * HH * WW * 16);
                        for (int b = 0; b < gb; ++b)
                            tma_load_4d(s_planes + pr.slot * p.pslot_bytes + b * p.plane_bytes, &xmap, (w0 - PAD) * 8, h0 - PAD,
                                        zi, n * p.cb + cg * p.cgb + b, b_pfull + 8 * pr.slot);
                        pr.next(NP);
                        const unsigned char* wsrc = wg + ((size_t)cg * pf * K3 + (size_t)kd * K2 * pairs) * mma_bytes;
                        const uint32_t tap_bytes = (uint32_t)pairs * mma_bytes;
                        for (int tap = 0; tap < K2; ++tap, wr.next(NW)) {
                            mbar_wait(b_wempty + 8 * wr.slot, wr.phase ^ 1);
                            mbar_expect_tx(b_wfull + 8 * wr.slot, tap_bytes);
                            bulk_load_1d(s_w + wr.slot * p.wslot_bytes, wsrc + (size_t)tap * tap_bytes, tap_bytes,
                                         b_wfull + 8 * wr.slot);
                        }
                    }
                }
            }
        }
    } else if (warp <= NTILE * NISS) {
        // ===================================================================== MMA issuer q of tile t
        const uint32_t leader = elect_one();
        const uint32_t t = (warp - 1) / NISS, q_iss = (warp - 1) % NISS;
        // D=f32, A=B=bf16, both K-major, N at [17,23), M at [24,29)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.nc >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t a_hi = (ROW >> 4) | (1u << 14);            // SBO = halo row pitch (next h)
        const uint32_t b_hi = 16u | (1u << 14);                   // SBO = 256 B between n-groups
        const uint32_t plane16 = p.plane_bytes >> 4;
        const uint32_t bstep = mma_bytes >> 4;
        const uint32_t d_tmem = tmem_base + (t * NISS + q_iss) * (uint32_t)p.nc;
        Ring pc = {0, 0}, wc = {0, 0};
        uint32_t it = 0;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
            const int z = (item / tiles_per_plane) % p.d;
            mbar_wait(b_aempty + 8 * t, (it & 1) ^ 1);
            tc_fence_after();
            uint32_t acc = 0;
            for (int cg = 0; cg < p.ncg; ++cg) {
                const int gb = wd_group_blocks(p.cb, p.cgb, cg), pairs = (gb + 1) / 2;
                for (int kd = 0; kd < K; ++kd) {
                    const int zi = z + kd - PAD;
                    if (zi < 0 || zi >= p.d) continue;
                    mbar_wait(b_pfull + 8 * pc.slot, pc.phase);
                    const uint32_t a16 = (s_planes + pc.slot * p.pslot_bytes + t * 128u) >> 4;
#pragma unroll 1
                    for (int tap = 0; tap < K2; ++tap, wc.next(NW)) {
                        if (NISS > 1 && (uint32_t)(tap % NISS) != q_iss) continue;   // (the for-increment still advances the ring)
                        mbar_wait(b_wfull + 8 * wc.slot, wc.phase);
                        tc_fence_after();
                        uint32_t b_lo = ((s_w + wc.slot * p.wslot_bytes) >> 4) | (8u << 16);   // LBO = 128 B (second K chunk)
                        const uint32_t a_tap = a16 + (uint32_t)(tap / K) * (ROW >> 4) + (uint32_t)(tap % K);
                        for (int q = 0; q < pairs; ++q) {
                            const uint32_t lbo = (2 * q + 1 < gb) ? plane16 : 0u;   // dummy half: same block x zero weights
                            umma_bf16_lead(leader, d_tmem, (a_tap + 2u * q * plane16) | (lbo << 16), a_hi, b_lo, b_hi, idesc, acc);
                            acc = 1;
                            b_lo += bstep;
                        }
                        umma_commit_lead(leader, b_wempty + 8 * wc.slot);
                    }
                    umma_commit_lead(leader, b_pempty + 8 * pc.slot);
                    pc.next(NP);
                }
            }
            umma_commit_lead(leader, b_afull + 8 * t);
        }
    } else {
        // ===================================================================== epilogue: 4 warps per tile
        const int te = (warp - 1 - NTILE * NISS) >> 2;
        const int quarter = warp & 3;
        const int row = (M == 128) ? quarter * 32 + lane : quarter * 16 + (lane & 15);
        const bool active = (M == 128) || lane < 16;
        const int hh = row >> 3, wl = row & 7;
        const long long plane = (long long)p.d * p.h * p.w;
        const int ob0 = (g * p.nc) / 8, nobc = p.nc / 8;
        const bool has_bias = p.bias != nullptr;
        uint32_t it = 0;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
            int r = item;
            const int twi = r % p.tiles_w; r /= p.tiles_w;
            const int thi = r % p.tiles_h; r /= p.tiles_h;
            const int z = r % p.d, n = r / p.d;
            const int gy = thi * TH + hh, gx = twi * TWID + te * 8 + wl;
            __nv_bfloat16* ycol = p.y + (((long long)n * p.cob_n + ob0) * plane + ((long long)z * p.h + gy) * p.w + gx) * 8;
            mbar_wait(b_afull + 8 * te, it & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(te * NISS) * p.nc;
            for (int ob = 0; ob < nobc; ob += 2) {
                uint32_t raw[NISS][2][8];
#pragma unroll
                for (int qi = 0; qi < NISS; ++qi) {
                    tmem_ld8_issue(taddr + qi * p.nc + ob * 8, raw[qi][0]);
                    tmem_ld8_issue(taddr + qi * p.nc + ob * 8 + 8, raw[qi][1]);      // nc is a multiple of 16
                }
                tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < 2; ++u) {
#pragma unroll
                    for (int qi = 0; qi < NISS; ++qi) tmem_ld8_pin(raw[qi][u]);
                    if (!active || ob0 + ob + u >= p.cob_n) continue;
                    V8 o;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        float v = __uint_as_float(raw[0][u][c]);
#pragma unroll
                        for (int qi = 1; qi < NISS; ++qi) v += __uint_as_float(raw[qi][u][c]);
                        const int ch = (ob0 + ob + u) * 8 + c;
                        if (has_bias && ch < p.cout) v += __ldg(p.bias + ch);
                        o.v[c] = v;
                    }
                    Vec8<__nv_bfloat16>::store(ycol + (long long)(ob + u) * plane * 8, o);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_aempty + 8 * te);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host
struct WdGeom {
    int th, ntile;               // 16 x 16 plane tiles (two M=128 MMAs) or 8 x 8 (one M=64 MMA)
    int cb, cgb, ncg, cob_n, nc, ngroups, total_pairs;
    uint32_t plane_bytes, pslot_bytes, wslot_bytes, tmem_cols, np, nw;
    unsigned long long group_bytes;
    size_t smem;
};

static bool wide_geometry(int k, int cb, int cout, int n, int d, int h, int w, WdGeom& g) {
    if ((k != 3 && k != 5) || cb < 1 || cout < 1 || n < 1 || d < 1) return false;
    if (h % 16 == 0 && w % 16 == 0) {
        g.th = 16; g.ntile = 2;
    } else if (h % 8 == 0 && w % 8 == 0) {
        g.th = 8; g.ntile = 1;
    } else {
        return false;
    }
    const int twid = 8 * g.ntile, m = g.th * 8;
    g.cb = cb;
    g.cgb = cb < WD_MAX_CGB ? cb : WD_MAX_CGB;
    g.ncg = (cb + g.cgb - 1) / g.cgb;
    g.cob_n = (cout + 7) / 8;
    g.total_pairs = 0;
    for (int cg = 0; cg < g.ncg; ++cg) g.total_pairs += (wd_group_blocks(cb, g.cgb, cg) + 1) / 2;
    const long long items = (long long)n * d * (h / g.th) * (w / twid);
    // output channels per CTA: the N that minimises waves x cycles per MMA (SS-mode UMMA is bound by the operand
    // reads from shared memory, ~128 B/clk: A = M x 32 B, B = N x 32 B)
    const int coutp = (cout + 15) / 16 * 16;
    double best = 1e30;
    g.nc = 16;
    for (int nc = 16; nc <= 128; nc *= 2) {
        const int ng = (coutp + nc - 1) / nc;
        long long gx = 148 / ng > 0 ? 148 / ng : 1;
        if (gx > items) gx = items;
        const double per_cta = (double)((items + gx - 1) / gx);
        const double cost = per_cta * (m * 32.0 + nc * 32.0);
        if (cost < best - 1e-9) {
            best = cost;
            g.nc = nc;
        }
        if (nc >= coutp) break;
    }
    g.ngroups = (coutp + g.nc - 1) / g.nc;
    g.plane_bytes = ((uint32_t)(g.th + k - 1) * (twid + k - 1) * 16 + 127u) & ~127u;
    g.pslot_bytes = g.plane_bytes * g.cgb;
    g.wslot_bytes = (uint32_t)((g.cgb + 1) / 2) * g.nc * 32;
    g.group_bytes = (unsigned long long)g.total_pairs * k * k * k * g.nc * 32;
    const size_t fixed = 8 * (4 * TC_MAX_SLOTS + 2 * g.ntile) + 64 + 1024;
    // the weight stream is latency bound (one tap tile per bulk copy): as many tiles in flight as the ring allows
    g.nw = TC_MAX_SLOTS;
    g.np = 2;
    while (g.nw > 4 && fixed + (size_t)g.np * g.pslot_bytes + (size_t)g.nw * g.wslot_bytes > 216 * 1024) --g.nw;
    while (g.np < 4 && fixed + (size_t)(g.np + 1) * g.pslot_bytes + (size_t)g.nw * g.wslot_bytes <= 216 * 1024) ++g.np;
    g.smem = fixed + (size_t)g.np * g.pslot_bytes + (size_t)g.nw * g.wslot_bytes;
    if (g.smem > 220 * 1024) return false;
    uint32_t cols = (uint32_t)g.ntile * (g.ntile == 1 ? WD_NISS_SMALL : 1) * g.nc;
    g.tmem_cols = 32;
    while (g.tmem_cols < cols) g.tmem_cols *= 2;
    return g.tmem_cols <= 512;
}

bool conv3d_wide_supported(int k, int cb, int cout, int n, int d, int h, int w) {
    WdGeom g;
    return wide_geometry(k, cb, cout, n, d, h, w, g);
}

long long conv3d_wide_wimg_bytes(int k, int cb, int cout, int n, int d, int h, int w) {
    WdGeom g;
    if (!wide_geometry(k, cb, cout, n, d, h, w, g)) return -1;
    return (long long)(g.group_bytes * g.ngroups);
}

int conv3d_wide_pack_weight(const float* wp, void* wimg, int k, int cb, int cout, int n, int d, int h, int w,
                            cudaStream_t stream) {
    WdGeom g;
    if (!wide_geometry(k, cb, cout, n, d, h, w, g)) {
        set_error("conv3d wide path: shape k=%d blocks=%d cout=%d %dx%dx%d not covered", k, cb, cout, d, h, w);
        return CTU_ERR_UNSUPPORTED;
    }
    const long long per_group = (long long)(g.group_bytes / 2), total = per_group * g.ngroups;
    wd_pack_wimg_kernel<<<cdiv(total, 256), 256, 0, stream>>>(wp, reinterpret_cast<__nv_bfloat16*>(wimg), k, g.cb, g.cgb, g.ncg,
                                                              g.cob_n, g.nc, per_group, total);
    return check_launch("ctu_conv_tc_pack_weight(wide)");
}

int conv3d_fprop_wide(const void* x, int cin, const void* wimg, const float* bias, void* y, int cout, int k, int n, int d,
                      int h, int w, cudaStream_t stream) {
    WdGeom g;
    const int cb = (cin + 7) / 8;
    if (!wide_geometry(k, cb, cout, n, d, h, w, g)) {
        set_error("conv3d wide path: shape k=%d blocks=%d cout=%d %dx%dx%d not covered", k, cb, cout, d, h, w);
        return CTU_ERR_UNSUPPORTED;
    }
    CUtensorMap xmap;
    int rc = make_map(&xmap, x, n * cb, d, h, w, 8 * g.ntile + k - 1, g.th + k - 1);
    if (rc != CTU_OK) return rc;
    WdParams p = {};
    p.wimg = reinterpret_cast<const __nv_bfloat16*>(wimg);
    p.bias = bias;
    p.y = reinterpret_cast<__nv_bfloat16*>(y);
    p.cb = g.cb; p.cgb = g.cgb; p.ncg = g.ncg; p.cob_n = g.cob_n; p.cout = cout; p.nc = g.nc;
    p.n = n; p.d = d; p.h = h; p.w = w;
    p.tiles_h = h / g.th; p.tiles_w = w / (8 * g.ntile);
    p.total_items = n * d * p.tiles_h * p.tiles_w;
    p.plane_bytes = g.plane_bytes; p.pslot_bytes = g.pslot_bytes; p.wslot_bytes = g.wslot_bytes;
    p.tmem_cols = g.tmem_cols; p.np = g.np; p.nw = g.nw; p.group_bytes = g.group_bytes;
    int gx = 148 / g.ngroups;
    if (gx < 1) gx = 1;
    if (gx > p.total_items) gx = p.total_items;
    auto go = [&](auto kern, int threads) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
        if (e != cudaSuccess) {
            set_error("conv3d wide path: smem %zu: %s", g.smem, cudaGetErrorString(e));
            return (int)e;
        }
        kern<<<dim3(gx, g.ngroups), threads, g.smem, stream>>>(xmap, p);
        return check_launch("ctu_conv3d_fprop(tcgen05 wide)");
    };
    if (k == 3 && g.ntile == 2) return go(conv3d_wide_kernel<3, 16, 2, 1>, 32 * 11);
    if (k == 3) return go(conv3d_wide_kernel<3, 8, 1, WD_NISS_SMALL>, 32 * (1 + WD_NISS_SMALL + 4));
    if (g.ntile == 2) return go(conv3d_wide_kernel<5, 16, 2, 1>, 32 * 11);
    return go(conv3d_wide_kernel<5, 8, 1, WD_NISS_SMALL>, 32 * (1 + WD_NISS_SMALL + 4));
}


// ================================================================================================ wgrad, small grids
// dW[co][ci][kd][kh][kw] = sum_{n,z,y,x} dy[n][co][z][y][x] * x[n][ci][z+kd-P][y+kh-P][x+kw-P] on 8 x 8 plane tiles
// (the 8^3 level: conv3d_wgrad_tc_kernel needs 16-wide rows).  TAP-STATIONARY: a CTA owns up to 512/N taps of one kd
// (one TMEM accumulator [M = input channels] x [N = output channels] per tap) and one sample, and walks the d-planes:
//   A = x halo plane, MN-major: M-groups = channel blocks (SBO = plane pitch), K = 16 voxels = two 8-voxel rows
//       (LBO = halo row pitch); start address = the tap's (kh, kw) shift;
//   B = dy plane [cob][8][8][8], MN-major: N-groups = channel blocks (SBO = 1 KB), K = the same two rows (LBO = 128 B).
// Four MMAs (row pairs) per (plane, tap).  Partial sums are flushed with fp32 atomics (dwp zeroed by the caller).
struct WsParams {
    float* dwp;
    int cb, cob_n, mblk, nblk;       // real blocks; blocks spanned by the MMA (M / 8, N / 8)
    int tpc, chunks;                 // taps per CTA, tap chunks per kd
    int zsplit;                      // CTAs sharing the d-planes of one (kd, chunk, sample)
    int n, d, h, w, tiles_h, tiles_w;
    uint32_t plane_bytes, xslot_bytes, dyslot_bytes, tmem_cols, ns;
};

template <int K>
__global__ void __launch_bounds__(32 * 5) conv3d_wgrad_small_kernel(const __grid_constant__ CUtensorMap xmap,
                                                                    const __grid_constant__ CUtensorMap dymap, WsParams p) {
    constexpr int PAD = K / 2, K2 = K * K, K3 = K2 * K;
    constexpr int WW = 8 + K - 1, HH = 8 + K - 1;
    constexpr uint32_t ROW = WW * 16;
    const uint32_t NS = p.ns;
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_x = s_base, s_dy = s_x + NS * p.xslot_bytes, s_bar = s_dy + NS * p.dyslot_bytes;
    const uint32_t b_full = s_bar, b_empty = s_bar + 8 * TC_MAX_SLOTS, b_done = s_bar + 16 * TC_MAX_SLOTS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (b_done + 8 - s_base));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kd = blockIdx.x / p.chunks, chunk = blockIdx.x % p.chunks;
    const int tap0 = chunk * p.tpc;                                   // first (kh, kw) tap of this CTA
    const int ntap = (K2 - tap0) < p.tpc ? (K2 - tap0) : p.tpc;
    const int n = blockIdx.y;
    const int ncol = p.nblk * 8, M = p.mblk * 8;
    const int zb = (int)(((long long)p.d * blockIdx.z) / p.zsplit), ze = (int)(((long long)p.d * (blockIdx.z + 1)) / p.zsplit);

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < NS; ++i) {
            mbar_init(b_full + 8 * i, 1);
            mbar_init(b_empty + 8 * i, 4);          // every issuer warp releases the slot
        }
        mbar_init(b_done, 4);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles = p.tiles_h * p.tiles_w;

    if (warp == 0) {
        if (lane == 0) {
            Ring pr = {0, 0};
            for (int z = zb; z < ze; ++z) {
                const int zi = z + kd - PAD;
                if (zi < 0 || zi >= p.d) continue;
                for (int t = 0; t < tiles; ++t, pr.next(NS)) {
                    const int h0 = (t / p.tiles_w) * 8, w0 = (t % p.tiles_w) * 8;
                    mbar_wait(b_empty + 8 * pr.slot, pr.phase ^ 1);
                    mbar_expect_tx(b_full + 8 * pr.slot, (uint32_t)p.mblk * HH * WW * 16 + (uint32_t)p.nblk * 1024u);
                    // blocks past the real channel count read the neighbouring sample (or zero fill past the tensor):
                    // they only feed accumulator rows / columns that are never flushed
                    for (int b = 0; b < p.mblk; ++b)
                        tma_load_4d(s_x + pr.slot * p.xslot_bytes + b * p.plane_bytes, &xmap, (w0 - PAD) * 8, h0 - PAD, zi,
                                    n * p.cb + b, b_full + 8 * pr.slot);
                    for (int b = 0; b < p.nblk; ++b)
                        tma_load_4d(s_dy + pr.slot * p.dyslot_bytes + b * 1024u, &dymap, w0 * 8, h0, z, n * p.cob_n + b,
                                    b_full + 8 * pr.slot);
                }
            }
        }
    } else {
        // ===================================================================== 4 issuer warps (taps dealt round-robin: one
        // issuing thread sustains one MMA per ~76 cycles, four reach the shared-memory operand limit), then the flush
        const int q4 = warp - 1;
        const uint32_t leader = elect_one();
        // D=f32, A=B=bf16, both MN-major (bits 15, 16), N at [17,23), M at [24,29)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(ncol >> 3) << 17) |
                               ((uint32_t)(M >> 4) << 24);
        const uint32_t a_hi = (p.plane_bytes >> 4) | (1u << 14);      // SBO: next channel block (M group)
        const uint32_t b_hi = (1024u >> 4) | (1u << 14);              // SBO: next output channel block (N group)
        const uint32_t a_lbo = (ROW >> 4) << 16, b_lbo = 8u << 16;    // LBO: the second row of the K = 16 voxels
        Ring cons = {0, 0};
        uint32_t first = 1;
        for (int z = zb; z < ze; ++z) {
            const int zi = z + kd - PAD;
            if (zi < 0 || zi >= p.d) continue;
            for (int t = 0; t < tiles; ++t, cons.next(NS)) {
                mbar_wait(b_full + 8 * cons.slot, cons.phase);
                tc_fence_after();
                const uint32_t x16 = (s_x + cons.slot * p.xslot_bytes) >> 4, dy16 = (s_dy + cons.slot * p.dyslot_bytes) >> 4;
                for (int tp = q4; tp < ntap; tp += 4) {
                    const int kh = (tap0 + tp) / K, kw = (tap0 + tp) % K;
                    const uint32_t a_tap = x16 + (uint32_t)kh * (ROW >> 4) + (uint32_t)kw;
                    const uint32_t d_tmem = tmem_base + (uint32_t)tp * ncol;
#pragma unroll
                    for (int j = 0; j < 4; ++j)      // output rows 2j, 2j+1
                        umma_bf16_lead(leader, d_tmem, (a_tap + 2u * j * (ROW >> 4)) | a_lbo, a_hi, (dy16 + 16u * j) | b_lbo, b_hi,
                                       idesc, (j == 0) ? (first ^ 1u) : 1u);
                }
                first = 0;
                umma_commit_lead(leader, b_empty + 8 * cons.slot);
            }
        }
        umma_commit_lead(leader, b_done);
        __syncwarp();
        // ===================================================================== flush (4 warps = 4 TMEM lane quarters)
        const int quarter = warp & 3;
        mbar_wait(b_done, 0);
        tc_fence_after();
        const int row = (M == 128) ? quarter * 32 + lane : quarter * 16 + (lane & 15);   // input channel
        const bool useful = ((M == 128) || lane < 16) && row < p.cb * 8;
        bool any = false;                                   // a CTA whose kd never meets the volume has nothing to flush
        for (int z = zb; z < ze; ++z) any = any || (z + kd - PAD >= 0 && z + kd - PAD < p.d);
        if (any) {
            for (int tp = 0; tp < ntap; ++tp) {
                const int tap = kd * K2 + tap0 + tp;
                for (int ob = 0; ob < p.cob_n; ++ob) {
                    float v[8];
                    tmem_ld8(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)tp * ncol + ob * 8, v);
                    if (useful) {
                        float* dst = p.dwp + ((((long long)ob * p.cb + (row >> 3)) * K3 + tap) * 64 + (row & 7) * 8);
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            if (v[c] != 0.f) atomicAdd(dst + c, v[c]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

struct WsGeom {
    int cb, cob_n, mblk, nblk, tpc, chunks;
    uint32_t plane_bytes, xslot_bytes, dyslot_bytes, tmem_cols, ns;
    size_t smem;
};

static bool wgrad_small_geometry(int k, int cin, int cout, int h, int w, WsGeom& g) {
    if ((k != 3 && k != 5) || cin < 1 || cout < 1 || h % 8 || w % 8) return false;
    g.cb = (cin + 7) / 8;
    g.cob_n = (cout + 7) / 8;
    if (g.cb > 16 || g.cob_n > 16) return false;
    g.mblk = g.cb <= 8 ? 8 : 16;                       // M = 64 or 128
    g.nblk = (g.cob_n + 1) / 2 * 2;                    // N a multiple of 16
    const int ncol = g.nblk * 8;
    g.tpc = 512 / ncol;
    if (g.tpc > 8) g.tpc = 8;
    g.chunks = (k * k + g.tpc - 1) / g.tpc;
    g.tpc = (k * k + g.chunks - 1) / g.chunks;         // even out the chunks
    g.plane_bytes = ((uint32_t)(8 + k - 1) * (8 + k - 1) * 16 + 127u) & ~127u;
    g.xslot_bytes = g.plane_bytes * g.mblk;
    g.dyslot_bytes = 1024u * g.nblk;
    const size_t fixed = 8 * (2 * TC_MAX_SLOTS + 1) + 64 + 1024;
    g.ns = 4;
    while (g.ns > 2 && fixed + (size_t)g.ns * (g.xslot_bytes + g.dyslot_bytes) > 200 * 1024) --g.ns;
    g.smem = fixed + (size_t)g.ns * (g.xslot_bytes + g.dyslot_bytes);
    if (g.smem > 220 * 1024) return false;
    g.tmem_cols = 32;
    while (g.tmem_cols < (uint32_t)(g.tpc * ncol)) g.tmem_cols *= 2;
    return g.tmem_cols <= 512;
}

bool conv3d_wgrad_small_supported(int k, int cin, int cout, int h, int w) {
    WsGeom g;
    return wgrad_small_geometry(k, cin, cout, h, w, g);
}

int conv3d_wgrad_small(const void* x, int cin, const void* dy, float* dwp, float* dbias, int cout, int k, int n, int d, int h,
                       int w, cudaStream_t stream) {
    WsGeom g;
    if (!wgrad_small_geometry(k, cin, cout, h, w, g)) {
        set_error("conv3d wgrad small-grid path: shape k=%d cin=%d cout=%d %dx%dx%d not covered", k, cin, cout, d, h, w);
        return CTU_ERR_UNSUPPORTED;
    }
    CUtensorMap xmap, dymap;
    int rc = make_map(&xmap, x, n * g.cb, d, h, w, 8 + k - 1, 8 + k - 1);
    if (rc == CTU_OK) rc = make_map(&dymap, dy, n * g.cob_n, d, h, w, 8, 8);
    if (rc != CTU_OK) return rc;
    WsParams p = {};
    p.dwp = dwp;
    p.cb = g.cb; p.cob_n = g.cob_n; p.mblk = g.mblk; p.nblk = g.nblk; p.tpc = g.tpc; p.chunks = g.chunks;
    p.n = n; p.d = d; p.h = h; p.w = w; p.tiles_h = h / 8; p.tiles_w = w / 8;
    p.plane_bytes = g.plane_bytes; p.xslot_bytes = g.xslot_bytes; p.dyslot_bytes = g.dyslot_bytes;
    p.tmem_cols = g.tmem_cols; p.ns = g.ns;
    // split the d-planes until about one CTA per SM (every CTA flushes its taps with atomics)
    p.zsplit = 148 / (k * g.chunks * n);
    if (p.zsplit > d) p.zsplit = d;
    if (p.zsplit < 1) p.zsplit = 1;
    auto go = [&](auto kern) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
        if (e != cudaSuccess) {
            set_error("conv3d wgrad small-grid path: smem %zu: %s", g.smem, cudaGetErrorString(e));
            return (int)e;
        }
        kern<<<dim3(k * g.chunks, n, p.zsplit), 32 * 5, g.smem, stream>>>(xmap, dymap, p);
        return check_launch("ctu_conv3d_wgrad(tcgen05 small grid)");
    };
    rc = k == 3 ? go(conv3d_wgrad_small_kernel<3>) : go(conv3d_wgrad_small_kernel<5>);
    if (rc == CTU_OK && dbias != nullptr) rc = channel_sum_bias(dy, dbias, cout, g.cob_n, n, (long long)d * h * w, stream);
    return rc;
}

}  // namespace ctu

using namespace ctu;

extern "C" {

int ctu_conv_wide_supported(int k, int cin, int cout, int n, int d, int h, int w) {
    return (cin > 0 && conv3d_wide_supported(k, (cin + 7) / 8, cout, n, d, h, w)) ? 1 : 0;
}

long long ctu_conv_wide_wimg_bytes(int k, int cin, int cout, int n, int d, int h, int w) {
    return cin > 0 ? conv3d_wide_wimg_bytes(k, (cin + 7) / 8, cout, n, d, h, w) : -1;
}

int ctu_conv_wide_pack_weight(const float* wp, void* wimg, int k, int cin, int cout, int n, int d, int h, int w,
                              ctu_stream stream) {
    CTU_REQUIRE(wp && wimg && cin > 0, "ctu_conv_wide_pack_weight: bad arguments");
    return conv3d_wide_pack_weight(wp, wimg, k, (cin + 7) / 8, cout, n, d, h, w, (cudaStream_t)stream);
}

int ctu_conv_wide_wgrad_supported(int k, int cin, int cout, int d, int h, int w) {
    (void)d;
    return conv3d_wgrad_small_supported(k, cin, cout, h, w) ? 1 : 0;
}

}  // extern "C"
