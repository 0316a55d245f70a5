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
#include "tc_ptx.cuh"

#include <stdlib.h>

namespace ctu {

// -DCTU_TCPROF: CTA 0 of every fprop/dgrad launch prints where its roles spent their cycles (debug builds only)
#ifdef CTU_TCPROF
#define TCPROF_T0() const long long _t0 = clock64()
#define TCPROF_ADD(acc) acc += clock64() - _t0
#else
#define TCPROF_T0()
#define TCPROF_ADD(acc)
#endif

constexpr int TC_TH = 16, TC_TW = 16;   // output tile (h, w) per plane = 2 MMA tiles of 16x8
constexpr int TC_WB = TC_TW / 8;
constexpr int TC_MAX_STAGES = 4;         // TMEM accumulator stages per 16x8 tile (fprop/dgrad)
// fprop/dgrad: warp 0 TMA producer, warps 1..2 MMA issuers (one per 16x8 tile), warps 3..10 epilogue (4 per tile)
// wgrad: warp 0 TMA producer, warps 1..4 MMA issuers (accumulators dealt round-robin) and final flush
constexpr int WG_ISSUERS = 4;
constexpr int WG_THREADS = 32 * (1 + WG_ISSUERS);
constexpr int WG_MAX_OWN = 16;          // accumulators per issuer warp


// ------------------------------------------------------------------------------------------------ schedule
// The input channel blocks of one plane are staged in GROUPS of `cbg` blocks (one ring slot per group, so wide
// inputs -- the concatenated sources of the fused up-sampling stage -- never need more than cbg planes of shared
// memory per slot); cbg is even whenever there is more than one group.
// MMAs of one kd-plane, group by group: first every (tap, channel-block pair) -- second K chunk = the next channel
// block -- then, when the (last) group has an odd number of blocks, its last block's taps pairwise -- second K
// chunk = the next tap.
__host__ __device__ inline int tc_groups(int cb, int cbg) { return (cb + cbg - 1) / cbg; }
__host__ __device__ inline int tc_group_mmas(int k, int gb) {
    const int k2 = k * k;
    return k2 * (gb / 2) + ((gb & 1) ? (k2 + 1) / 2 : 0);
}
__host__ __device__ inline int tc_mmas_per_kd(int k, int cb, int cbg) {
    const int ng = tc_groups(cb, cbg);
    return (ng - 1) * tc_group_mmas(k, cbg) + tc_group_mmas(k, cb - (ng - 1) * cbg);
}
// first MMA of group g
__host__ __device__ inline int tc_group_start(int k, int cb, int cbg, int g) {
    const int ng = tc_groups(cb, cbg);
    return g < ng ? g * tc_group_mmas(k, cbg) : tc_mmas_per_kd(k, cb, cbg);
}
// chunk c (0/1) of MMA m: global block, 2-D tap and group; returns false for the zero-weight dummy half of an odd tail
__host__ __device__ inline bool tc_chunk(int k, int cb, int cbg, int m, int c, int& blk, int& tap2d, int& grp) {
    const int k2 = k * k, ng = tc_groups(cb, cbg), per = tc_group_mmas(k, cbg);
    grp = (ng > 1 && per > 0) ? m / per : 0;
    if (grp > ng - 1) grp = ng - 1;
    const int ml = m - grp * (ng > 1 ? per : 0);
    const int gb = (grp == ng - 1) ? cb - grp * cbg : cbg;
    const int pairs = gb / 2;
    if (ml < k2 * pairs) {
        tap2d = ml / pairs;
        blk = grp * cbg + 2 * (ml % pairs) + c;
        return true;
    }
    const int j = ml - k2 * pairs;
    blk = grp * cbg + gb - 1;
    tap2d = 2 * j + c;
    if (tap2d >= k2) {
        tap2d = k2 - 1;
        return false;
    }
    return true;
}

// up to CTU_MAX_SRC concatenated sources, one tensor map each
struct TcMaps {
    CUtensorMap m[CTU_MAX_SRC];
};

struct TcParams {
    const __nv_bfloat16* wimg;   // [mma m][NT/8][2][8][8] bf16: UMMA B tiles in order of use, n = kd*cpad + co
    const float* bias;
    __nv_bfloat16* y;
    double* stats;               // nullable: [2][cpad_out] sum, sum of squares (of the bf16-rounded outputs)
    int cb, cob_n, nt, cout;     // nt: MMA N = K*COB*8 rounded up to 16
    int cbg, ncg;                // channel blocks per ring slot, number of such groups
    int nsrc, src_cb[CTU_MAX_SRC], src_cboff[CTU_MAX_SRC];   // concatenated sources: blocks and first block of each
    int cobo, cstat;             // statistics: natural channel block = output block % cobo, cstat real channels
    int n, d, h, w;
    int tiles_h, tiles_w, dchunks, dc;
    int total_items;
    uint32_t wimg_bytes, plane_bytes, slot_bytes;   // plane_bytes: one channel block of one plane, padded to 128
    uint32_t tmem_cols;
    uint32_t ns;                                    // plane ring depth (more slots = more TMA loads in flight)
    uint32_t stmask, stshift;                       // TMEM accumulator stages per tile = stmask + 1 = 1 << stshift (2 or 4)
    int balance;                                    // 1: every CTA walks an equal share of the (column, z) plane sequence
    // BNRED kernels (data gradient feeding a BatchNorm+ReLU backward): y of that BatchNorm (natural or phase-major
    // layout), its ss = scale | shift | mean | invstd; `stats` then receives sum(dz) | sum(dz * xhat)
    const __nv_bfloat16* bn_y;
    const float* bn_ss;
    int bn_pm;
};

// fp32 packed weights [cob][cib][tap][ci][co] -> the bf16 B-tile images, one per output-block group of `cobg`
__global__ void tc_pack_wimg_kernel(const float* __restrict__ wp, __nv_bfloat16* __restrict__ wimg, int k, int cb, int cbg,
                                    int cob_n, int cobg, int nt, int nm, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int e = (int)(i & 7), r = (int)((i >> 3) & 7), c = (int)((i >> 6) & 1);
    long long q = i >> 7;
    const int ngroups = nt / 8;
    const int g = (int)(q % ngroups);
    q /= ngroups;
    const int m = (int)(q % nm);
    const int grp = (int)(q / nm);
    const int kd = g / cobg, cob = grp * cobg + g % cobg;
    int blk, tap2d, grp_;
    float v = 0.f;
    if (kd < k && cob < cob_n && tc_chunk(k, cb, cbg, m, c, blk, tap2d, grp_)) {
        const int taps = k * k * k, tap = kd * k * k + tap2d;
        v = wp[(((long long)cob * cb + blk) * taps + tap) * 64 + e * 8 + r];   // e = input lane (K), r = output lane (N)
    }
    wimg[i] = __float2bfloat16_rn(v);
}

// The work of one launch is the sequence of output planes (n, h-tile, w-tile, z), z fastest.  balance = 1: CTA b of G
// owns the contiguous share [total*b/G, total*(b+1)/G) of it -- every CTA computes the same number of planes (+-1), cut
// into segments at the column ends (a segment re-reads K - 1 halo planes, like a d-chunk); balance = 0: fixed d-chunks of
// dc planes dealt round-robin (the round count quantises: 1024 items on 296 CTAs are 4 rounds for 3.46 rounds of work).
struct TcWalk {
    long long p, pend;
    int item;
    template <class P>
    __device__ __forceinline__ void init(const P& q) {
        const long long total = (long long)q.n * q.tiles_h * q.tiles_w * q.d;
        p = total * blockIdx.x / gridDim.x;
        pend = total * (blockIdx.x + 1) / gridDim.x;
        item = blockIdx.x;
    }
    template <class P>
    __device__ __forceinline__ bool next(const P& q, int& n, int& thi, int& twi, int& z0, int& nd) {
        const int tiles = q.tiles_h * q.tiles_w;
        if (q.balance) {
            if (p >= pend) return false;
            const long long col = p / q.d;
            z0 = (int)(p - col * q.d);
            const long long rem = pend - p;
            nd = (long long)(q.d - z0) < rem ? (q.d - z0) : (int)rem;
            p += nd;
            n = (int)(col / tiles);
            const int r = (int)(col % tiles);
            thi = r / q.tiles_w;
            twi = r % q.tiles_w;
            return true;
        }
        if (item >= q.total_items) return false;
        const int items_per_n = tiles * q.dchunks;
        n = item / items_per_n;
        int r = item % items_per_n;
        const int dci = r % q.dchunks;
        r /= q.dchunks;
        twi = r % q.tiles_w;
        thi = r / q.tiles_w;
        z0 = dci * q.dc;
        nd = (q.d - z0) < q.dc ? (q.d - z0) : q.dc;
        item += gridDim.x;
        return true;
    }
};

// One CTA walks (n, 16x16 h-w tile, d-chunk) items plane by plane.  For every INPUT plane and every 16x8 tile a
// single accumulation group of tc_mmas_per_kd() MMAs computes P[voxel][kd][co] = sum_{kh,kw,ci} x * W, i.e. the
// three (five) kd taps ride in the MMA N dimension: the activation tile is read from shared memory once per
// plane instead of once per kd (SS-mode UMMA is bound by the A-operand read, ~64 B/clk, not by the math, when N
// is this small).  The epilogue thread of a voxel column adds P[kd] of K consecutive planes in registers.
// BNRED: the output IS dA of a BatchNorm+ReLU stage whose only consumer this convolution is (models.py:26-32): instead of
// the sums of the outputs the epilogue accumulates the two reductions of the BatchNorm backward pass,
// sum(dz) and sum(dz * xhat) with dz = dA * [scale * y + shift > 0], reading y (16 B per voxel and block) while the
// gradient tile is still in registers -- the separate reduce pass over y and dA (ctu_bn_relu_bwd_reduce) disappears.
template <int K, int COB, bool BNRED = false>
// (one-block variants: at most 88 registers, so that TWO CTAs -- four MMA issuers -- share an SM; for K = 5 that costs 64 bytes
//  of spills in the epilogue and buys 8 -> 8 at 4x128^3 200 -> 174 us with the balanced shares, recAE_v2_fixed 8.66 -> 8.2 ms)
__global__ void __launch_bounds__(32 * (1 + TC_WB + 4 * TC_WB), (COB == 1 && !BNRED) ? 2 : 1)
    conv3d_tc_kernel(const __grid_constant__ TcMaps maps, TcParams p) {
    constexpr int PAD = K / 2;
    constexpr int HH = TC_TH + K - 1, WW = TC_TW + K - 1;
    constexpr uint32_t ROW = WW * 16;              // bytes per halo row of one channel block
    constexpr int CP = COB * 8;                    // padded output channels
    constexpr int NSLOT = COB < 2 ? COB : 2;       // statistics accumulators (blocks) per thread
    const uint32_t NS = p.ns;
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_w = s_base;                                            // weights image
    const uint32_t s_planes = s_base + ((p.wimg_bytes + 1023u) & ~1023u);   // NS slots
    const uint32_t s_tab = s_planes + NS * p.slot_bytes;                    // per-MMA descriptor low words (A, B)
    const int nm = tc_mmas_per_kd(K, p.cb, p.cbg);
    const uint32_t s_bar = s_tab + ((nm * 8u + 15u) & ~15u);
    // barriers: plane_full[NS], plane_empty[NS], acc_full[<= 4 stages][TC_WB tiles], acc_empty[<= 4][TC_WB], w_full
    const uint32_t b_full = s_bar, b_empty = s_bar + 8 * TC_MAX_SLOTS, b_afull = s_bar + 16 * TC_MAX_SLOTS,
                   b_aempty = b_afull + 8 * TC_MAX_STAGES * TC_WB, b_w = b_aempty + 8 * TC_MAX_STAGES * TC_WB;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (b_w + 8 - s_base));
    float* st_red = reinterpret_cast<float*>(smem + (b_w + 32 - s_base));   // [4*TC_WB epilogue warps][NSLOT*16]
    float* st_bias = st_red + 4 * TC_WB * 32;                               // [CP] bias of this CTA's output blocks
    float* st_bn = st_bias + 32;                                            // BNRED: [4][CP] scale | shift | mean | invstd

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef CTU_TCPROF
    long long prof_a = 0, prof_b = 0, prof_c = 0, prof_d = 0;
    const long long prof_start = clock64();
#endif

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < NS; ++i) {
            mbar_init(b_full + 8 * i, 1);
            mbar_init(b_empty + 8 * i, TC_WB);      // every issuer warp releases the slot
        }
        for (int i = 0; i < TC_MAX_STAGES * TC_WB; ++i) {
            mbar_init(b_afull + 8 * i, 1);
            mbar_init(b_aempty + 8 * i, 4);
        }
        mbar_init(b_w, 1);
        fence_barrier_init();
    }
    if (warp == 1) {   // TMEM allocation (this warp also frees it)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x >= 64 && threadIdx.x < 64 + CP) {   // bias -> shared memory (read per plane by every epilogue thread)
        const int ch = blockIdx.y * CP + (int)threadIdx.x - 64;
        st_bias[threadIdx.x - 64] = (p.bias != nullptr && ch < p.cout) ? p.bias[ch] : 0.f;
        if (BNRED) {
            const int cpad = p.cobo * 8;
#pragma unroll
            for (int q = 0; q < 4; ++q) st_bn[q * CP + threadIdx.x - 64] = ch < cpad ? p.bn_ss[q * cpad + ch] : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            mbar_expect_tx(b_w, p.wimg_bytes);
            for (uint32_t off = 0; off < p.wimg_bytes; off += 32768u) {
                const uint32_t sz = p.wimg_bytes - off < 32768u ? p.wimg_bytes - off : 32768u;
                bulk_load_1d(s_w + off, reinterpret_cast<const unsigned char*>(p.wimg) + (size_t)blockIdx.y * p.wimg_bytes + off, sz, b_w);
            }
            Ring pr = {0, 0};
            TcWalk walk;
            walk.init(p);
            int n, thi, twi, z0, nd;
            while (walk.next(p, n, thi, twi, z0, nd)) {
                const int h0 = thi * TC_TH, w0 = twi * TC_TW;
                for (int pl = 0; pl < nd + K - 1; ++pl) {
                    for (int g = 0; g < p.ncg; ++g, pr.next(NS)) {
                        const int b0 = g * p.cbg;
                        const int gb = (p.cb - b0) < p.cbg ? (p.cb - b0) : p.cbg;
                        { TCPROF_T0(); mbar_wait(b_empty + 8 * pr.slot, pr.phase ^ 1); TCPROF_ADD(prof_a); }
#ifdef CTU_DBG_NOTMA
                        mbar_arrive(b_full + 8 * pr.slot);
                        for (int b = 0; b < 0; ++b) {
#else
                        mbar_expect_tx(b_full + 8 * pr.slot, (uint32_t)gb * HH * WW * 16);
                        for (int b = 0; b < gb; ++b) {
#endif
                            int s = 0;
#pragma unroll
                            for (int q = 1; q < CTU_MAX_SRC; ++q)
                                if (q < p.nsrc && b0 + b >= p.src_cboff[q]) s = q;
                            tma_load_4d(s_planes + pr.slot * p.slot_bytes + b * p.plane_bytes, &maps.m[s], (w0 - PAD) * 8,
                                        h0 - PAD, z0 + pl - PAD, n * p.src_cb[s] + (b0 + b - p.src_cboff[s]),
                                        b_full + 8 * pr.slot);
                        }
                    }
                }
            }
        }
    } else if (warp <= TC_WB) {
        // ===================================================================== MMA issuers (tile t = warp - 1)
        // All 32 lanes walk the schedule (so the descriptor arithmetic stays warp-uniform and branch-free); the
        // MMAs and commits of the elected lane are the ones that issue.  The order is the weight image's order
        // (tc_chunk): per group, (tap, block pair) with the tap outermost, then the odd block's tap pairs.
        {
            const uint32_t leader = elect_one();
            // instruction descriptor: D=f32, A=B=bf16, both K-major, N at [17,23), M=128 at [24,29)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.nt >> 3) << 17) | (8u << 24);
            const uint32_t a_hi = (ROW >> 4) | (1u << 14);           // SBO = row pitch (next h), version 1
            const uint32_t b_hi = (16u) | (1u << 14);                 // B: SBO = 256 B between n-groups
            const uint32_t t = warp - 1;
            const uint32_t plane16 = p.plane_bytes >> 4;
            const uint32_t lbo_pair = plane16 << 16;                  // second K chunk = the next channel block
            const uint32_t bstep = (uint32_t)p.nt * 2u;              // one MMA's B tile in 16-byte units
            constexpr int K2 = K * K, NTAIL = (K2 + 1) / 2;
            mbar_wait(b_w, 0);
            uint32_t step = 0;
            Ring cons = {0, 0};
            TcWalk walk;
            walk.init(p);
            int n_, thi_, twi_, z0_, nd;
            while (walk.next(p, n_, thi_, twi_, z0_, nd)) {
                for (int pl = 0; pl < nd + K - 1; ++pl, ++step) {
                    const uint32_t stage = step & p.stmask;
                    const uint32_t d_tmem = tmem_base + (stage * TC_WB + t) * p.nt;
                    uint32_t b_lo = (s_w >> 4) | (8u << 16);          // B: LBO = 128 B (second 8-wide K chunk)
                    uint32_t acc = 0;
                    for (int g = 0; g < p.ncg; ++g, cons.next(NS)) {
                        { TCPROF_T0(); mbar_wait(b_full + 8 * cons.slot, cons.phase); TCPROF_ADD(prof_a); }
                        if (g == 0) { TCPROF_T0(); mbar_wait(b_aempty + 8 * (stage * TC_WB + t), ((step >> p.stshift) & 1) ^ 1); TCPROF_ADD(prof_b); }
                        tc_fence_after();
                        TCPROF_T0();
                        const uint32_t a16 = (s_planes + cons.slot * p.slot_bytes + t * 128u) >> 4;
                        const int gb = (p.cb - g * p.cbg) < p.cbg ? (p.cb - g * p.cbg) : p.cbg;
                        const int pairs = gb >> 1;
                        if (pairs > 0) {
#pragma unroll
                            for (int tp = 0; tp < K2; ++tp) {
                                const uint32_t a_tap = a16 + (tp / K) * (ROW >> 4) + (tp % K);
                                for (int pr = 0; pr < pairs; ++pr) {
#ifdef CTU_DBG_NOMMA
                                    if (p.cb == 12345)
#endif
                                    umma_bf16_lead(leader, d_tmem, (a_tap + 2u * pr * plane16) | lbo_pair, a_hi, b_lo, b_hi,
                                                   idesc, acc);
                                    acc = 1;
                                    b_lo += bstep;
                                }
                            }
                        }
                        if (gb & 1) {
                            const uint32_t a_blk = a16 + (uint32_t)(gb - 1) * plane16;
#pragma unroll
                            for (int j = 0; j < NTAIL; ++j) {
                                // taps 2j and 2j+1 of the last block; the dummy half of the last MMA has zero weights
                                constexpr uint32_t R16 = ROW >> 4;
                                const int t0 = 2 * j, t1 = (2 * j + 1 < K2) ? 2 * j + 1 : 2 * j;
                                const uint32_t o0 = (t0 / K) * R16 + (t0 % K), o1 = (t1 / K) * R16 + (t1 % K);
#ifdef CTU_DBG_NOMMA
                                if (p.cb == 12345)
#endif
                                umma_bf16_lead(leader, d_tmem, (a_blk + o0) | ((o1 - o0) << 16), a_hi, b_lo, b_hi, idesc, acc);
                                acc = 1;
                                b_lo += bstep;
                            }
                        }
                        TCPROF_ADD(prof_c);
                        { TCPROF_T0(); umma_commit_lead(leader, b_empty + 8 * cons.slot); TCPROF_ADD(prof_d); }
                    }
                    { TCPROF_T0(); umma_commit_lead(leader, b_afull + 8 * (stage * TC_WB + t)); TCPROF_ADD(prof_d); }
                }
            }
        }
    } else {
        // ===================================================================== epilogue: 4 warps per tile
        const int te = (warp - 1 - TC_WB) >> 2;        // tile this warp serves
        const int quarter = warp & 3;                  // TMEM lane quarter this warp may read
        const int row = quarter * 32 + lane;           // row of the 128-row MMA tile
        const int hh = row >> 3, wl = row & 7;
        const bool want_stats = p.stats != nullptr;
        const long long plane = (long long)p.d * p.h * p.w;
        const int ob0 = blockIdx.y * COB;              // output blocks [ob0, ob0 + nob) belong to this CTA row
        const int nob = (p.cob_n - ob0) < COB ? (p.cob_n - ob0) : COB;
        float part[K][CP];                             // partial sums of the K output planes in flight
        float s1[NSLOT * 8], s2[NSLOT * 8];
        const bool has_bias = p.bias != nullptr;       // only the legacy 5^3 family: read on use, not held in registers
#pragma unroll
        for (int i = 0; i < NSLOT * 8; ++i) s1[i] = s2[i] = 0.f;
        uint32_t step = 0;
        TcWalk walk;
        walk.init(p);
        int n, thi, twi, z0, nd;
        while (walk.next(p, n, thi, twi, z0, nd)) {
            const int h0 = thi * TC_TH, w0 = twi * TC_TW;
            const int gy = h0 + hh, gx = w0 + te * 8 + wl;
            const bool inb = gy < p.h && gx < p.w;
            __nv_bfloat16* ycol = p.y + (((long long)n * p.cob_n + ob0) * plane + ((long long)gy) * p.w + gx) * 8;
#pragma unroll
            for (int i = 0; i < K; ++i)
#pragma unroll
                for (int c = 0; c < CP; ++c) part[i][c] = 0.f;
            // The partial sums of output plane j live in part[j % K]: the plane loop is unrolled K-fold so that every slot
            // index is a compile-time constant (no register rotation per plane).
            const int npl = nd + K - 1;
            for (int pl0 = 0; pl0 < npl; pl0 += K) {
#pragma unroll
            for (int u = 0; u < K; ++u) {
                const int pl = pl0 + u;
                if (pl >= npl) break;
                const uint32_t stage = step & p.stmask;
                { TCPROF_T0(); mbar_wait(b_afull + 8 * (stage * TC_WB + te), (step >> p.stshift) & 1); TCPROF_ADD(prof_a); }
                ++step;
                TCPROF_T0();
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (stage * TC_WB + te) * p.nt;
                // P[kd] of input plane pl feeds output plane pl - kd, kept in part[(pl - kd) % K] = part[(u - kd + K) % K]
                // all loads of up to two output blocks in flight, one wait
                constexpr int OBB = COB < 2 ? COB : 2;
#ifdef CTU_DBG_NOEPI
                if (taddr == 0xffffffffu) part[0][0] += 1.f;
                for (int ob0b = 0; ob0b < 0; ob0b += OBB) {
#else
#pragma unroll
                for (int ob0b = 0; ob0b < COB; ob0b += OBB) {
#endif
                    uint32_t raw[K][OBB][8];
#pragma unroll
                    for (int kd = 0; kd < K; ++kd)
#pragma unroll
                        for (int ob = 0; ob < OBB; ++ob) tmem_ld8_issue(taddr + kd * CP + (ob0b + ob) * 8, raw[kd][ob]);
                    tmem_ld_wait();
#pragma unroll
                    for (int kd = 0; kd < K; ++kd)
#pragma unroll
                        for (int ob = 0; ob < OBB; ++ob) {
                            tmem_ld8_pin(raw[kd][ob]);
#pragma unroll
                            for (int c = 0; c < 8; ++c)
                                part[(u - kd + K) % K][(ob0b + ob) * 8 + c] += __uint_as_float(raw[kd][ob][c]);
                        }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(b_aempty + 8 * (stage * TC_WB + te));
                TCPROF_ADD(prof_b);
                // output plane j = pl - (K-1) is complete; its slot (u + 1) % K then starts over for plane pl + 1
                const int j = pl - (K - 1);
                float (&done)[CP] = part[(u + 1) % K];
                if (j >= 0) {
                    const int gz = z0 + j;
#pragma unroll
                    for (int ob = 0; ob < COB; ++ob) {
                        V8 o;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            float v = done[ob * 8 + c];
                            if (has_bias) v += st_bias[ob * 8 + c];
                            o.v[c] = round_to<__nv_bfloat16>(v);
                        }
                        if (inb && ob < nob) {
#ifndef CTU_DBG_NOSTORE
                            Vec8<__nv_bfloat16>::store(ycol + ((long long)ob * plane + (long long)gz * p.h * p.w) * 8, o);
#else
                            if (o.v[0] == 1234.567f) Vec8<__nv_bfloat16>::store(ycol, o);
#endif
                            if (BNRED) {
                                long long yoff;
                                if (p.bn_pm) {      // y phase-major: [n][8 phases x cob][d/2][h/2][w/2][8]
                                    const int q = ((gz & 1) << 2) | ((gy & 1) << 1) | (gx & 1);
                                    const long long ps = ((long long)(gz >> 1) * (p.h >> 1) + (gy >> 1)) * (p.w >> 1) + (gx >> 1);
                                    yoff = ((((long long)n * 8 + q) * p.cob_n + ob0 + ob) * (plane >> 3) + ps) * 8;
                                } else {
                                    yoff = (((long long)n * p.cob_n + ob0 + ob) * plane + ((long long)gz * p.h + gy) * p.w + gx) * 8;
                                }
                                const V8 yv = Vec8<__nv_bfloat16>::load(p.bn_y + yoff);
#pragma unroll
                                for (int c = 0; c < 8; ++c) {
                                    const int cc = ob * 8 + c;
                                    const float av = fmaf(yv.v[c], st_bn[cc], st_bn[CP + cc]);
                                    const float dz = av > 0.f ? o.v[c] : 0.f;
                                    const float xhat = (yv.v[c] - st_bn[2 * CP + cc]) * st_bn[3 * CP + cc];
                                    s1[(ob % NSLOT) * 8 + c] += dz;
                                    s2[(ob % NSLOT) * 8 + c] = fmaf(dz, xhat, s2[(ob % NSLOT) * 8 + c]);
                                }
                            } else if (want_stats) {
#pragma unroll
                                for (int c = 0; c < 8; ++c) {
                                    s1[(ob % NSLOT) * 8 + c] += o.v[c];
                                    s2[(ob % NSLOT) * 8 + c] = fmaf(o.v[c], o.v[c], s2[(ob % NSLOT) * 8 + c]);
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < CP; ++c) done[c] = 0.f;
            }
            }
        }
        if (want_stats) {
            // warp totals -> shared memory; one atomic per channel and CTA after the teardown barrier (same-address
            // double atomics from every warp of every CTA serialise at ~1 ns each: tens of microseconds of tail)
            float* red = st_red + (warp - 1 - TC_WB) * (NSLOT * 16);
#pragma unroll
            for (int i = 0; i < NSLOT * 8; ++i) {
                const float a1 = warp_sum(s1[i]), a2 = warp_sum(s2[i]);
                if (lane == 0) {
                    red[i] = a1;
                    red[NSLOT * 8 + i] = a2;
                }
            }
        }
    }
    // ------------------------------------------------------------------------- teardown
#ifdef CTU_TCPROF
    if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && (warp <= 1 || warp == 3)) {
        const long long tot = clock64() - prof_start;
        if (warp == 0) printf("TCPROF producer: total %lld wait_empty %lld\n", tot, prof_a);
        else if (warp == 1) printf("TCPROF issuer  : total %lld wait_full %lld wait_aempty %lld mma %lld commit %lld\n", tot, prof_a, prof_b, prof_c, prof_d);
        else printf("TCPROF epilogue: total %lld wait_afull %lld ld+add+release %lld\n", tot, prof_a, prof_b);
    }
#endif
    tc_fence_before();
    __syncthreads();
    if (p.stats != nullptr && threadIdx.x < NSLOT * 16) {
        // accumulator slot s holds the output blocks ob0 + s, ob0 + s + NSLOT, ...: all the same natural block
        // (the host only fuses the statistics when that holds)
        const int i = threadIdx.x % (NSLOT * 8), q = threadIdx.x / (NSLOT * 8);     // q: 0 = sum, 1 = sum of squares
        const int ob0 = blockIdx.y * COB;
        const int nob = (p.cob_n - ob0) < COB ? (p.cob_n - ob0) : COB;
        const int ch = ((ob0 + (i >> 3)) % p.cobo) * 8 + (i & 7);
        if ((i >> 3) < nob && ch < p.cstat) {
            double tot = 0.0;
#pragma unroll
            for (int wq = 0; wq < 4 * TC_WB; ++wq) tot += (double)st_red[wq * (NSLOT * 16) + q * (NSLOT * 8) + i];
            atomicAdd(p.stats + q * (p.cobo * 8) + ch, tot);
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host
// TcWalk mode of a launch.  Equal plane shares remove the round quantisation of fixed d-chunks, but the CTAs then sit at
// staggered depths (neighbouring tiles no longer read their common halo at the same time) and a share that straddles a
// column end pays the K - 1 halo planes twice.  Measured on B200 (scripts/bench_kernels.py, us, chunks -> shares):
//   fprop 7->7 at 4x128^3 94.9 -> 84.7, 2->7 95.2 -> 85.6, 64->14 at 4x64^3 88.3 -> 84.5, 14->14 at 4x64^3 47.5 -> 47.6;
//   but 4-block outputs 146 -> 169 ([14,14,1]->64 at 4x64^3), 53 -> 58 (128->28 at 4x32^3), 30 -> 33 (28->28 at 4x32^3);
//   wgrad (first formulation) 244 -> 240, 135 -> 127, 84.4 -> 79.5, 49.3 -> 47.0; (kd,kh)-in-N formulation 116 -> 124.
// (5^3 layers -- four halo planes per segment: one-block kernel, two CTAs per SM, 8->8 at 4x128^3 194 -> 174 us; two-block
//  kernel 16->16 at 4x64^3 168 -> 187 us.)
// Hence: shares for 3^3 fprop / dgrad with <= 2 output blocks per CTA, 5^3 with one, >= 16 planes per CTA, and for the first wgrad
// formulation; fixed chunks elsewhere.  CTU_TC_BALANCE=0 / 1 forces one mode everywhere (A/B runs).
static int tc_balance(bool by_rule) {
    static const int forced = getenv("CTU_TC_BALANCE") ? atoi(getenv("CTU_TC_BALANCE")) : -1;
    return forced >= 0 ? (forced ? 1 : 0) : (by_rule ? 1 : 0);
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
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
    int cb, cbg, ncg, cob_n, cobg, ngroups, nt, nm;
    uint32_t wimg_bytes, plane_bytes, slot_bytes, tmem_cols, ns;   // wimg_bytes: ONE group's image
    uint32_t stages;                                               // TMEM accumulator stages per 16x8 tile
    size_t smem;
};

// Ring depth: as deep as a ~100 KB budget allows (two CTAs per SM), at least `min_slots`, at most TC_MAX_SLOTS.
static uint32_t pick_slots(int min_slots, size_t fixed_bytes, size_t slot_bytes, size_t budget = 100 * 1024) {
    long long ns = fixed_bytes < budget ? (long long)((budget - fixed_bytes) / slot_bytes) : 0;
    if (ns < min_slots) ns = min_slots;
    if (ns > TC_MAX_SLOTS) ns = TC_MAX_SLOTS;
    return (uint32_t)ns;
}

// cb: total channel blocks over all concatenated sources
static bool tc_geometry(int k, int cb, int cout, int h, int w, TcGeom& g) {
    if (k != 3 && k != 5) return false;
    if (h % TC_TH || w % TC_TW || cb < 1 || cout < 1) return false;
    g.cb = cb;
    g.cbg = cb <= 6 ? cb : 4;          // wide inputs are staged four blocks at a time
    g.ncg = tc_groups(cb, g.cbg);
    g.cob_n = (cout + 7) / 8;
    g.nm = tc_mmas_per_kd(k, g.cb, g.cbg);
    const int hh = TC_TH + k - 1, ww = TC_TW + k - 1;
    g.plane_bytes = ((uint32_t)hh * ww * 16 + 127u) & ~127u;
    g.slot_bytes = g.plane_bytes * g.cbg;
    // output blocks per launch: the instantiated COB in {1,2,4} (k=3) / {1,2} (k=5) -- N = K*COB*8 <= 128 so that
    // 2 stages x 2 tiles fit the 512 TMEM columns -- shrunk until weights + a 3-slot ring fit shared memory
    int cobg = (k == 3) ? 4 : 2;
    while (cobg > 1 && cobg / 2 >= g.cob_n) cobg /= 2;
    for (;; cobg /= 2) {
        g.cobg = cobg;
        g.nt = (k * cobg * 8 + 15) / 16 * 16;
        g.wimg_bytes = (uint32_t)g.nm * g.nt * 32;
        const size_t fixed = ((g.wimg_bytes + 1023u) & ~1023u) + ((g.nm * 8 + 15) & ~15) +
                             8 * (2 * TC_MAX_SLOTS + 2 * TC_MAX_STAGES * TC_WB + 2) + 32 + 4 * TC_WB * 32 * 4 + 128 + 512 + 1024;
        // the wide-output variants are limited to one CTA per SM by registers: give their ring the whole SM
        g.ns = pick_slots(g.ncg > 1 ? 4 : 3, fixed, g.slot_bytes, cobg >= 2 ? 200 * 1024 : 100 * 1024);
        g.smem = fixed + (size_t)g.ns * g.slot_bytes;
        if (g.smem > 220 * 1024 && g.ns > 3) {
            g.ns = 3;
            g.smem = fixed + (size_t)g.ns * g.slot_bytes;
        }
        if (g.smem <= 220 * 1024) break;
        if (cobg == 1) return false;
    }
    g.ngroups = (g.cob_n + g.cobg - 1) / g.cobg;
    // accumulator stages per tile: two.  (Four -- CTU_TC_STAGES=4, possible while 4 x 2 x N <= 512 TMEM columns -- were measured
    // neutral on the narrow layers: 7->7 at 4x128^3 84.7 vs 84.5 us; the issuers do not wait on the epilogue.)
    g.stages = 2u;
    static const int stages_env = getenv("CTU_TC_STAGES") ? atoi(getenv("CTU_TC_STAGES")) : 0;
    if (stages_env == 4 && (uint32_t)(4 * TC_WB * g.nt) <= 512u) g.stages = 4u;
    uint32_t cols = g.stages * TC_WB * g.nt;
    g.tmem_cols = 32;
    while (g.tmem_cols < cols) g.tmem_cols *= 2;
    return g.tmem_cols <= 512;
}

static int total_blocks(int nsrc, const int* h_src_channels) {
    if (nsrc < 1 || nsrc > CTU_MAX_SRC || h_src_channels == nullptr) return -1;
    int cb = 0;
    for (int i = 0; i < nsrc; ++i) {
        if (h_src_channels[i] < 1) return -1;
        cb += (h_src_channels[i] + 7) / 8;
    }
    return cb;
}

int make_map(CUtensorMap* map, const void* ptr, int nblocks, int d, int h, int w, int box_w, int box_h) {
    auto encode = get_encode();
    if (!encode) {
        set_error("cuTensorMapEncodeTiled entry point not found");
        return CTU_ERR_UNSUPPORTED;
    }
    const cuuint64_t gdim[4] = {(cuuint64_t)w * 8, (cuuint64_t)h, (cuuint64_t)d, (cuuint64_t)nblocks};
    const cuuint64_t gstr[3] = {(cuuint64_t)w * 16, (cuuint64_t)h * w * 16, (cuuint64_t)d * h * w * 16};
    const cuuint32_t box[4] = {(cuuint32_t)box_w * 8, (cuuint32_t)box_h, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d)", (int)cr);
        return CTU_ERR_INVALID;
    }
    return CTU_OK;
}

// one halo-box tensor map per concatenated source
static int make_src_maps(TcMaps& maps, const void* const* h_srcs, const int* h_src_channels, int nsrc, int n, int d, int h,
                         int w, int k) {
    for (int i = 0; i < CTU_MAX_SRC; ++i) {
        const int j = i < nsrc ? i : 0;
        int rc = make_map(&maps.m[i], h_srcs[j], n * ((h_src_channels[j] + 7) / 8), d, h, w, TC_TW + k - 1, TC_TH + k - 1);
        if (rc != CTU_OK) return rc;
    }
    return CTU_OK;
}

// stat_cout: number of NATURAL channels the statistics are taken over: cout for an ordinary convolution, cout / 8
// phases for the phase-major output of the fused up-sampling stage (natural block = output block % ceil(stat_cout/8))
int conv3d_fprop_tc(const void* const* h_srcs, const int* h_src_channels, int nsrc, const float* wp, const float* bias,
                    void* y, double* stats, int stat_cout, int cout, int k, int n, int d, int h, int w, cudaStream_t stream,
                    const void* bn_y, const float* bn_ss, int bn_pm, int prezeroed) {
    TcGeom g;
    const int cb = total_blocks(nsrc, h_src_channels);
    if (cb < 0 || !tc_geometry(k, cb, cout, h, w, g)) {
        set_error("conv3d tensor path: shape k=%d blocks=%d cout=%d %dx%dx%d not covered", k, cb, cout, d, h, w);
        return CTU_ERR_UNSUPPORTED;
    }
    TcMaps maps;
    int rc = make_src_maps(maps, h_srcs, h_src_channels, nsrc, n, d, h, w, k);
    if (rc != CTU_OK) return rc;
    TcParams p = {};
    p.wimg = reinterpret_cast<const __nv_bfloat16*>(wp);
    p.bias = bias;
    p.y = reinterpret_cast<__nv_bfloat16*>(y);
    if (stat_cout <= 0) stat_cout = cout;
    p.cobo = (stat_cout + 7) / 8;
    p.cstat = stat_cout;
    const bool fuse_stats = stats != nullptr && (g.cobg <= 2 || p.cobo <= 2);
    p.stats = fuse_stats ? stats : nullptr;
    const bool bnred = bn_y != nullptr;
    if (bnred && !(k == 3 && g.cobg <= 2 && bn_ss != nullptr && stats != nullptr && stat_cout == cout)) {
        set_error("conv3d tensor path: BatchNorm-backward reduction fused only for 3x3x3 with <= 16 output channels");
        return CTU_ERR_UNSUPPORTED;
    }
    p.bn_y = reinterpret_cast<const __nv_bfloat16*>(bn_y);
    p.bn_ss = bn_ss;
    p.bn_pm = bn_pm;
    p.cb = g.cb; p.cbg = g.cbg; p.ncg = g.ncg; p.cob_n = g.cob_n; p.nt = g.nt; p.cout = cout;
    p.nsrc = nsrc;
    for (int i = 0, off = 0; i < CTU_MAX_SRC; ++i) {
        p.src_cb[i] = i < nsrc ? (h_src_channels[i] + 7) / 8 : 0;
        p.src_cboff[i] = off;
        off += p.src_cb[i];
    }
    p.n = n; p.d = d; p.h = h; p.w = w;
    p.tiles_h = h / TC_TH; p.tiles_w = w / TC_TW;
    const int tiles = n * p.tiles_h * p.tiles_w;
    p.wimg_bytes = g.wimg_bytes; p.plane_bytes = g.plane_bytes; p.slot_bytes = g.slot_bytes; p.tmem_cols = g.tmem_cols;
    p.ns = g.ns;
    p.stmask = g.stages - 1;
    p.stshift = g.stages == 4 ? 2 : 1;
    if (fuse_stats && !prezeroed) {
        cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(double) * 2 * p.cobo * 8, stream);
        if (e != cudaSuccess) {
            set_error("conv3d tensor path: memset: %s", cudaGetErrorString(e));
            return (int)e;
        }
    }
    int ctas_per_sm = (int)((227 * 1024) / (g.smem + 1024));
    if (ctas_per_sm > 512 / (int)g.tmem_cols) ctas_per_sm = 512 / (int)g.tmem_cols;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm > 2) ctas_per_sm = 2;
    const int threads = 32 * (1 + TC_WB + 4 * TC_WB);
    auto go = [&](auto kern) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
        if (e != cudaSuccess) {
            set_error("conv3d tensor path: smem %zu: %s", g.smem, cudaGetErrorString(e));
            return (int)e;
        }
        // persistent grid = exactly the CTAs that are resident at once (registers limit the wide-output variants
        // to one per SM); one grid row per group of COB output blocks -- all rows stream the same input planes, so
        // their re-reads hit L2
        // (shared memory and TMEM give ctas_per_sm; the register file is accounted for here -- the occupancy API
        // was observed to answer 1 where two CTAs do become resident)
        int occ = ctas_per_sm;
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, kern) == cudaSuccess && fa.numRegs > 0) {
            const int regs_per_warp = (fa.numRegs * 32 + 255) / 256 * 256;
            const int by_regs = 65536 / (regs_per_warp * (threads / 32));
            if (occ > by_regs) occ = by_regs;
        }
        if (occ < 1) occ = 1;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        // d-chunk: every chunk walks dc + K - 1 input planes (K - 1 of them halo re-reads AND re-computed MMAs), and the
        // resident CTAs take the items in rounds: minimise rounds x planes per item (ties: the smaller chunk, for balance)
        int dc = d;
        {
            const long long slots = (148LL * occ) / g.ngroups > 0 ? (148LL * occ) / g.ngroups : 1;
            long long best = -1;
            for (int c = d; c >= 1; --c) {
                const long long items = (long long)tiles * ((d + c - 1) / c);
                const long long rounds = (items + slots - 1) / slots;
                const long long cost = rounds * (c + k - 1);
                if (best < 0 || cost <= best) {
                    best = cost;
                    dc = c;
                }
            }
        }
        p.dc = dc;
        p.dchunks = (d + dc - 1) / dc;
        p.total_items = tiles * p.dchunks;
        int gx = (148 * occ) / g.ngroups;
        p.balance = tc_balance(((k == 3 && g.cobg <= 2) || (k == 5 && g.cobg == 1)) && (long long)tiles * d >= 16LL * (gx > 0 ? gx : 1));
        const long long units = p.balance ? (long long)tiles * d : (long long)p.total_items;
        if (gx > units) gx = (int)units;
        if (gx < 1) gx = 1;
        kern<<<dim3(gx, g.ngroups), threads, g.smem, stream>>>(maps, p);
        return check_launch("ctu_conv3d_fprop(tcgen05)");
    };
    if (bnred && g.cobg == 1) rc = go(conv3d_tc_kernel<3, 1, true>);
    else if (bnred) rc = go(conv3d_tc_kernel<3, 2, true>);
    else if (k == 3 && g.cobg == 1) rc = go(conv3d_tc_kernel<3, 1>);
    else if (k == 3 && g.cobg == 2) rc = go(conv3d_tc_kernel<3, 2>);
    else if (k == 3 && g.cobg == 4) rc = go(conv3d_tc_kernel<3, 4>);
    else if (k == 5 && g.cobg == 1) rc = go(conv3d_tc_kernel<5, 1>);
    else rc = go(conv3d_tc_kernel<5, 2>);
    if (rc == CTU_OK && stats != nullptr && !fuse_stats)   // wide layers (low resolution): separate statistics pass
        rc = ctu_bn_stats(CTU_BF16, y, stat_cout, (g.cob_n / p.cobo) | (prezeroed ? CTU_ACCUM_PREZEROED : 0), n, (long long)d * h * w, stats, stream);
    return rc;
}

// ================================================================================================ wgrad
// dW[co, ci, kd, kh, kw] = sum_v dy[co, v] * x[ci, v + tap - pad]  as UMMA with the VOXELS as the K dimension.
// Both operands are MN-major straight out of the blocked layout (for one voxel the 8 channels are contiguous):
//   A = x halo row:  M = 64 = 8 "lag" groups x 8 input lanes, group g = the row shifted by g voxels (SBO = 16 B),
//       so lags 0..K-1 are the kw taps (lags K..7 are computed and discarded: 3/8 or 5/8 of the MMA is useful);
//   B = K dy rows:   N = K (kh) x output channels.  The dy tile is staged as [h][cob][w][8] (one 5-D TMA box), so
//       for the x halo row rho the rows rho-(K-1)..rho of dy (kh = K-1..0) x all channel blocks are n-groups at a
//       uniform SBO of 256 B; K-1 zero rows above and below the tile stand for the rows owned by neighbour tiles;
//   K = 16 consecutive w-voxels (two 8-voxel core matrices, LBO = 128 B).
// The MMA is bound by the A-operand read from shared memory (~64 B/clk), so carrying the kh taps in N (instead of
// one MMA per kh) cuts the work to HH = 16+K-1 MMAs per (plane, kd, input block).
// One TMEM accumulator [64 x K*Ng] per (kd, input block) lives for the whole CTA; a CTA walks its share of the
// (n, h-tile, w-tile, d-chunk) items and flushes K^3 x 8 x Cbg x Ng partial sums with fp32 atomics at the end.
struct WgParams {
    float* dwp;                 // packed fp32 [cob][cib][tap][ci][co], zeroed by the caller
    int cb, cob_n;              // totals
    int cbg, ng;                // input blocks / output channels handled per CTA group
    int n_cbgroups, n_ngroups;
    int n, d, h, w;
    int tiles_h, tiles_w, dchunks, dc, total_items;
    int balance;                // TcWalk: equal shares of the plane sequence instead of round-robin d-chunks
    uint32_t plane_bytes, xslot_bytes, dyslot_bytes, tmem_cols;
    uint32_t ns, nds;           // x-plane / dy-plane ring depths
    int nsrc, src_cb[CTU_MAX_SRC], src_cboff[CTU_MAX_SRC];   // concatenated sources of x
    // phase-sparse mode (weight gradient of the fused up-sampling stage, K = 3, output channel = q*8P*... phase-major):
    // a group of output channels = the phases with one qd (sparse == 1, 4P blocks) or one (qd, qh) (sparse == 2, 2P
    // blocks); only the taps kd in {qd, qd+1} and kh in {qh, qh+1} of the composed weights are structurally non-zero,
    // so every (kd, input block) accumulator covers 2 of 3 kd and a contiguous run of 8P (4P) n-groups.  0 = dense.
    int sparse, pblk;           // pblk = P = channel blocks per phase
};

template <int K>
__global__ void __launch_bounds__(WG_THREADS) conv3d_wgrad_tc_kernel(const __grid_constant__ TcMaps xmaps,
                                                                     const __grid_constant__ CUtensorMap dymap,
                                                                     WgParams p) {
    constexpr int PAD = K / 2;
    constexpr int HH = TC_TH + K - 1, WW = TC_TW + K - 1;
    constexpr uint32_t ROW = WW * 16;
    constexpr uint32_t DYBLK = TC_TW * 16;         // one channel block of one dy row: 16 voxels x 16 B
    const uint32_t NS = p.ns, NDS = p.nds;
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_x = s_base;
    const uint32_t s_dy = s_x + NS * p.xslot_bytes;
    const uint32_t s_bar = s_dy + NDS * p.dyslot_bytes;
    const uint32_t b_xfull = s_bar, b_xempty = s_bar + 8 * TC_MAX_SLOTS, b_dyfull = s_bar + 16 * TC_MAX_SLOTS,
                   b_dyempty = s_bar + 24 * TC_MAX_SLOTS, b_done = s_bar + 32 * TC_MAX_SLOTS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (b_done + 8 - s_base));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = blockIdx.y;
    const int cbg_i = grp % p.n_cbgroups, ng_i = grp / p.n_cbgroups;
    const int cb0 = cbg_i * p.cbg;
    const int ncb = (p.cb - cb0) < p.cbg ? (p.cb - cb0) : p.cbg;         // input blocks of this group
    const int ob0 = ng_i * (p.ng / 8);
    // valid output blocks of this group; the LAYOUT (TMA box, row pitch, MMA N) always spans the full group of ng/8
    // blocks -- a partial last group is zero-filled by TMA (out-of-range coordinates) and skipped at the flush
    const int nob = (p.cob_n - ob0) < (p.ng / 8) ? (p.cob_n - ob0) : (p.ng / 8);
    const int nobx = p.ng / 8;
    const int nn = nobx * 8;                                             // output channels of this group
    const uint32_t dyrow = (uint32_t)nobx * DYBLK;                       // one dy row, all channel blocks

    // zero rows above / below every dy slot (never written by TMA), visible to the async proxy before any MMA
    for (uint32_t sl = 0; sl < NDS; ++sl) {
        uint4* top = reinterpret_cast<uint4*>(smem + (s_dy - s_base) + sl * p.dyslot_bytes);
        uint4* bot = reinterpret_cast<uint4*>(smem + (s_dy - s_base) + sl * p.dyslot_bytes + (K - 1 + TC_TH) * dyrow);
        for (uint32_t i = threadIdx.x; i < (K - 1) * dyrow / 16; i += WG_THREADS) {
            top[i] = make_uint4(0, 0, 0, 0);
            bot[i] = make_uint4(0, 0, 0, 0);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < NS; ++i) {
            mbar_init(b_xfull + 8 * i, 1);
            mbar_init(b_xempty + 8 * i, WG_ISSUERS);
        }
        for (uint32_t i = 0; i < NDS; ++i) {
            mbar_init(b_dyfull + 8 * i, 1);
            mbar_init(b_dyempty + 8 * i, WG_ISSUERS);
        }
        mbar_init(b_done, WG_ISSUERS);
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
    if (warp == 0) {
        if (lane == 0) {
            Ring pr = {0, 0}, dr = {0, 0};
            TcWalk walk;
            walk.init(p);
            int n, thi, twi, z0, nd;
            while (walk.next(p, n, thi, twi, z0, nd)) {
                const int h0 = thi * TC_TH, w0 = twi * TC_TW;
                for (int pl = 0; pl < nd + K - 1; ++pl, pr.next(NS)) {
                    mbar_wait(b_xempty + 8 * pr.slot, pr.phase ^ 1);
                    mbar_expect_tx(b_xfull + 8 * pr.slot, (uint32_t)ncb * HH * WW * 16);
                    for (int b = 0; b < ncb; ++b) {
                        int s = 0;
#pragma unroll
                        for (int q = 1; q < CTU_MAX_SRC; ++q)
                            if (q < p.nsrc && cb0 + b >= p.src_cboff[q]) s = q;
                        tma_load_4d(s_x + pr.slot * p.xslot_bytes + b * p.plane_bytes, &xmaps.m[s], (w0 - PAD) * 8, h0 - PAD,
                                    z0 + pl - PAD, n * p.src_cb[s] + (cb0 + b - p.src_cboff[s]), b_xfull + 8 * pr.slot);
                    }
                    if (pl >= K - 1) {
                        mbar_wait(b_dyempty + 8 * dr.slot, dr.phase ^ 1);
                        mbar_expect_tx(b_dyfull + 8 * dr.slot, (uint32_t)TC_TH * dyrow);
                        tma_load_5d(s_dy + dr.slot * p.dyslot_bytes + (K - 1) * dyrow, &dymap, w0 * 8, ob0, h0,
                                    z0 + pl - (K - 1), n, b_dyfull + 8 * dr.slot);
                        dr.next(NDS);
                    }
                }
            }
        }
    } else {
        // ===================================================================== MMA issuers + flush (warps 1..4)
        // D=f32, A=B=bf16, both MN-major (bits 15, 16), N at [17,23), M=64 at [24,29)
        // dense: N = kh taps x output channels, K accumulators (kd) per input block.  phase-sparse: see WgParams
        const int P = p.pblk;
        const int qd = p.sparse == 1 ? ng_i : (ng_i >> 1), qh = ng_i & 1;
        const int nkd = p.sparse ? 2 : K, kd0 = p.sparse ? qd : 0;
        const int ncol = p.sparse == 1 ? 64 * P : (p.sparse == 2 ? 32 * P : K * nn);
        // first n-group of the MMA inside the (row r) window of the dy slot, in 16-byte units
        const uint32_t b_off16 = p.sparse == 1 ? (uint32_t)(2 * P) * (DYBLK >> 4)
                                               : (p.sparse == 2 ? (uint32_t)(1 - qh) * (dyrow >> 4) : 0u);
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(ncol >> 3) << 17) | (4u << 24);
        const uint32_t a_hi = 1u | (1u << 14);                    // SBO = 16 B: m-group g = lag g
        const uint32_t b_hi = (DYBLK >> 4) | (1u << 14);          // SBO = 256 B: next (row, channel block) n-group
        const uint32_t lbo = 8u << 16;                            // second 8-voxel core matrix: +128 B
        const int q = warp - 1;
        const int nacc = nkd * ncb;                               // one accumulator per (kd, input block)
        // All 32 lanes walk the schedule (warp-uniform, branch-free descriptor arithmetic); only the elected lane's
        // MMAs / commits issue.  Per accumulator e = (kd, input block): HH back-to-back MMAs, A and B descriptors
        // advancing by one halo row / one dy row.
        const uint32_t leader = elect_one();
        uint32_t base = 0, waited = 0, first = 1;
        Ring cons = {0, 0}, head = {0, 0}, dyr = {0, 0};
        TcWalk walk;
        walk.init(p);
        int n_, thi_, twi_, z0_, nd;
        while (walk.next(p, n_, thi_, twi_, z0_, nd)) {
            for (int j = 0; j < nd; ++j) {
                while (waited < base + j + K) {
                    mbar_wait(b_xfull + 8 * cons.slot, cons.phase);
                    cons.next(NS);
                    ++waited;
                }
                mbar_wait(b_dyfull + 8 * dyr.slot, dyr.phase);
                tc_fence_after();
                const uint32_t dy16 = (((s_dy + dyr.slot * p.dyslot_bytes) >> 4) + b_off16) | lbo;
                int b = q % ncb, kd = q / ncb;                // accumulator e = q, q + WG_ISSUERS, ...
                for (int e = q; e < nacc; e += WG_ISSUERS) {
                    uint32_t slot = head.slot + kd0 + kd;
                    if (slot >= NS) slot -= NS;
                    const uint32_t x16 = ((s_x + slot * p.xslot_bytes + b * p.plane_bytes) >> 4) | lbo;
                    const uint32_t tcol = tmem_base + (uint32_t)e * ncol;
#pragma unroll
                    for (int r = 0; r < HH; ++r)    // x halo row r meets dy rows r-(K-1)..r (slot rows r..r+K-1)
                        umma_bf16_lead(leader, tcol, x16 + r * (ROW >> 4), a_hi, dy16 + r * (dyrow >> 4), b_hi, idesc,
                                       (r == 0) ? (first ^ 1u) : 1u);
                    b += WG_ISSUERS;
                    while (b >= ncb) {
                        b -= ncb;
                        ++kd;
                    }
                }
                first = 0;
                umma_commit_lead(leader, b_xempty + 8 * head.slot);
                umma_commit_lead(leader, b_dyempty + 8 * dyr.slot);
                head.next(NS);
                dyr.next(NDS);
            }
            for (int qq = 0; qq < K - 1; ++qq, head.next(NS)) umma_commit_lead(leader, b_xempty + 8 * head.slot);
            base += nd + K - 1;
        }
        umma_commit_lead(leader, b_done);
        __syncwarp();
        // flush: M=64 accumulators sit in lanes 0-15 of every 32-lane quarter (row = quarter*16 + lane)
        const int quarter = warp & 3;
        mbar_wait(b_done, 0);
        tc_fence_after();
        const int row = quarter * 16 + (lane & 15);
        const int lag = row >> 3, ci = row & 7;
        const bool useful = lane < 16 && lag < K;
        const int taps = K * K * K;
        for (int a = 0; a < nacc; ++a) {
            const int b = a % ncb, kd = kd0 + a / ncb;
            for (int g = 0; g < ncol / 8; ++g) {           // n-group g of the accumulator -> (kh, output block)
                int kh, ob;
                if (p.sparse == 1) {                       // [kh=2: blocks 2P..4P) | kh=1: all 4P | kh=0: blocks 0..2P)
                    if (g < 2 * P) { kh = 2; ob = 2 * P + g; }
                    else if (g < 6 * P) { kh = 1; ob = g - 2 * P; }
                    else { kh = 0; ob = g - 6 * P; }
                } else if (p.sparse == 2) {                // [kh = qh+1: 2P blocks | kh = qh: 2P blocks]
                    kh = g < 2 * P ? qh + 1 : qh;
                    ob = g < 2 * P ? g : g - 2 * P;
                } else {                                   // n-group row g / nob holds kh = K-1-row
                    kh = K - 1 - g / nobx;
                    ob = g % nobx;
                }
                if (ob >= nob) continue;
                const int tap = (kd * K + kh) * K + lag;
                float v[8];
                tmem_ld8(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)a * ncol + g * 8, v);
                if (useful) {
                    float* dst = p.dwp + ((((long long)(ob0 + ob) * p.cb + cb0 + b) * taps + tap) * 64 + ci * 8);
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (v[c] != 0.f) atomicAdd(dst + c, v[c]);
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


// ================================================================================================ wgrad, kd and kh in N
// Second formulation of the 3x3x3 weight gradient for narrow layers (round 2).  The kernel above issues one MMA per
// (dy plane, kd, x halo row) with N = kh x Cout = 24..48: for 8-channel layers the tensor pipe is saturated by MMA COUNT
// (tiny N, ~25-29 cycles each at 8 issuing warps per SM) while DRAM idles at 18 %.  Here the work is owned by X planes
// and rows: x row (z, y) of a 16 x 16 tile (no halo rows or planes) meets the 3 x 3 neighbourhood of dy rows
// (z - 1 .. z + 1) x (y - 1 .. y + 1).  The dy box is fetched by ONE 5-D TMA per step through a tensor map whose
// dimension order is (w, cob, d, h, n), so it lands in shared memory as [h (18 rows)][d (3 planes)][cob][w][8]: for a fixed
// x row the nine (kh, kd) rows x all output blocks are n-groups at a uniform 256-byte stride, i.e. ONE MMA
// (M = 64 = 8 kw lags x 8 input lanes, N = 9 x Cout, K = 16 voxels) replaces three, and the d / h borders (volume edge or
// neighbouring tile) are TMA's out-of-bounds zero fill -- every product is computed exactly once, by the owner of the x
// voxel.  16 MMAs per (plane, tile, input block) instead of 54.  Accumulators [64 x 9*Cout] per (input block, row parity)
// stay in TMEM for the whole CTA and are flushed with fp32 atomics into the packed gradient.
struct Wg2Params {
    float* dwp;
    int cb, cob_n;              // totals
    int cbg, nobx;              // input blocks / output blocks per CTA group
    int n_cbgroups, n_ngroups;
    int split;                  // accumulators per input block: the 16 rows of a plane are dealt to `split` issuers
    int n, d, h, w;
    int tiles_h, tiles_w, dchunks, dc, total_items;
    int balance;                // TcWalk
    uint32_t xplane_bytes, xslot_bytes, dyslot_bytes, tmem_cols, ns;
    int nsrc, src_cb[CTU_MAX_SRC], src_cboff[CTU_MAX_SRC];
};

__global__ void __launch_bounds__(WG_THREADS) conv3d_wgrad_tc2_kernel(const __grid_constant__ TcMaps xmaps,
                                                                      const __grid_constant__ CUtensorMap dymap, Wg2Params p) {
    constexpr int K = 3, PAD = 1;
    constexpr int WW = TC_TW + K - 1;
    constexpr uint32_t ROW = WW * 16;              // one x row of one channel block (w halo included)
    constexpr uint32_t DYBLK = TC_TW * 16;         // one channel block of one dy row
    constexpr int DYROWS = TC_TH + K - 1;
    const uint32_t NS = p.ns;
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_x = s_base;
    const uint32_t s_dy = s_x + NS * p.xslot_bytes;
    const uint32_t s_bar = s_dy + NS * p.dyslot_bytes;
    const uint32_t b_full = s_bar, b_empty = s_bar + 8 * TC_MAX_SLOTS, b_done = s_bar + 16 * TC_MAX_SLOTS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (b_done + 8 - s_base));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = blockIdx.y;
    const int cbg_i = grp % p.n_cbgroups, ng_i = grp / p.n_cbgroups;
    const int cb0 = cbg_i * p.cbg;
    const int ncb = (p.cb - cb0) < p.cbg ? (p.cb - cb0) : p.cbg;
    const int ob0 = ng_i * p.nobx;
    const int nob = (p.cob_n - ob0) < p.nobx ? (p.cob_n - ob0) : p.nobx;
    const int nobx = p.nobx;
    const uint32_t dyrow = (uint32_t)(K * nobx) * DYBLK;      // one dy row: 3 planes x all channel blocks
    const int ncol = K * K * nobx * 8;                        // MMA N = TMEM columns per accumulator

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < NS; ++i) {
            mbar_init(b_full + 8 * i, 1);
            mbar_init(b_empty + 8 * i, WG_ISSUERS);
        }
        mbar_init(b_done, WG_ISSUERS);
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
    if (warp == 0) {
        if (lane == 0) {
            Ring pr = {0, 0};
            const uint32_t bytes = (uint32_t)ncb * TC_TH * ROW + (uint32_t)DYROWS * dyrow;
            TcWalk walk;
            walk.init(p);
            int n, thi, twi, z0, nd;
            while (walk.next(p, n, thi, twi, z0, nd)) {
                const int h0 = thi * TC_TH, w0 = twi * TC_TW;
                for (int j = 0; j < nd; ++j, pr.next(NS)) {
                    mbar_wait(b_empty + 8 * pr.slot, pr.phase ^ 1);
                    mbar_expect_tx(b_full + 8 * pr.slot, bytes);
                    for (int b = 0; b < ncb; ++b) {
                        int s = 0;
#pragma unroll
                        for (int q = 1; q < CTU_MAX_SRC; ++q)
                            if (q < p.nsrc && cb0 + b >= p.src_cboff[q]) s = q;
                        tma_load_4d(s_x + pr.slot * p.xslot_bytes + b * p.xplane_bytes, &xmaps.m[s], (w0 - PAD) * 8, h0, z0 + j,
                                    n * p.src_cb[s] + (cb0 + b - p.src_cboff[s]), b_full + 8 * pr.slot);
                    }
                    // dy rows h0-1 .. h0+16, planes z-1 .. z+1, all output blocks of the group
                    tma_load_5d(s_dy + pr.slot * p.dyslot_bytes, &dymap, w0 * 8, ob0, z0 + j - PAD, h0 - PAD, n,
                                b_full + 8 * pr.slot);
                }
            }
        }
    } else {
        // D=f32, A=B=bf16, both MN-major (bits 15, 16), N at [17,23), M=64 at [24,29)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(ncol >> 3) << 17) |
                               (4u << 24);
        const uint32_t a_hi = 1u | (1u << 14);                    // SBO = 16 B: m-group g = lag g (kw)
        const uint32_t b_hi = (DYBLK >> 4) | (1u << 14);          // SBO = 256 B: next (kh row, kd plane, channel block)
        const uint32_t lbo = 8u << 16;                            // second 8-voxel core matrix: +128 B
        const int q = warp - 1;
        const int S = p.split;
        const int nacc = ncb * S;                                 // accumulator e = b * S + s
        const uint32_t leader = elect_one();
        uint32_t first = 1;
        Ring cons = {0, 0};
        TcWalk walk;
        walk.init(p);
        int n_, thi_, twi_, z0_, nd;
        while (walk.next(p, n_, thi_, twi_, z0_, nd)) {
            for (int j = 0; j < nd; ++j, cons.next(NS)) {
                mbar_wait(b_full + 8 * cons.slot, cons.phase);
                tc_fence_after();
                const uint32_t dy16 = ((s_dy + cons.slot * p.dyslot_bytes) >> 4) | lbo;
                for (int e = q; e < nacc; e += WG_ISSUERS) {
                    const int b = e / S, s = e - b * S;
                    const uint32_t x16 = ((s_x + cons.slot * p.xslot_bytes + b * p.xplane_bytes) >> 4) | lbo;
                    const uint32_t tcol = tmem_base + (uint32_t)e * ncol;
                    for (int r = s; r < TC_TH; r += S)   // x row r meets dy rows r-1 .. r+1 = box rows r .. r+2
                        umma_bf16_lead(leader, tcol, x16 + r * (ROW >> 4), a_hi, dy16 + r * (dyrow >> 4), b_hi, idesc,
                                       (r == s) ? (first ^ 1u) : 1u);
                }
                first = 0;
                umma_commit_lead(leader, b_empty + 8 * cons.slot);
            }
        }
        umma_commit_lead(leader, b_done);
        __syncwarp();
        // flush: M=64 accumulators sit in lanes 0-15 of every 32-lane quarter (row = quarter*16 + lane)
        const int quarter = warp & 3;
        mbar_wait(b_done, 0);
        tc_fence_after();
        const int row = quarter * 16 + (lane & 15);
        const int lag = row >> 3, ci = row & 7;
        const bool useful = lane < 16 && lag < K;
        constexpr int taps = K * K * K;
        for (int a = 0; a < nacc; ++a) {
            const int b = a / S;
            for (int g = 0; g < ncol / 8; ++g) {           // n-group g = (box row kh', box plane i, output block)
                const int khp = g / (K * nobx), i = (g / nobx) % K, ob = g % nobx;
                if (ob >= nob) continue;
                // dy row = x row + kh' - 1 = x row - kh + 1  ->  kh = 2 - kh';  dy plane = x plane + i - 1  ->  kd = 2 - i
                const int tap = ((K - 1 - i) * K + (K - 1 - khp)) * K + lag;
                float v[8];
                tmem_ld8(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)a * ncol + g * 8, v);
                if (useful) {
                    float* dst = p.dwp + ((((long long)(ob0 + ob) * p.cb + cb0 + b) * taps + tap) * 64 + ci * 8);
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (v[c] != 0.f) atomicAdd(dst + c, v[c]);
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

struct Wg2Geom {
    int cb, cob_n, cbg, nobx, n_cbgroups, n_ngroups, split;
    uint32_t xplane_bytes, xslot_bytes, dyslot_bytes, tmem_cols, ns;
    size_t smem;
};

// dense 3x3x3 only; N = 72 * nobx <= 216, all accumulators of a CTA in <= 512 TMEM columns
static bool wg2_geometry(int k, int cb, int cout, int h, int w, Wg2Geom& g) {
    if (k != 3 || h % TC_TH || w % TC_TW || cb < 1 || cout < 1) return false;
    g.cb = cb;
    g.cob_n = (cout + 7) / 8;
    g.nobx = g.cob_n < 3 ? g.cob_n : (g.cob_n % 3 == 0 ? 3 : 2);
    const int ncol = 72 * g.nobx;
    // input blocks per CTA and row split: aim at 4 busy issuers in <= 256 columns (two CTAs per SM), never above 512
    g.cbg = cb;
    while (g.cbg > 1 && g.cbg * ncol > 512) g.cbg = (g.cbg + 1) / 2;
    if (g.cbg * ncol > 512) return false;
    {
        const int ngr = (cb + g.cbg - 1) / g.cbg;
        g.cbg = (cb + ngr - 1) / ngr;
    }
    // Measured on B200 (scripts/bench_wgrad.py, 7->7 at 4x128^3): one issuer per CTA and two CTAs per SM with a 6-deep
    // ring (117 us) beats two issuers alternating rows into two accumulators (139 us) and one deep CTA per SM (145 us).
    g.split = 1;
    if (const char* e = getenv("CTU_WG2_SPLIT")) g.split = atoi(e);       // (tuning knobs, scripts/bench_wgrad.py)
    const int ww = TC_TW + k - 1;
    g.xplane_bytes = ((uint32_t)TC_TH * ww * 16 + 127u) & ~127u;
    g.xslot_bytes = g.xplane_bytes * g.cbg;
    g.dyslot_bytes = (uint32_t)(TC_TH + k - 1) * k * g.nobx * TC_TW * 16;
    const size_t fixed = 8 * (2 * TC_MAX_SLOTS + 3) + 16 + 1024;
    const uint32_t cols = (uint32_t)g.cbg * g.split * ncol;
    g.tmem_cols = 32;
    while (g.tmem_cols < cols) g.tmem_cols *= 2;
    if (g.tmem_cols > 512) return false;
    const size_t budget = g.tmem_cols > 256 ? 200 * 1024 : 112 * 1024;
    g.ns = 2;
    while (g.ns < 6 && fixed + (size_t)(g.ns + 1) * (g.xslot_bytes + g.dyslot_bytes) <= budget) ++g.ns;
    if (const char* e = getenv("CTU_WG2_NS")) g.ns = (uint32_t)atoi(e);
    g.smem = fixed + (size_t)g.ns * (g.xslot_bytes + g.dyslot_bytes);
    if (g.smem > 220 * 1024) return false;
    g.n_cbgroups = (cb + g.cbg - 1) / g.cbg;
    g.n_ngroups = (g.cob_n + g.nobx - 1) / g.nobx;
    return true;
}

static int conv3d_wgrad_tc2(const void* const* h_srcs, const int* h_src_channels, int nsrc, const void* dy, float* dwp,
                            int cout, int n, int d, int h, int w, const Wg2Geom& g, cudaStream_t stream) {
    TcMaps xmaps;
    CUtensorMap dymap;
    for (int i = 0; i < CTU_MAX_SRC; ++i) {       // x boxes: 16 rows x 18 voxels (w halo only)
        const int j = i < nsrc ? i : 0;
        int rc = make_map(&xmaps.m[i], h_srcs[j], n * ((h_src_channels[j] + 7) / 8), d, h, w, TC_TW + 2, TC_TH);
        if (rc != CTU_OK) return rc;
    }
    {
        // dy as (w*8, cob, d, h, n): the box [18 rows][3 planes][nobx][16 voxels] lands in shared memory as [h][d][cob][w][8]
        auto encode = get_encode();
        if (!encode) {
            set_error("cuTensorMapEncodeTiled entry point not found");
            return CTU_ERR_UNSUPPORTED;
        }
        const cuuint64_t plane_b = (cuuint64_t)d * h * w * 16;
        const cuuint64_t gdim[5] = {(cuuint64_t)w * 8, (cuuint64_t)g.cob_n, (cuuint64_t)d, (cuuint64_t)h, (cuuint64_t)n};
        const cuuint64_t gstr[4] = {plane_b, (cuuint64_t)h * w * 16, (cuuint64_t)w * 16, plane_b * g.cob_n};
        const cuuint32_t box[5] = {(cuuint32_t)TC_TW * 8, (cuuint32_t)g.nobx, 3, (cuuint32_t)(TC_TH + 2), 1};
        const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult cr = encode(&dymap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(dy), gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled(dy, wgrad2) failed (%d)", (int)cr);
            return CTU_ERR_INVALID;
        }
    }
    Wg2Params p = {};
    p.dwp = dwp;
    p.cb = g.cb; p.cob_n = g.cob_n; p.cbg = g.cbg; p.nobx = g.nobx; p.n_cbgroups = g.n_cbgroups; p.n_ngroups = g.n_ngroups;
    p.split = g.split;
    p.n = n; p.d = d; p.h = h; p.w = w;
    p.tiles_h = h / TC_TH; p.tiles_w = w / TC_TW;
    p.xplane_bytes = g.xplane_bytes; p.xslot_bytes = g.xslot_bytes; p.dyslot_bytes = g.dyslot_bytes; p.tmem_cols = g.tmem_cols;
    p.ns = g.ns;
    p.nsrc = nsrc;
    for (int i = 0, off = 0; i < CTU_MAX_SRC; ++i) {
        p.src_cb[i] = i < nsrc ? (h_src_channels[i] + 7) / 8 : 0;
        p.src_cboff[i] = off;
        off += p.src_cb[i];
    }
    int ctas_per_sm = (int)((227 * 1024) / (g.smem + 1024));
    if (ctas_per_sm > 512 / (int)g.tmem_cols) ctas_per_sm = 512 / (int)g.tmem_cols;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm > 4) ctas_per_sm = 4;
    if (const char* e = getenv("CTU_WG2_CTAS")) ctas_per_sm = atoi(e);
    const int groups = g.n_cbgroups * g.n_ngroups;
    int gx = (148 * ctas_per_sm + groups - 1) / groups;
    // d-chunks cost nothing here (no halo planes): cut until the resident CTAs get >= 4 items each, chunks of >= 4 planes
    const int tiles = n * p.tiles_h * p.tiles_w;
    int dc = d;
    while (dc > 32 && (long long)tiles * ((d + dc - 1) / dc) < 4LL * gx) dc = (dc + 1) / 2;
    while (dc > 4 && (long long)tiles * ((d + dc - 1) / dc) < 2LL * gx) dc = (dc + 1) / 2;
    if (const char* e = getenv("CTU_WG2_DC")) dc = atoi(e);
    p.dc = dc;
    p.dchunks = (d + dc - 1) / dc;
    p.total_items = tiles * p.dchunks;
    p.balance = tc_balance(false);
    if (gx > p.total_items) gx = p.total_items;     // (every CTA ends with an atomic flush of its accumulators: no more CTAs than chunks)
    if (gx < 1) gx = 1;
    cudaError_t e = cudaFuncSetAttribute(conv3d_wgrad_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    if (e != cudaSuccess) {
        set_error("conv3d wgrad2 tensor path: smem %zu: %s", g.smem, cudaGetErrorString(e));
        return (int)e;
    }
    conv3d_wgrad_tc2_kernel<<<dim3(gx, groups), WG_THREADS, g.smem, stream>>>(xmaps, dymap, p);
    return check_launch("ctu_conv3d_wgrad(tcgen05, kd+kh in N)");
}

// sum of dy over voxels per channel (bias gradient of convolutions that carry a bias: the legacy 5^3 family)
__global__ void tc_channel_sum_kernel(const __nv_bfloat16* __restrict__ dy, float* __restrict__ dbias, int cout, int cob_n,
                                      long long spatial) {
    const int ob = blockIdx.y, n = blockIdx.z;
    const __nv_bfloat16* base = dy + ((long long)n * cob_n + ob) * spatial * 8;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; s + 7 * stride < spatial; s += 8 * stride) {          // eight 16-byte loads in flight per thread
        Raw8<__nv_bfloat16> r[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) r[u].load(base + (s + u * stride) * 8);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const V8 v = r[u].get();
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] += v.v[c];
        }
    }
    for (; s < spatial; s += stride) {
        V8 v = Vec8<__nv_bfloat16>::load(base + s * 8);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] += v.v[c];
    }
    // one atomic per channel and BLOCK (same-address atomics serialise at ~1 ns each)
    __shared__ float red[8][8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float t = warp_sum(acc[c]);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][c] = t;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        float t = 0.f;
        for (int wq = 0; wq < (int)(blockDim.x >> 5); ++wq) t += red[wq][threadIdx.x];
        const int ch = ob * 8 + threadIdx.x;
        if (ch < cout && t != 0.f) atomicAdd(dbias + ch, t);
    }
}

struct WgGeom {
    int cb, cob_n, cbg, ng, n_cbgroups, n_ngroups, sparse, pblk;
    uint32_t plane_bytes, xslot_bytes, dyslot_bytes, tmem_cols, ns, nds;
    size_t smem;
};

// phase_cout: 0, or the natural channel count of a phase-major dy (fused up-sampling stage: cout = 8 phases x 8P)
static bool wg_geometry(int k, int cb, int cout, int h, int w, WgGeom& g, int phase_cout = 0) {
    if (k != 3 && k != 5) return false;
    if (h % TC_TH || w % TC_TW || cb < 1 || cout < 1) return false;
    g.cb = cb;
    g.cob_n = (cout + 7) / 8;
    const int k2 = k * k;
    const int hh = TC_TH + k - 1, ww = TC_TW + k - 1;
    g.plane_bytes = ((uint32_t)hh * ww * 16 + 127u) & ~127u;
    g.sparse = 0;
    g.pblk = 0;
    int cols_per_block;       // TMEM columns per input block
    if (phase_cout > 0 && k == 3 && cout == 8 * ((phase_cout + 7) / 8 * 8)) {
        // phase-sparse: group = one qd (N = 64P) while that fits 128 columns, else one (qd, qh) (N = 32P)
        g.pblk = (phase_cout + 7) / 8;
        g.sparse = 64 * g.pblk <= 128 ? 1 : 2;
        if (g.sparse == 2 && 32 * g.pblk > 256) return false;
        g.ng = (g.sparse == 1 ? 4 : 2) * g.pblk * 8;
        cols_per_block = 2 * (g.sparse == 1 ? 64 : 32) * g.pblk;
    } else {
        // choose (cbg, ng): k2*cbg*ng <= 512 TMEM columns; prefer all output channels, then as many input blocks as fit
        g.ng = g.cob_n * 8;
        while (k2 * g.ng > 512 || k * g.ng > 256) g.ng = ((g.ng / 8 + 1) / 2) * 8;
        cols_per_block = k2 * g.ng;
    }
    g.cbg = 512 / cols_per_block;
    if (g.cbg > g.cb) g.cbg = g.cb;
    if (g.cbg * (g.sparse ? 2 : k) > WG_ISSUERS * WG_MAX_OWN) g.cbg = WG_ISSUERS * WG_MAX_OWN / (g.sparse ? 2 : k);
    // shared memory: x ring (k+1 slots x cbg planes) + dy ring
    const size_t fixed = 8 * (4 * TC_MAX_SLOTS + 3) + 16 + WG_ISSUERS * WG_MAX_OWN * 8 + 1024;
    for (;;) {
        g.xslot_bytes = g.plane_bytes * g.cbg;
        g.dyslot_bytes = (uint32_t)(g.ng / 8) * (TC_TH + 2 * (k - 1)) * TC_TW * 16;   // + zero rows above / below
        g.smem = fixed + (size_t)(k + 1) * g.xslot_bytes + 2 * (size_t)g.dyslot_bytes;
        if (g.smem <= 200 * 1024 || g.cbg == 1) break;
        --g.cbg;
    }
    if (g.smem > 220 * 1024) return false;
    // balance the input-block groups (5 blocks as 3 + 2, not 4 + 1): every group gets the same number of CTAs, so the
    // largest group sets the kernel's duration
    {
        const int ngr = (g.cb + g.cbg - 1) / g.cbg;
        g.cbg = (g.cb + ngr - 1) / ngr;
        g.xslot_bytes = g.plane_bytes * g.cbg;
        g.smem = fixed + (size_t)(k + 1) * g.xslot_bytes + 2 * (size_t)g.dyslot_bytes;
    }
    // deepen both rings while a ~100 KB budget (two CTAs per SM) allows: more TMA loads in flight
    g.ns = k + 1;
    g.nds = 2;
    const size_t wg_budget = (uint32_t)cols_per_block * g.cbg > 256 ? 200 * 1024 : 100 * 1024;   // > 256 TMEM columns: one CTA per SM
    while (g.ns < TC_MAX_SLOTS && g.smem + g.xslot_bytes + g.dyslot_bytes <= wg_budget) {
        ++g.ns;
        if (g.nds < TC_MAX_SLOTS) ++g.nds;
        g.smem += g.xslot_bytes + g.dyslot_bytes;
    }
    g.n_cbgroups = (g.cb + g.cbg - 1) / g.cbg;
    g.n_ngroups = (g.cob_n * 8 + g.ng - 1) / g.ng;
    uint32_t cols = (uint32_t)cols_per_block * g.cbg;
    g.tmem_cols = 32;
    while (g.tmem_cols < cols) g.tmem_cols *= 2;
    return g.tmem_cols <= 512;
}

int conv3d_wgrad_tc(const void* const* h_srcs, const int* h_src_channels, int nsrc, const void* dy, float* dwp,
                    float* dbias, int phase_cout, int cout, int k, int n, int d, int h, int w, cudaStream_t stream) {
    WgGeom g;
    const int cbt = total_blocks(nsrc, h_src_channels);
    if (cbt < 0 || !wg_geometry(k, cbt, cout, h, w, g, phase_cout)) {
        set_error("conv3d wgrad tensor path: shape k=%d blocks=%d cout=%d %dx%dx%d not covered", k, cbt, cout, d, h, w);
        return CTU_ERR_UNSUPPORTED;
    }
    {
        // narrow dense 3x3x3 layers: the (kd, kh)-in-N formulation (a third of the MMAs); CTU_WGRAD_V1=1 keeps the first one
        static const bool v1 = getenv("CTU_WGRAD_V1") != nullptr;
        Wg2Geom g2;
        // (measured: a gain for one input and one output block -- 7->7 180 -> 117 us, 2->7 178 -> 115 us at 4x128^3 -- and
        // none once N = 72 * blocks exceeds ~100: 14->14 at 4x64^3 81 vs 82 us; CTU_WGRAD_V2_ALL=1 routes every covered shape)
        static const bool all2 = getenv("CTU_WGRAD_V2_ALL") != nullptr;
        if (!v1 && !g.sparse && k == 3 && (all2 ? (cbt <= 4 && cout <= 32) : (cbt == 1 && cout <= 8)) &&
            wg2_geometry(k, cbt, cout, h, w, g2)) {
            int rc2 = conv3d_wgrad_tc2(h_srcs, h_src_channels, nsrc, dy, dwp, cout, n, d, h, w, g2, stream);
            if (rc2 == CTU_OK && dbias != nullptr) rc2 = channel_sum_bias(dy, dbias, cout, g2.cob_n, n, (long long)d * h * w, stream);
            return rc2;
        }
    }
    TcMaps xmaps;
    CUtensorMap dymap;
    int rc = make_src_maps(xmaps, h_srcs, h_src_channels, nsrc, n, d, h, w, k);
    if (rc == CTU_OK) {
        // dy as (w*8, cob, h, d, n): one box [TH][ng/8][TW*8] lands in shared memory as [h][cob][w][8]
        auto encode = get_encode();
        const cuuint64_t plane_b = (cuuint64_t)d * h * w * 16;
        const cuuint64_t gdim[5] = {(cuuint64_t)w * 8, (cuuint64_t)g.cob_n, (cuuint64_t)h, (cuuint64_t)d, (cuuint64_t)n};
        const cuuint64_t gstr[4] = {plane_b, (cuuint64_t)w * 16, (cuuint64_t)h * w * 16, plane_b * g.cob_n};
        const cuuint32_t box[5] = {(cuuint32_t)TC_TW * 8, (cuuint32_t)(g.ng / 8), (cuuint32_t)TC_TH, 1, 1};
        const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult cr = encode(&dymap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(dy), gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled(dy) failed (%d)", (int)cr);
            rc = CTU_ERR_INVALID;
        }
    }
    if (rc != CTU_OK) return rc;
    WgParams p = {};
    p.dwp = dwp;
    p.cb = g.cb; p.cob_n = g.cob_n; p.cbg = g.cbg; p.ng = g.ng; p.n_cbgroups = g.n_cbgroups; p.n_ngroups = g.n_ngroups;
    p.n = n; p.d = d; p.h = h; p.w = w;
    p.tiles_h = h / TC_TH; p.tiles_w = w / TC_TW;
    const int tiles = n * p.tiles_h * p.tiles_w;
    const int groups = g.n_cbgroups * g.n_ngroups;
    int dc = d;
    while (dc > 8 && (long long)tiles * ((d + dc - 1) / dc) * groups < 148 * 4) dc = (dc + 1) / 2;
    p.dc = dc;
    p.dchunks = (d + dc - 1) / dc;
    p.total_items = tiles * p.dchunks;
    p.plane_bytes = g.plane_bytes; p.xslot_bytes = g.xslot_bytes; p.dyslot_bytes = g.dyslot_bytes; p.tmem_cols = g.tmem_cols;
    p.ns = g.ns; p.nds = g.nds;
    p.sparse = g.sparse; p.pblk = g.pblk;
    p.nsrc = nsrc;
    for (int i = 0, off = 0; i < CTU_MAX_SRC; ++i) {
        p.src_cb[i] = i < nsrc ? (h_src_channels[i] + 7) / 8 : 0;
        p.src_cboff[i] = off;
        off += p.src_cb[i];
    }
    int ctas_per_sm = (int)((227 * 1024) / (g.smem + 1024));
    if (ctas_per_sm > 512 / (int)g.tmem_cols) ctas_per_sm = 512 / (int)g.tmem_cols;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    int gx = (148 * ctas_per_sm + groups - 1) / groups;
    p.balance = tc_balance(true);
    if (gx > p.total_items) gx = p.total_items;     // (every CTA ends with an atomic flush of its accumulators: no more CTAs than chunks)
    if (gx < 1) gx = 1;
    dim3 grid(gx, groups);
    auto go = [&](auto kern) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
        if (e != cudaSuccess) {
            set_error("conv3d wgrad tensor path: smem %zu: %s", g.smem, cudaGetErrorString(e));
            return (int)e;
        }
        kern<<<grid, WG_THREADS, g.smem, stream>>>(xmaps, dymap, p);
        return check_launch("ctu_conv3d_wgrad(tcgen05)");
    };
    rc = k == 3 ? go(conv3d_wgrad_tc_kernel<3>) : go(conv3d_wgrad_tc_kernel<5>);
    if (rc == CTU_OK && dbias != nullptr) rc = channel_sum_bias(dy, dbias, cout, g.cob_n, n, (long long)d * h * w, stream);
    return rc;
}

int channel_sum_bias(const void* dy, float* dbias, int cout, int cob_n, int n, long long spatial, cudaStream_t stream) {
    dim3 bgrid((unsigned)((spatial + 256 * 32 - 1) / (256 * 32)), cob_n, n);
    tc_channel_sum_kernel<<<bgrid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(dy), dbias, cout, cob_n, spatial);
    return check_launch("ctu_conv3d_wgrad(dbias)");
}

}  // namespace ctu

using namespace ctu;

extern "C" {

int ctu_has_tensor_path(void) { return 1; }

int ctu_conv_tc_supported(int k, int nsrc, const int* h_src_channels, int cout, int d, int h, int w) {
    (void)d;
    TcGeom g;
    const int cb = total_blocks(nsrc, h_src_channels);
    return (cb > 0 && tc_geometry(k, cb, cout, h, w, g)) ? 1 : 0;
}

int ctu_conv_tc_bnred_supported(int k, int nsrc, const int* h_src_channels, int cout, int d, int h, int w) {
    (void)d;
    TcGeom g;
    const int cb = total_blocks(nsrc, h_src_channels);
    return (k == 3 && cb > 0 && tc_geometry(k, cb, cout, h, w, g) && g.cobg <= 2) ? 1 : 0;
}

int ctu_conv_tc_wgrad_supported(int k, int nsrc, const int* h_src_channels, int cout, int d, int h, int w) {
    (void)d;
    WgGeom g;
    const int cb = total_blocks(nsrc, h_src_channels);
    return (cb > 0 && wg_geometry(k, cb, cout, h, w, g)) ? 1 : 0;     // (the phase-sparse mode covers a superset)
}

long long ctu_conv_tc_wimg_bytes(int k, int nsrc, const int* h_src_channels, int cout) {
    TcGeom g;
    const int cb = total_blocks(nsrc, h_src_channels);
    if (cb < 0 || !tc_geometry(k, cb, cout, TC_TH, TC_TW, g)) return -1;
    return (long long)g.wimg_bytes * g.ngroups;
}

int ctu_conv_tc_pack_weight(const float* wp, void* wimg, int k, int nsrc, const int* h_src_channels, int cout,
                            ctu_stream stream) {
    TcGeom g;
    const int cb = total_blocks(nsrc, h_src_channels);
    CTU_REQUIRE(wp && wimg && cb > 0 && tc_geometry(k, cb, cout, TC_TH, TC_TW, g), "ctu_conv_tc_pack_weight: bad arguments");
    const long long total = (long long)g.wimg_bytes / 2 * g.ngroups;
    tc_pack_wimg_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(wp, reinterpret_cast<__nv_bfloat16*>(wimg), k, g.cb, g.cbg, g.cob_n, g.cobg, g.nt, g.nm, total);
    return check_launch("ctu_conv_tc_pack_weight");
}

}  // extern "C"
