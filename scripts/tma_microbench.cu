// How fast does TMA stage the halo planes of the convolution kernels?  A producer thread per CTA walks the (16 x 16 tile,
// plane) sequence of a blocked bf16 activation [4][128][128][128][8] (134 MB) exactly as conv3d_tc_kernel does and loads one
// box per plane into a ring of shared-memory slots; a consumer warp only waits for the slot and hands it back.  Variants:
//   box 18 x 18 voxels at (w0 - 1, h0 - 1)       -- what the kernels use (288-byte rows, 16-byte aligned)
//   box 32 x 18 voxels at (w0 - 8, h0 - 1)       -- 512-byte rows on 128-byte boundaries (1.78 x the bytes)
//   box 24 x 18 voxels at (w0 - 4, h0 - 1)       -- 384-byte rows on 64-byte boundaries
//   two planes per box (box depth 2)
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/_bin/tma_microbench scripts/tma_microbench.cu -lcuda
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}

constexpr int D = 128, H = 128, W = 128, NB = 4, NS_MAX = 16;

__global__ void __launch_bounds__(64) stream_kernel(const __grid_constant__ CUtensorMap map, int box_w, int box_h, int box_d, int woff,
                                                   int ns, uint32_t slot_bytes, int blocks_per_plane, int par) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bars[2 * NS_MAX];
    const uint32_t s_base = smem_u32(smem), b_full = smem_u32(bars), b_empty = b_full + 8 * NS_MAX;
    if (threadIdx.x == 0) {
        for (int i = 0; i < ns; ++i) {
            mbar_init(b_full + 8 * i, 1);
            mbar_init(b_empty + 8 * i, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int tiles = (H / 16) * (W / 16);
    const long long total = (long long)NB * tiles * (D / box_d);                  // (block, tile, plane group) sequence
    const long long p0 = total * blockIdx.x / gridDim.x, p1 = total * (blockIdx.x + 1) / gridDim.x;
    const uint32_t bbytes = (uint32_t)box_w * box_h * box_d * 16, bstride = (bbytes + 127u) & ~127u;
    const uint32_t bytes = bbytes * blocks_per_plane;
    if (threadIdx.x < 32) {
        // par = 0: lane 0 issues every box of a slot; par = 1: lane b issues box b (the whole warp walks the sequence)
        if (!par && threadIdx.x != 0) return;
        uint32_t slot = 0, phase = 0;
        for (long long p = p0; p < p1; ++p) {
            const int zg = (int)(p % (D / box_d));
            const int col = (int)(p / (D / box_d));
            const int t = col % tiles, nb = col / tiles;
            const int h0 = (t / (W / 16)) * 16, w0 = (t % (W / 16)) * 16;
            mbar_wait(b_empty + 8 * slot, phase ^ 1);
            if (threadIdx.x == 0) mbar_expect_tx(b_full + 8 * slot, bytes);
            if (par) {
                __syncwarp();
                const int b = threadIdx.x;
                if (b < blocks_per_plane)
                    tma_load_4d(s_base + slot * slot_bytes + b * bstride, &map, (w0 - woff) * 8, h0 - 1, zg * box_d - 1, (nb + b) % NB,
                                b_full + 8 * slot);
            } else {
                for (int b = 0; b < blocks_per_plane; ++b)
                    tma_load_4d(s_base + slot * slot_bytes + b * bstride, &map, (w0 - woff) * 8, h0 - 1, zg * box_d - 1, (nb + b) % NB,
                                b_full + 8 * slot);
            }
            if (++slot == (uint32_t)ns) { slot = 0; phase ^= 1; }
        }
    } else if (threadIdx.x == 32) {
        uint32_t slot = 0, phase = 0;
        for (long long p = p0; p < p1; ++p) {
            mbar_wait(b_full + 8 * slot, phase);
            mbar_arrive(b_empty + 8 * slot);
            if (++slot == (uint32_t)ns) { slot = 0; phase ^= 1; }
        }
    }
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
}

int main() {
    void* buf;
    const size_t bytes = (size_t)NB * D * H * W * 16;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 0, bytes);
    void* flush;
    cudaMalloc(&flush, 512u << 20);
    auto encode = get_encode();
    printf("TMA staging of a [4][128][128][128][8] bf16 tensor (134 MB), 16x16 tiles, plane by plane, ring of slots, grid = CTAs/SM x 148\n");
    printf("%-28s %-6s %-5s %-5s | %8s %9s %10s\n", "box (w x h x d voxels)", "blocks", "CTAs", "slots", "us", "GB/s", "useful GB/s");
    struct Cfg { int bw, bh, bd, woff, blocks, ctas, ns, par; };
    const Cfg cfgs[] = {
        {18, 18, 1, 1, 1, 2, 16, 0}, {18, 18, 1, 1, 1, 1, 16, 0}, {18, 18, 1, 1, 1, 4, 8, 0}, {32, 18, 1, 8, 1, 2, 8, 0},
        {18, 18, 2, 1, 1, 2, 8, 0},  {16, 16, 1, 0, 1, 2, 16, 0}, {18, 18, 1, 1, 2, 1, 8, 0}, {18, 18, 1, 1, 2, 1, 8, 1},
        {18, 18, 1, 1, 4, 1, 4, 0},  {18, 18, 1, 1, 4, 1, 4, 1},  {18, 18, 1, 1, 4, 1, 8, 1}, {18, 18, 1, 1, 4, 2, 4, 1},
        {18, 18, 1, 1, 8, 1, 3, 0},  {18, 18, 1, 1, 8, 1, 3, 1},
    };
    for (const Cfg& c : cfgs) {
        CUtensorMap map;
        const cuuint64_t gdim[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)NB};
        const cuuint64_t gstr[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16};
        const cuuint32_t box[4] = {(cuuint32_t)c.bw * 8, (cuuint32_t)c.bh, (cuuint32_t)c.bd, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) { printf("encode failed %d\n", (int)cr); continue; }
        const uint32_t slot_bytes = (((uint32_t)c.bw * c.bh * c.bd * 16 + 127u) & ~127u) * c.blocks;
        const size_t smem = (size_t)slot_bytes * c.ns + 1024;
        cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaMemsetAsync(flush, rep, 512u << 20);
            cudaEvent_t a, b;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            cudaEventRecord(a);
            stream_kernel<<<148 * c.ctas, 64, smem>>>(map, c.bw, c.bh, c.bd, c.woff, c.ns, slot_bytes, c.blocks, c.par);
            cudaEventRecord(b);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            if (ms < best) best = ms;
        }
        const double tiles = (double)NB * 64 * (D / c.bd);
        const double moved = tiles * c.bw * c.bh * c.bd * 16.0 * c.blocks, useful = (double)bytes * c.blocks;   // (blocks > 1: every tile loads `blocks` boxes)
        char name[64];
        snprintf(name, sizeof name, "%d x %d x %d @w0-%d%s", c.bw, c.bh, c.bd, c.woff, c.par ? " lanes" : "");
        printf("%-28s %-6d %-5d %-5d | %8.1f %9.0f %10.0f\n", name, c.blocks, c.ctas, c.ns, best * 1e3, moved / best / 1e6, useful / best / 1e6);
    }
    return 0;
}
