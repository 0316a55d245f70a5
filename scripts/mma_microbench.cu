// Micro-benchmark of tcgen05.mma issue behaviour on sm_100a for the SMALL shapes the U-Net convolutions use
// (M in {64,128}, N in {24..256}, K = 16, bf16, no-swizzle K-major operands in shared memory).
// Question it answers: how many cycles does one MMA cost as a function of
//   W = number of issuing warps, A = independent TMEM accumulators each warp cycles through, and the shape,
// i.e. are back-to-back MMAs limited by the tensor pipe, by the shared-memory operand fetch, or by a per-MMA latency
// that only independent chains can hide.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/mma_microbench scripts/mma_microbench.cu
//   ./scripts/mma_microbench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void mma(uint32_t leader, uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                    uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %7, 0;\n\tsetp.ne.b32 q, %0, 0;\n\t"
        "mov.b64 da, {%2, %3};\n\tmov.b64 db, {%4, %5};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], da, db, %6, p;\n\t}"
        ::"r"(leader), "r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void commit(uint32_t leader, uint32_t bar) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%1];\n\t}" ::"r"(leader), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}

// W issuing warps; warp w cycles through A accumulators; L MMAs per warp.  mn_major: both operands MN-major (wgrad).
__global__ void bench(int M, int N, int W, int A, int L, int mn_major, long long* out, int sbo16 = 16, int astep16 = 2,
                      int lbo16 = 8, int tcols = 512) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bars[8];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 64 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(tcols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (warp < W) {
        const uint32_t leader = elect_one();
        uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        if (mn_major) idesc |= (1u << 15) | (1u << 16);
        const uint32_t base = smem_u32(smem) >> 4;
        // K-major: SBO = 256 B between 8-row groups, LBO = 128 B; MN-major: SBO = 16 B (lag groups), LBO = 128 B
        const uint32_t ahi = (mn_major ? 1u : (uint32_t)sbo16) | (1u << 14), bhi = 16u | (1u << 14);
        const uint32_t alo = (base + warp * 64) | ((uint32_t)lbo16 << 16), blo = (base + 2048) | (8u << 16);
        long long t0 = clock64();
        for (int i = 0; i < L; ++i) {
            const uint32_t d = tmem + (uint32_t)((warp * A + (i % A)) * N);
            mma(leader, d, alo + (i & 7) * astep16, ahi, blo + (i & 3) * 2, bhi, idesc, i >= A ? 1u : 0u);
        }
        commit(leader, smem_u32(&bars[warp]));
        mbar_wait(smem_u32(&bars[warp]), 0);
        long long t1 = clock64();
        if (lane == 0 && blockIdx.x == 0) out[warp] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tcols) : "memory");
}

// Commit cost: every iteration issues `nm` MMAs (M=128, N=32) and `nc` tcgen05.commit onto mbarriers with a huge
// arrival count (never waited on inside the loop).
__global__ void bench_commit(int W, int nm, int nc, int L, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bars[8];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 64 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(i < 4 ? 1 : 1000000) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (warp < W) {
        const uint32_t leader = elect_one();
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t base = smem_u32(smem) >> 4;
        const uint32_t ahi = 16u | (1u << 14), bhi = 16u | (1u << 14);
        const uint32_t alo = (base + warp * 64) | (8u << 16), blo = (base + 2048) | (8u << 16);
        long long t0 = clock64();
        for (int i = 0; i < L; ++i) {
            for (int m = 0; m < nm; ++m) mma(leader, tmem + warp * 32, alo + m * 2, ahi, blo, bhi, idesc, (i | m) ? 1u : 0u);
            for (int c = 0; c < nc; ++c) commit(leader, smem_u32(&bars[4 + ((warp + c) & 3)]));
        }
        commit(leader, smem_u32(&bars[warp]));
        mbar_wait(smem_u32(&bars[warp]), 0);
        long long t1 = clock64();
        if (lane == 0 && blockIdx.x == 0) out[warp] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    {
        // two CTAs per SM (grid = 2 x 148, 256 TMEM columns each) against one CTA with the same number of issuing warps
        {
            long long* d_q;
            cudaMalloc(&d_q, 8 * sizeof(long long));
            cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            printf("CTAs per SM x issuing warps per CTA (M=128, N=32): cycles per MMA per warp (CTA 0)\n");
            for (int ctas = 1; ctas <= 2; ++ctas)
                for (int W = 1; W <= 4; W *= 2) {
                    bench<<<148 * ctas, 128, 64 * 1024>>>(128, 32, W, 2, 4096, 0, d_q, 16, 2, 8, 256);
                    if (cudaDeviceSynchronize() != cudaSuccess) { printf("cta sweep failed\n"); return 1; }
                    long long h[8];
                    cudaMemcpy(h, d_q, sizeof(h), cudaMemcpyDeviceToHost);
                    long long mx = 0;
                    for (int w = 0; w < W; ++w) mx = h[w] > mx ? h[w] : mx;
                    printf("ctas/SM=%d W=%d : %6.1f per warp, %6.1f aggregate per SM\n", ctas, W, (double)mx / 4096.0, (double)mx / (4096.0 * W * ctas));
                }
            cudaFree(d_q);
        }
        // A-operand layout sweep (K-major, M=128, N=32): SBO = pitch between 8-voxel row groups, start address step, LBO
        long long* d_o;
        cudaMalloc(&d_o, 8 * sizeof(long long));
        cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        printf("A-operand layout sweep: M=128 N=32 K-major; SBO bytes, start-address step bytes, LBO bytes -> aggregate cycles per MMA\n");
        const int sbos[] = {16, 18};
        const int steps[] = {1};
        const int lbos[] = {8, 328};
        for (int W = 2; W <= 4; W *= 2)
            for (int sb : sbos)
                for (int st : steps)
                    for (int lb : lbos) {
                        bench<<<1, 128, 64 * 1024>>>(128, 32, W, 2, 4096, 0, d_o, sb, st, lb);
                        if (cudaDeviceSynchronize() != cudaSuccess) { printf("layout sweep failed\n"); return 1; }
                        long long h[8];
                        cudaMemcpy(h, d_o, sizeof(h), cudaMemcpyDeviceToHost);
                        long long mx = 0;
                        for (int w = 0; w < W; ++w) mx = h[w] > mx ? h[w] : mx;
                        printf("W=%d SBO=%4d step=%4d LBO=%5d : %6.1f\n", W, sb * 16, st * 16, lb * 16, (double)mx / (4096.0 * W));
                    }
        cudaFree(d_o);
    }
    {
        long long* d_o;
        cudaMalloc(&d_o, 8 * sizeof(long long));
        cudaFuncSetAttribute(bench_commit, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        printf("commit cost: W warps, nm MMAs (M=128,N=32) + nc commits per iteration -> cycles per iteration (max over warps)\n");
        for (int W = 1; W <= 4; W *= 2)
            for (int nm = 0; nm <= 5; nm += 5)
                for (int nc = 0; nc <= 2; ++nc) {
                    if (nm == 0 && nc == 0) continue;
                    bench_commit<<<1, 128, 64 * 1024>>>(W, nm, nc, 2000, d_o);
                    if (cudaDeviceSynchronize() != cudaSuccess) { printf("commit bench failed\n"); return 1; }
                    long long h[8];
                    cudaMemcpy(h, d_o, sizeof(h), cudaMemcpyDeviceToHost);
                    long long mx = 0;
                    for (int w = 0; w < W; ++w) mx = h[w] > mx ? h[w] : mx;
                    printf("W=%d nm=%d nc=%d : %8.1f cycles/iter\n", W, nm, nc, (double)mx / 2000);
                }
    }
    long long* d_out;
    cudaMalloc(&d_out, 8 * sizeof(long long));
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int L = 4096;
    printf("%-4s %-4s %-3s %-3s %-3s | cycles per MMA per warp | aggregate cycles per MMA | tensor floor max(M,128)*N/256\n", "M", "N",
           "W", "A", "MN");
    const int Ms[] = {128, 64};
    const int Ns[] = {32, 40, 48, 64, 80, 96, 128, 256};
    for (int mn = 0; mn < 2; ++mn)
        for (int M : Ms)
            for (int N : Ns)
                for (int W = 1; W <= 4; W *= 2)
                    for (int A = 1; A <= 4; A *= 2) {
                        if (W * A * N > 512) continue;
                        if (mn && M != 64) continue;
                        bench<<<1, 32 * 4, 64 * 1024>>>(M, N, W, A, L, mn, d_out);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) {
                            printf("M=%d N=%d W=%d A=%d mn=%d: %s\n", M, N, W, A, mn, cudaGetErrorString(e));
                            return 1;
                        }
                        long long h[8];
                        cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
                        long long mx = 0;
                        for (int w = 0; w < W; ++w) mx = h[w] > mx ? h[w] : mx;
                        printf("%-4d %-4d %-3d %-3d %-3d | %8.1f | %8.1f | %d\n", M, N, W, A, mn, (double)mx / L, (double)mx / (L * W),
                               (M > 128 ? M : 128) * N / 256);
                    }
    return 0;
}
