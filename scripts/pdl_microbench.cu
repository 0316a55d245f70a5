// How much of a dependent kernel chain's time is launch gap on B200, and how much of it does programmatic dependent launch
// (griddepcontrol.wait / launch_dependents + cudaLaunchAttributeProgrammaticStreamSerialization) recover inside a CUDA graph?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/pdl_microbench scripts/pdl_microbench.cu && /tmp/pdl_microbench
#include <cuda_runtime.h>
#include <stdio.h>

__global__ void work(float* buf, int n, int iters, int pdl, int prologue) {
    // prologue that does not touch dependent data (stands for barrier init / TMEM alloc / weight image load)
    __shared__ float s[256];
    float w = threadIdx.x;
    for (int i = 0; i < prologue; ++i) w = w * 1.0001f + 0.5f;
    s[threadIdx.x] = w;
    __syncthreads();
    if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = i0; i < n; i += gridDim.x * blockDim.x) {
        float v = buf[i];
        for (int k = 0; k < iters; ++k) v = v * 1.0001f + 0.001f;
        buf[i] = v + s[(threadIdx.x + 1) & 255] * 1e-30f;
    }
}

static float run(float* buf, int n, int iters, int pdl, int prologue, int chain, int grid) {
    cudaStream_t st;
    cudaStreamCreate(&st);
    cudaGraph_t g;
    cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < chain; ++i) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(256);
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = pdl ? 1 : 0;
        cudaError_t e = cudaLaunchKernelEx(&cfg, work, buf, n, iters, pdl, prologue);
        if (e != cudaSuccess) printf("launch: %s\n", cudaGetErrorString(e));
    }
    cudaError_t e = cudaStreamEndCapture(st, &g);
    if (e != cudaSuccess) printf("capture: %s\n", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&ge, g, 0);
    if (e != cudaSuccess) printf("instantiate: %s\n", cudaGetErrorString(e));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaGraphLaunch(ge, st);
        cudaStreamSynchronize(st);
        cudaEventRecord(a, st);
        cudaGraphLaunch(ge, st);
        cudaEventRecord(b, st);
        cudaStreamSynchronize(st);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    cudaGraphExecDestroy(ge);
    cudaGraphDestroy(g);
    cudaStreamDestroy(st);
    return best * 1e3f / chain;
}

int main() {
    const int n = 1 << 24;
    float* buf;
    cudaMalloc(&buf, sizeof(float) * n);
    cudaMemset(buf, 0, sizeof(float) * n);
    printf("per-kernel us in a captured chain of 100 dependent launches (grid 296 x 256 threads)\n");
    printf("%8s %9s | %9s %9s %9s\n", "elements", "prologue", "plain", "pdl", "saved");
    const int sizes[] = {1 << 14, 1 << 20, 1 << 22, 1 << 24};
    const int pros[] = {0, 2000};
    for (int si = 0; si < 4; ++si)
        for (int pi = 0; pi < 2; ++pi) {
            const float t0 = run(buf, sizes[si], 8, 0, pros[pi], 100, 296);
            const float t1 = run(buf, sizes[si], 8, 1, pros[pi], 100, 296);
            printf("%8d %9d | %9.2f %9.2f %9.2f\n", sizes[si], pros[pi], t0, t1, t0 - t1);
        }
    cudaFree(buf);
    return 0;
}
