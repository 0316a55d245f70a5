// Data-parallel gradient exchange fused with the optimizer over NVLink / NVSwitch peer memory.
//
// The reference averages gradients inside nn.DataParallel (ctunet/pytorch/Model.py:481-486: reduce to GPU 0, step, re-broadcast
// every iteration).  One process per GPU, the flat gradient buffer of every rank lives in a cudaMalloc'ed, IPC-shared
// allocation; the optimizer kernel of rank r reads element f from ALL ranks' buffers through peer pointers (plain ld.global
// over NVLink -- NVSwitch gives every pair full bandwidth), sums them in rank order (so every rank computes bit-identical
// averages and the replicas never drift), scales by 1/world and applies the update in the same pass: all-reduce and
// optimizer are ONE kernel, there is no NCCL call in the step and the whole iteration stays one CUDA graph.
//
// Synchronisation (flags live in the same shared allocation, one 32-bit slot per peer, system-scope release / acquire):
//   arrive[r] on rank q = the last step for which rank r's gradients are complete   (written by r, read by q's optimizer)
//   done[r]   on rank q = the last step for which rank r has finished READING q's gradients (written by r; q waits for it
//                          before the first gradient write of the next step -- normally long satisfied)
// Every spin is bounded (~4 s of clock64): on time-out the kernel records an error code and proceeds, so a lost peer cannot
// hang the GPU.
#include "common.cuh"

namespace ctu {

constexpr int kPeerMax = 16;
constexpr long long kSpinLimit = 8000000000LL;      // clock cycles (~4 s)

struct PeerFlags {                // layout of the flag page at the head of every rank's shared allocation
    unsigned int arrive[kPeerMax];
    unsigned int done[kPeerMax];
    unsigned int finished_blocks;   // local: blocks of the current optimizer launch that have finished reading
    unsigned int error;             // local: 0 ok, 1 arrive time-out, 2 done time-out
    unsigned int seq;               // local: completed steps
    unsigned int pad;
};

struct PeerPtrs {
    const float* grad[kPeerMax];    // flat gradient buffer of every rank (own rank included)
    PeerFlags* flags[kPeerMax];
    int world, rank;
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// after the last gradient kernel of step seq+1: tell every peer that this rank's buffer is complete
__global__ void peer_signal_kernel(PeerPtrs pp) {
    const int r = threadIdx.x;
    if (r >= pp.world) return;
    const unsigned int step = pp.flags[pp.rank]->seq + 1;
    __threadfence_system();
    st_release_sys(&pp.flags[r]->arrive[pp.rank], step);
}

// before the first gradient write of step seq+1: every peer has finished reading this rank's gradients of step seq
__global__ void peer_wait_done_kernel(PeerPtrs pp) {
    const int r = threadIdx.x;
    if (r >= pp.world) return;
    PeerFlags* mine = pp.flags[pp.rank];
    const unsigned int want = mine->seq;
    const long long t0 = clock64();
    while (ld_acquire_sys(&mine->done[r]) < want) {
        if (clock64() - t0 > kSpinLimit) {
            atomicMax(&mine->error, 2u);
            break;
        }
    }
}

struct OptChunkP {
    float* param;
    long long flat_off;
    int count;
    int pad;
};
struct OptHyperP {
    float beta1, beta2, eps, weight_decay, momentum, alpha, grad_scale;
    int kind, amsgrad;
    double b1d, b2d;
    float omb1, omb2, oma;
};

// The single-GPU rule of optim.cu (optim_step_kernel) with the gradient taken as the rank-ordered mean over all peers.
__global__ void __launch_bounds__(256) optim_step_peer_kernel(const OptChunkP* __restrict__ chunks, PeerPtrs pp, float* __restrict__ s0,
                                                              float* __restrict__ s1, float* __restrict__ s2,
                                                              const double* __restrict__ lr_p, const long long* __restrict__ step_p,
                                                              float* __restrict__ tail_out, long long tail_off, int tail_n,
                                                              int n_chunks, OptHyperP hp) {
    PeerFlags* mine = pp.flags[pp.rank];
    const unsigned int want = mine->seq + 1;
    if (threadIdx.x < pp.world) {
        const long long t0 = clock64();
        while (ld_acquire_sys(&mine->arrive[threadIdx.x]) < want) {
            if (clock64() - t0 > kSpinLimit) {
                atomicMax(&mine->error, 1u);
                break;
            }
        }
    }
    __syncthreads();
    const float inv_world = 1.f / (float)pp.world;
    if (blockIdx.x < n_chunks) {
        const OptChunkP ck = chunks[blockIdx.x];
        const double lr = *lr_p;
        const long long step = *step_p + 1;
        float step_size = (float)lr, bc2_sqrt = 1.f;
        if (hp.kind <= 1) {
            const double bc1 = 1.0 - pow(hp.b1d, (double)step);
            const double bc2 = 1.0 - pow(hp.b2d, (double)step);
            step_size = (float)(lr / bc1);
            bc2_sqrt = (float)sqrt(bc2);
        }
        for (int i = threadIdx.x; i < ck.count; i += 256) {
            const long long f = ck.flat_off + i;
            float g = 0.f;
            for (int r = 0; r < pp.world; ++r) g += __ldcv(pp.grad[r] + f);     // rank order: identical on every rank
            g *= inv_world * hp.grad_scale;
            float p = ck.param[i];
            if (hp.kind <= 1) {
                if (hp.weight_decay != 0.f) {
                    if (hp.kind == 0) g = fmaf(p, hp.weight_decay, g);
                    else p *= 1.f - (float)lr * hp.weight_decay;
                }
                float m = s0[f], v = s1[f];
                m = m + (g - m) * hp.omb1;
                v = v * hp.beta2 + hp.omb2 * g * g;
                s0[f] = m;
                s1[f] = v;
                float vv = v;
                if (hp.amsgrad) {
                    vv = fmaxf(s2[f], v);
                    s2[f] = vv;
                }
                p -= step_size * (m / (sqrtf(vv) / bc2_sqrt + hp.eps));
            } else if (hp.kind == 2) {
                if (hp.weight_decay != 0.f) g = fmaf(p, hp.weight_decay, g);
                float sq = s0[f];
                sq = sq * hp.alpha + hp.oma * g * g;
                s0[f] = sq;
                const float avg = sqrtf(sq) + hp.eps;
                if (hp.momentum > 0.f) {
                    const float buf = s1[f] * hp.momentum + g / avg;
                    s1[f] = buf;
                    p -= (float)lr * buf;
                } else {
                    p -= (float)lr * (g / avg);
                }
            } else {
                if (hp.weight_decay != 0.f) g = fmaf(p, hp.weight_decay, g);
                if (hp.momentum != 0.f) {
                    const float buf = step == 1 ? g : s0[f] * hp.momentum + g;
                    s0[f] = buf;
                    g = buf;
                }
                p -= (float)lr * g;
            }
            ck.param[i] = p;
        }
    } else if (threadIdx.x < tail_n) {
        // the extra block: the loss components at the tail of the buffers, averaged the same way
        float g = 0.f;
        for (int r = 0; r < pp.world; ++r) g += __ldcv(pp.grad[r] + tail_off + threadIdx.x);
        tail_out[threadIdx.x] = g * inv_world;
    }
    // the last block to finish reading tells every peer that their buffers are free again
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int prev = atomicAdd(&mine->finished_blocks, 1u);
        if (prev == gridDim.x - 1) {
            mine->finished_blocks = 0;
            __threadfence_system();
            for (int r = 0; r < pp.world; ++r) st_release_sys(&pp.flags[r]->done[pp.rank], want);
        }
    }
}

// one thread, after the optimizer: seq += 1 (the step counter itself is advanced by ctu_optim_post)
__global__ void peer_advance_kernel(PeerFlags* mine) { mine->seq += 1; }

}  // namespace ctu

using namespace ctu;

static int fill_ptrs(PeerPtrs& pp, const void* const* h_grads, void* const* h_flags, int world, int rank, const char* what) {
    CTU_REQUIRE(world >= 1 && world <= kPeerMax && rank >= 0 && rank < world && h_grads && h_flags, "%s: world %d rank %d", what, world, rank);
    for (int r = 0; r < kPeerMax; ++r) {
        pp.grad[r] = r < world ? (const float*)h_grads[r] : nullptr;
        pp.flags[r] = r < world ? (PeerFlags*)h_flags[r] : nullptr;
        if (r < world) CTU_REQUIRE(pp.grad[r] && pp.flags[r], "%s: null peer pointer %d", what, r);
    }
    pp.world = world;
    pp.rank = rank;
    return CTU_OK;
}

extern "C" {

int ctu_peer_flag_bytes(void) { return (int)((sizeof(PeerFlags) + 255) / 256 * 256); }

/* allocation that other processes of the node can map: cudaMalloc + cudaIpcGetMemHandle (64-byte handle) */
int ctu_peer_alloc(long long bytes, void** ptr, unsigned char* handle64) {
    CTU_REQUIRE(bytes > 0 && ptr && handle64, "ctu_peer_alloc: bad arguments");
    cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
    if (e == cudaSuccess) e = cudaMemset(*ptr, 0, (size_t)bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, *ptr);
    if (e != cudaSuccess) {
        set_error("ctu_peer_alloc: %s", cudaGetErrorString(e));
        return (int)e;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle64, &h, 64);
    return CTU_OK;
}

int ctu_peer_open(const unsigned char* handle64, void** ptr) {
    CTU_REQUIRE(handle64 && ptr, "ctu_peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        set_error("ctu_peer_open: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return CTU_OK;
}

int ctu_peer_close(void* ptr) {
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    if (e != cudaSuccess) {
        set_error("ctu_peer_close: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return CTU_OK;
}

int ctu_peer_free(void* ptr) {
    cudaError_t e = cudaFree(ptr);
    if (e != cudaSuccess) {
        set_error("ctu_peer_free: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return CTU_OK;
}

int ctu_peer_signal(const void* const* h_grads, void* const* h_flags, int world, int rank, ctu_stream stream) {
    PeerPtrs pp;
    int rc = fill_ptrs(pp, h_grads, h_flags, world, rank, "ctu_peer_signal");
    if (rc != CTU_OK) return rc;
    peer_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pp);
    return check_launch("peer_signal_kernel");
}

int ctu_peer_wait_done(const void* const* h_grads, void* const* h_flags, int world, int rank, ctu_stream stream) {
    PeerPtrs pp;
    int rc = fill_ptrs(pp, h_grads, h_flags, world, rank, "ctu_peer_wait_done");
    if (rc != CTU_OK) return rc;
    peer_wait_done_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pp);
    return check_launch("peer_wait_done_kernel");
}

int ctu_optim_step_peer(int kind, const void* chunks, int n_chunks, const void* const* h_grads, void* const* h_flags, int world,
                        int rank, float* state0, float* state1, float* state2, const double* lr, const long long* step,
                        double beta1, double beta2, double eps, double weight_decay, double momentum, double alpha, int amsgrad,
                        double grad_scale, float* tail_out, long long tail_off, int tail_n, ctu_stream stream) {
    CTU_REQUIRE(kind >= 0 && kind <= 3 && chunks && n_chunks >= 1 && lr && step, "ctu_optim_step_peer: bad arguments");
    CTU_REQUIRE(tail_n >= 0 && tail_n <= 32 && (tail_n == 0 || tail_out), "ctu_optim_step_peer: tail of at most 32 floats");
    const bool need0 = kind != 3 || momentum != 0.0, need1 = kind <= 1 || (kind == 2 && momentum > 0.0);
    CTU_REQUIRE((!need0 || state0) && (!need1 || state1) && (!(kind <= 1 && amsgrad) || state2), "ctu_optim_step_peer: missing state");
    PeerPtrs pp;
    int rc = fill_ptrs(pp, h_grads, h_flags, world, rank, "ctu_optim_step_peer");
    if (rc != CTU_OK) return rc;
    OptHyperP hp;
    hp.beta1 = (float)beta1, hp.beta2 = (float)beta2, hp.eps = (float)eps, hp.weight_decay = (float)weight_decay;
    hp.momentum = (float)momentum, hp.alpha = (float)alpha, hp.grad_scale = (float)grad_scale;
    hp.omb1 = (float)(1.0 - beta1), hp.omb2 = (float)(1.0 - beta2), hp.oma = (float)(1.0 - alpha);
    hp.kind = kind, hp.amsgrad = amsgrad, hp.b1d = beta1, hp.b2d = beta2;
    cudaStream_t st = (cudaStream_t)stream;
    optim_step_peer_kernel<<<n_chunks + 1, 256, 0, st>>>((const OptChunkP*)chunks, pp, state0, state1, state2, lr, step, tail_out,
                                                         tail_off, tail_n, n_chunks, hp);
    rc = check_launch("optim_step_peer_kernel");
    if (rc != CTU_OK) return rc;
    peer_advance_kernel<<<1, 1, 0, st>>>(pp.flags[rank]);
    return check_launch("peer_advance_kernel");
}

/* host read of the local error flag (0 ok, 1 a peer's gradients never arrived, 2 a peer never released this rank's
 * buffer); synchronises the device */
int ctu_peer_error(const void* flags) {
    PeerFlags h;
    if (cudaMemcpy(&h, flags, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int)h.error;
}

}  // extern "C"
