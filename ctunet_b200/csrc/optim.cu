// The tail of a training iteration on the device: loss weighting, the optimizer update over ALL parameters in one
// launch, and the learning-rate scheduler -- so that a captured step contains no library (at::) kernels and no host logic.
//   loss weighting / sum          ctunet/pytorch/ProblemHandler.py:59-91, 241-298   (lambda * term, sum(list))
//   optimizers                    ctunet/pytorch/Model.py:510-541   (Adam / AdamW with amsgrad=True, RMSprop, SGD)
//   ReduceLROnPlateau per ITERATION on the training loss     Model.py:369-371, 544-546  (torch defaults: mode 'min',
//                                 factor 0.1, patience 10, threshold 1e-4 'rel', cooldown 0, min_lr 0, eps 1e-8)
// The update rules restate torch.optim's single-tensor implementations (torch 2.11) in fp32 with the bias corrections
// and the step size formed in double, like the Python code does.
#include "common.cuh"

namespace ctu {

struct OptChunk {        // 24 bytes; built once on the host (optim.py), lives on the device
    float* param;        // first element of this chunk in the parameter tensor
    long long flat_off;  // offset of the chunk in the flat gradient / state buffers
    int count;           // elements (<= kChunk)
    int pad;
};
constexpr int kChunk = 1024;
constexpr int kOptThreads = 256;

struct OptHyper {
    float beta1, beta2, eps, weight_decay, momentum, alpha, grad_scale;
    int kind;            // 0 adam, 1 adamw, 2 rmsprop, 3 sgd
    int amsgrad;
    double b1d, b2d;         // the betas in double for the bias corrections (Python: 1 - beta ** step)
    float omb1, omb2, oma;   // 1 - beta1, 1 - beta2, 1 - alpha formed in double on the host (as Python does)
};

__global__ void __launch_bounds__(kOptThreads) optim_step_kernel(const OptChunk* __restrict__ chunks, const float* __restrict__ grad,
                                                                 float* __restrict__ s0, float* __restrict__ s1,
                                                                 float* __restrict__ s2, const double* __restrict__ lr_p,
                                                                 const long long* __restrict__ step_p, OptHyper hp) {
    const OptChunk ck = chunks[blockIdx.x];
    const double lr = *lr_p;
    const long long step = *step_p + 1;                       // this call is step number `step` (1-based, as torch counts)
    float step_size = (float)lr, bc2_sqrt = 1.f;
    if (hp.kind <= 1) {
        const double bc1 = 1.0 - pow(hp.b1d, (double)step);
        const double bc2 = 1.0 - pow(hp.b2d, (double)step);
        step_size = (float)(lr / bc1);
        bc2_sqrt = (float)sqrt(bc2);
    }
    for (int i = threadIdx.x; i < ck.count; i += kOptThreads) {
        const long long f = ck.flat_off + i;
        float p = ck.param[i];
        float g = grad[f] * hp.grad_scale;
        if (hp.kind == 0 || hp.kind == 1) {
            if (hp.weight_decay != 0.f) {
                if (hp.kind == 0) g = fmaf(p, hp.weight_decay, g);          // Adam: L2 term added to the gradient
                else p *= 1.f - (float)lr * hp.weight_decay;                // AdamW: decoupled decay
            }
            float m = s0[f], v = s1[f];
            m = m + (g - m) * hp.omb1;                                   // exp_avg.lerp_(grad, 1 - beta1)
            v = v * hp.beta2 + hp.omb2 * g * g;                            // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
            s0[f] = m;
            s1[f] = v;
            float vv = v;
            if (hp.amsgrad) {
                vv = fmaxf(s2[f], v);
                s2[f] = vv;
            }
            const float denom = sqrtf(vv) / bc2_sqrt + hp.eps;
            p -= step_size * (m / denom);                                   // param.addcdiv_(exp_avg, denom, value=-step_size)
        } else if (hp.kind == 2) {                                           // RMSprop (alpha, eps, momentum; not centered)
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
        } else {                                                             // SGD with momentum (dampening 0, no nesterov)
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
}

// state: double[8] = { lr, best, num_bad_epochs, cooldown_counter, factor, patience, threshold, min_lr }
//        (cooldown length and eps ride in [8], [9]); step counter is a separate long long.
__global__ void optim_post_kernel(long long* __restrict__ step_p, double* __restrict__ st, const float* __restrict__ loss,
                                  int use_plateau) {
    *step_p += 1;
    if (!use_plateau) return;
    const double current = (double)*loss;
    double best = st[1], bad = st[2], cool = st[3];
    const double factor = st[4], patience = st[5], threshold = st[6], min_lr = st[7], cooldown = st[8], eps = st[9];
    if (current < best * (1.0 - threshold)) {     // mode 'min', threshold_mode 'rel'
        best = current;
        bad = 0;
    } else {
        bad += 1;
    }
    if (cool > 0) {
        cool -= 1;
        bad = 0;
    }
    if (bad > patience) {
        const double old_lr = st[0];
        const double new_lr = fmax(old_lr * factor, min_lr);
        if (old_lr - new_lr > eps) st[0] = new_lr;
        cool = cooldown;
        bad = 0;
    }
    st[1] = best;
    st[2] = bad;
    st[3] = cool;
}

// comps[i] = lambda[i] * terms[i] (i < n), comps[n] = their sum in order (sum(list) starts from int 0: exact);
// mirror (nullable): a second copy, e.g. the tail of the data-parallel flat gradient buffer (averaged with the gradients)
struct TermPtrs {
    const float* p[8];
    float lambda[8];
};
__global__ void loss_combine_kernel_v(TermPtrs tp, int n, float* __restrict__ comps, float* __restrict__ mirror) {
    float total = 0.f;
    for (int i = 0; i < n; ++i) {
        const float t = tp.lambda[i] * *tp.p[i];
        comps[i] = t;
        if (mirror) mirror[i] = t;
        total = i == 0 ? t : total + t;
    }
    comps[n] = total;
    if (mirror) mirror[n] = total;
}

}  // namespace ctu

using namespace ctu;

extern "C" int ctu_optim_chunk_bytes(void) { return (int)sizeof(OptChunk); }
extern "C" int ctu_optim_chunk_elems(void) { return kChunk; }

extern "C" int ctu_optim_step(int kind, const void* chunks, int n_chunks, const float* flat_grad, float* state0, float* state1,
                              float* state2, const double* lr, const long long* step, double beta1, double beta2, double eps,
                              double weight_decay, double momentum, double alpha, int amsgrad, double grad_scale,
                              ctu_stream stream) {
    CTU_REQUIRE(kind >= 0 && kind <= 3, "ctu_optim_step: kind %d (0 adam, 1 adamw, 2 rmsprop, 3 sgd)", kind);
    CTU_REQUIRE(chunks && flat_grad && lr && step && n_chunks >= 1, "ctu_optim_step: null pointer / no chunks");
    const bool need0 = kind != 3 || momentum != 0.0, need1 = kind <= 1 || (kind == 2 && momentum > 0.0);
    CTU_REQUIRE((!need0 || state0 != nullptr) && (!need1 || state1 != nullptr), "ctu_optim_step: missing state buffer");
    CTU_REQUIRE(!(kind <= 1 && amsgrad) || state2, "ctu_optim_step: amsgrad needs state2");
    OptHyper hp;
    hp.beta1 = (float)beta1, hp.beta2 = (float)beta2, hp.eps = (float)eps, hp.weight_decay = (float)weight_decay;
    hp.momentum = (float)momentum, hp.alpha = (float)alpha, hp.grad_scale = (float)grad_scale;
    hp.omb1 = (float)(1.0 - beta1), hp.omb2 = (float)(1.0 - beta2), hp.oma = (float)(1.0 - alpha);
    hp.kind = kind, hp.amsgrad = amsgrad, hp.b1d = beta1, hp.b2d = beta2;
    optim_step_kernel<<<n_chunks, kOptThreads, 0, (cudaStream_t)stream>>>((const OptChunk*)chunks, flat_grad, state0, state1,
                                                                         state2, lr, step, hp);
    return check_launch("optim_step_kernel");
}

extern "C" int ctu_optim_post(long long* step, double* sched_state, const float* loss, int use_plateau, ctu_stream stream) {
    CTU_REQUIRE(step && (!use_plateau || (sched_state && loss)), "ctu_optim_post: null pointer");
    optim_post_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step, sched_state, loss, use_plateau);
    return check_launch("optim_post_kernel");
}

extern "C" int ctu_loss_combine(const float* const* h_terms, const float* h_lambdas, int n, float* comps, float* mirror,
                                ctu_stream stream) {
    CTU_REQUIRE(h_terms && h_lambdas && comps && n >= 1 && n <= 8, "ctu_loss_combine: 1..8 terms");
    TermPtrs tp;
    for (int i = 0; i < 8; ++i) {
        tp.p[i] = i < n ? h_terms[i] : nullptr;
        tp.lambda[i] = i < n ? h_lambdas[i] : 0.f;
    }
    for (int i = 0; i < n; ++i) CTU_REQUIRE(tp.p[i], "ctu_loss_combine: term %d is null", i);
    loss_combine_kernel_v<<<1, 1, 0, (cudaStream_t)stream>>>(tp, n, comps, mirror);
    return check_launch("loss_combine_kernel");
}
