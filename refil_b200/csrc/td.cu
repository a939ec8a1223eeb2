// TD-target plumbing of QLearner.train (kernels K5/K8) and the optimiser (row L9, K10).
//
// Replaces, on [N = B*T] rows of (b, t):
//   /root/reference/src/learners/q_learner.py:92,94-96   gather of the chosen action utilities (3 copies)
//   /root/reference/src/learners/q_learner.py:111-126    unavailable -> -9999999, double-Q argmax / gather
//   /root/reference/src/learners/q_learner.py:68-72,157-172  mask, targets, TD error, masked L2 losses
//   /root/reference/src/learners/q_learner.py:175-178    clip_grad_norm_ + RMSprop.step
// Losses are produced un-normalised (sums) together with sum(mask): the division by sum(mask) happens in the
// optimiser kernel AFTER the (optional) NCCL all-reduce of [grads | stats], which makes the N-GPU step equal to
// the single-GPU step on the concatenated batch.
#include "common.cuh"

#define NEG_UNAVAIL (-9999999.0f)

// chosen[c, n, a] = Q[c, n, a, actions[n, a]]
__global__ void gather_chosen_kernel(const float* __restrict__ Q, const long long* __restrict__ actions,
                                     float* __restrict__ chosen, int C, long long Nna, int A) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)C * Nna) return;
    long long r = idx % Nna;
    int act = (int)actions[r];
    chosen[idx] = (act >= 0 && act < A) ? Q[idx * A + act] : 0.f;
}

// dQ[c, n, a, :] = onehot(actions[n, a]) * dchosen[c, n, a]   (zero rows at t = T-1: no transition starts there)
__global__ void scatter_dq_kernel(const float* __restrict__ dchosen, const long long* __restrict__ actions,
                                  float* __restrict__ dQ, int C, long long Nna, int A, int T, int na) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)C * Nna * A) return;
    long long row = idx / A;
    int k = (int)(idx - row * A);
    long long r = row % Nna;
    int t = (int)((r / na) % T);
    float v = 0.f;
    if (t < T - 1 && (int)actions[r] == k) v = dchosen[row];
    dQ[idx] = v;
}

// tgt[n, a] = Q_target[n, a, argmax_k masked(Q_online[n, a, k])]  (double-Q) or max_k masked(Q_target)
__global__ void target_max_kernel(const float* __restrict__ q_online, const float* __restrict__ q_target,
                                  const int32_t* __restrict__ avail, float* __restrict__ tgt,
                                  long long* __restrict__ cur_max, long long Nna, int A, int double_q) {
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= Nna) return;
    const float* qo = q_online + r * A;
    const float* qt = q_target + r * A;
    const int32_t* av = avail + r * A;
    int best = 0;
    float bv = 0.f, out;
    if (double_q) {
        for (int k = 0; k < A; k++) {
            float v = av[k] ? qo[k] : NEG_UNAVAIL;
            if (k == 0 || v > bv) { bv = v; best = k; }
        }
        out = av[best] ? qt[best] : NEG_UNAVAIL;
    } else {
        for (int k = 0; k < A; k++) {
            float v = av[k] ? qt[k] : NEG_UNAVAIL;
            if (k == 0 || v > bv) { bv = v; best = k; }
        }
        out = bv;
    }
    tgt[r] = out;
    if (cur_max) cur_max[r] = best;
}

// stats (f64): 0 sum(mask) 1 sum((mask*td)^2) 2 sum((mask*td_im)^2) 3 sum|mask*td| 4 sum(q_tot*mask) 5 sum(targets*mask)
__global__ void td_loss_kernel(const float* __restrict__ qtot, const float* __restrict__ qtot_im,
                               const float* __restrict__ tgt_tot, const float* __restrict__ reward,
                               const uint8_t* __restrict__ terminated, const long long* __restrict__ filled,
                               float* __restrict__ g_plain, float* __restrict__ g_im, float* __restrict__ targets_out,
                               double* __restrict__ stats, int B, int T, float gamma, float lmbda) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    double s[6] = {0, 0, 0, 0, 0, 0};
    if (n < B * T) {
        int t = n % T;
        float gp = 0.f, gi = 0.f, target = 0.f;
        if (t < T - 1) {
            float mask = (float)filled[n];
            if (t > 0) mask *= 1.f - (float)terminated[n - 1];
            const float term = (float)terminated[n];
            target = reward[n] + gamma * (1.f - term) * tgt_tot[n + 1];
            const float td = (qtot[n] - target) * mask;
            s[0] = mask;
            s[1] = (double)td * td;
            s[3] = fabsf(td);
            s[4] = (double)qtot[n] * mask;
            s[5] = (double)target * mask;
            if (qtot_im) {
                const float tdi = (qtot_im[n] - target) * mask;
                s[2] = (double)tdi * tdi;
                gp = 2.f * (1.f - lmbda) * td * mask;
                gi = 2.f * lmbda * tdi * mask;
            } else {
                gp = 2.f * td * mask;
            }
        }
        g_plain[n] = gp;
        if (g_im) g_im[n] = gi;
        if (targets_out) targets_out[n] = target;
    }
#pragma unroll
    for (int k = 0; k < 6; k++) {
        double v = warp_sum_d(s[k]);
        if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(stats + k, v);
    }
}

// sum of squares of the flat gradient -> stats[slot] (f64)
__global__ void sumsq_kernel(const float* __restrict__ g, long long P, double* __restrict__ out) {
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) {
        float v = g[i];
        s += (double)v * v;
    }
    s = warp_sum_d(s);
    __shared__ double red[32];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        v = warp_sum_d(v);
        if (threadIdx.x == 0) atomicAdd(out, v);
    }
}

// tail[k] = (float) stats[k]: the loss statistics ride behind the flat gradient so ONE all-reduce covers both
__global__ void pack_stats_kernel(const double* __restrict__ stats, float* __restrict__ tail, int n) {
    int k = threadIdx.x;
    if (k < n) tail[k] = (float)stats[k];
}

// g <- g / sum(mask); clip by global norm; RMSprop (torch.optim.RMSprop, centered=False, momentum=0)
__global__ void clip_rmsprop_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ sq,
                                    long long P, const float* __restrict__ mask_sum,
                                    const double* __restrict__ sumsq, float* __restrict__ grad_norm_out, float clip,
                                    float lr, float alpha, float eps, float weight_decay) {
    const float msum = mask_sum ? *mask_sum : 1.f;
    const float inv = 1.f / msum;
    const float norm = (float)sqrt(*sumsq) * inv;
    float coef = clip / (norm + 1e-6f);
    if (coef > 1.f) coef = 1.f;
    const float scale = inv * coef;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) {
        float gv = g[i] * scale;
        g[i] = gv;
        const float pv = p[i];
        if (weight_decay != 0.f) gv = fmaf(weight_decay, pv, gv);
        const float v = alpha * sq[i] + (1.f - alpha) * gv * gv;
        sq[i] = v;
        p[i] = pv - lr * gv / (sqrtf(v) + eps);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && grad_norm_out) *grad_norm_out = norm;
}

extern "C" int refil_gather_chosen(const float* Q, const long long* actions, float* chosen, int copies,
                                   long long rows_per_copy, int n_actions, cudaStream_t stream) {
    REFIL_CHECK_ARG(Q && actions && chosen && copies > 0 && rows_per_copy > 0 && n_actions > 0, "gather_chosen: bad arguments");
    long long n = (long long)copies * rows_per_copy;
    gather_chosen_kernel<<<refil_cdiv(n, 256), 256, 0, stream>>>(Q, actions, chosen, copies, rows_per_copy, n_actions);
    REFIL_CHECK_LAUNCH("gather_chosen");
    return REFIL_OK;
}

extern "C" int refil_scatter_dq(const float* dchosen, const long long* actions, float* dQ, int copies,
                                long long rows_per_copy, int n_actions, int T, int n_agents, cudaStream_t stream) {
    REFIL_CHECK_ARG(dchosen && actions && dQ && copies > 0 && rows_per_copy > 0 && n_actions > 0 && T > 0 && n_agents > 0,
                    "scatter_dq: bad arguments");
    long long n = (long long)copies * rows_per_copy * n_actions;
    scatter_dq_kernel<<<refil_cdiv(n, 256), 256, 0, stream>>>(dchosen, actions, dQ, copies, rows_per_copy, n_actions, T,
                                                             n_agents);
    REFIL_CHECK_LAUNCH("scatter_dq");
    return REFIL_OK;
}

extern "C" int refil_target_max(const float* q_online, const float* q_target, const int32_t* avail, float* tgt,
                                long long* cur_max_actions, long long rows, int n_actions, int double_q,
                                cudaStream_t stream) {
    REFIL_CHECK_ARG(q_target && avail && tgt && rows > 0 && n_actions > 0, "target_max: bad arguments");
    REFIL_CHECK_ARG(!double_q || q_online, "target_max: double_q needs the online utilities");
    target_max_kernel<<<refil_cdiv(rows, 256), 256, 0, stream>>>(q_online, q_target, avail, tgt, cur_max_actions, rows,
                                                                n_actions, double_q);
    REFIL_CHECK_LAUNCH("target_max");
    return REFIL_OK;
}

extern "C" int refil_td_loss(const float* qtot, const float* qtot_im, const float* tgt_tot, const float* reward,
                             const uint8_t* terminated, const long long* filled, float* g_plain, float* g_im,
                             float* targets_out, double* stats, int B, int T, float gamma, float lmbda,
                             cudaStream_t stream) {
    REFIL_CHECK_ARG(qtot && tgt_tot && reward && terminated && filled && g_plain && stats && B > 0 && T > 1,
                    "td_loss: bad arguments");
    REFIL_CHECK_ARG(!qtot_im || g_im, "td_loss: imagine loss needs g_im");
    td_loss_kernel<<<refil_cdiv((long long)B * T, 256), 256, 0, stream>>>(qtot, qtot_im, tgt_tot, reward, terminated,
                                                                         filled, g_plain, g_im, targets_out, stats, B,
                                                                         T, gamma, lmbda);
    REFIL_CHECK_LAUNCH("td_loss");
    return REFIL_OK;
}

extern "C" int refil_grad_sumsq(const float* grads, long long n_params, double* out, cudaStream_t stream) {
    REFIL_CHECK_ARG(grads && out && n_params > 0, "grad_sumsq: bad arguments");
    int blocks = refil_cdiv(n_params, 256 * 8);
    int cap = 2 * refil_num_sms();
    if (blocks > cap) blocks = cap;
    sumsq_kernel<<<blocks, 256, 0, stream>>>(grads, n_params, out);
    REFIL_CHECK_LAUNCH("grad_sumsq");
    return REFIL_OK;
}

extern "C" int refil_pack_stats(const double* stats, float* tail, int n, cudaStream_t stream) {
    REFIL_CHECK_ARG(stats && tail && n > 0 && n <= 32, "pack_stats: bad arguments");
    pack_stats_kernel<<<1, 32, 0, stream>>>(stats, tail, n);
    REFIL_CHECK_LAUNCH("pack_stats");
    return REFIL_OK;
}

extern "C" int refil_clip_rmsprop_step(float* params, float* grads, float* square_avg, long long n_params,
                                       const float* mask_sum, const double* sumsq, float* grad_norm_out,
                                       float grad_clip, float lr, float alpha, float eps, float weight_decay,
                                       cudaStream_t stream) {
    REFIL_CHECK_ARG(params && grads && square_avg && sumsq && n_params > 0, "clip_rmsprop_step: bad arguments");
    int blocks = refil_cdiv(n_params, 256 * 4);
    int cap = 4 * refil_num_sms();
    if (blocks > cap) blocks = cap;
    clip_rmsprop_kernel<<<blocks, 256, 0, stream>>>(params, grads, square_avg, n_params, mask_sum, sumsq,
                                                   grad_norm_out, grad_clip, lr, alpha, eps, weight_decay);
    REFIL_CHECK_LAUNCH("clip_rmsprop_step");
    return REFIL_OK;
}
