// Action selection.
//   /root/reference/src/components/action_selectors.py:45-63  EpsilonGreedyActionSelector.select_action
// Greedy part: unavailable -> -inf, FIRST max index (torch max(dim)[1] on CPU).  Exploration: the caller passes
// two uniform [0,1) draws per (env, agent) (torch.rand on device): u_pick < eps picks the floor(u_act * n_avail)-th
// available action (== Categorical(avail) in distribution; bit parity is only defined for the greedy index).
#include "common.cuh"

__global__ void select_actions_kernel(const float* __restrict__ q, long long q_stride_b,
                                      const int32_t* __restrict__ avail, long long avail_stride_b,
                                      const float* __restrict__ u_pick, const float* __restrict__ u_act,
                                      const int32_t* __restrict__ est_flags, float eps_host,
                                      const float* __restrict__ eps_dev, long long* __restrict__ out,
                                      long long out_stride_b, int B, int na, int A) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * na) return;
    // a device-resident epsilon lets one captured CUDA graph of a rollout serve every point of the schedule (and test mode: 0)
    const float eps = eps_dev ? __ldg(eps_dev) : eps_host;
    int b = idx / na, a = idx - b * na;
    if (est_flags && !(est_flags[b] & 2)) return;  // env not in the runner's envs_not_terminated list
    const float* qr = q + (size_t)b * q_stride_b + (size_t)a * A;
    const int32_t* av = avail + (size_t)b * avail_stride_b + (size_t)a * A;
    int best = 0, n_avail = 0;
    float bv = -INFINITY;
    bool have = false;
    for (int k = 0; k < A; k++) {
        bool ok = av[k] != 0;
        n_avail += ok;
        float v = ok ? qr[k] : -INFINITY;
        if (!have || v > bv) { if (!have) { best = k; bv = v; have = true; } else { best = k; bv = v; } }
    }
    int pick = best;
    if (u_pick && eps > 0.f && u_pick[idx] < eps && n_avail > 0) {
        int target = min((int)(u_act[idx] * (float)n_avail), n_avail - 1), seen = 0;
        for (int k = 0; k < A; k++)
            if (av[k] != 0) { if (seen == target) { pick = k; break; } seen++; }
    }
    out[(size_t)b * out_stride_b + a] = pick;
}

extern "C" int refil_select_actions(const float* q, long long q_stride_b, const int32_t* avail,
                                    long long avail_stride_b, const float* u_pick, const float* u_act,
                                    const int32_t* est_flags, float epsilon, const float* epsilon_dev,
                                    long long* actions_out, long long out_stride_b, int B, int n_agents, int n_actions,
                                    cudaStream_t stream) {
    REFIL_CHECK_ARG(q && avail && actions_out && B > 0 && n_agents > 0 && n_actions > 0, "select_actions: bad arguments");
    REFIL_CHECK_ARG((epsilon <= 0.f && !epsilon_dev) || (u_pick && u_act), "select_actions: epsilon > 0 needs u_pick and u_act");
    int n = B * n_agents;
    select_actions_kernel<<<refil_cdiv(n, 128), 128, 0, stream>>>(q, q_stride_b, avail, avail_stride_b, u_pick, u_act,
                                                                  est_flags, epsilon, epsilon_dev, actions_out,
                                                                  out_stride_b, B, n_agents, n_actions);
    REFIL_CHECK_LAUNCH("select_actions");
    return REFIL_OK;
}
