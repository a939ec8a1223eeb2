// Fused acting step of the feed-forward entity-attention agent.
//
// One launch computes, for timestep t of the rollout tensors and every environment, the utilities
//   q = fc2( relu( rowmask( out_trans( MHA( in_trans( relu( fc1(entities) ) ), obs_mask ) ) ) ) )         [E, na, A]
// i.e. EntityAttentionFFAgent.forward (/root/reference/src/modules/agents/entity_ff_agent.py:30-57) on the inputs of
// EntityMAC._build_inputs (controllers/entity_controller.py:11-30), read IN PLACE from the EpisodeBatch tensors (row stride T: no
// time-slice copies, no packed input).  The generic acting path runs the same five layers as ~12 launches (three tensor-core GEMMs of
// < 2 us of work each, the attention kernel, a head GEMM and half a dozen slice copies): with the small Group Matching networks
// (d = 64, <= 8 entities of <= 32 features) a rollout timestep was bound by their fixed costs.  Here the weights sit in one CTA's
// shared memory and registers (54 KB + 96 registers per thread: three CTAs per SM) and 16 entity rows (two 8-slot or four 4-slot environments) advance per iteration through the five stages on the fp32 pipe:
//   stage GEMMs: thread = (output column j, row group g), inner loop over k with W^T[k][j] (conflict-free) and the stage input stored
//   TRANSPOSED [k][row] so that the 8 rows of a group arrive as two broadcast 128-bit loads; every stage writes its output transposed
//   for the next one;
//   attention: thread = (env, agent i, head h, half of the head dim): 8 partial products + one shuffle per logit, softmax in
//   registers (a fully masked row gives zeros, attention.py:58-60), 8 output features per thread.
// fp32 FFMA throughout (no TF32): utilities agree with the reference to ~1e-6; the greedy index is compared bit-exactly in the tests.
#include "common.cuh"

#define FA_D 64
#define FA_H 4
#define FA_HD 16
#define FA_NE_MAX 8
#define FA_R 16                  // rows of a stage tile: 16 / NE environments advance per iteration (NE = entity slots per env: 4 or 8)
#define FA_THREADS 128
#define FA_EIN_MAX 32
#define FA_A_MAX 16

struct FfActArgs {
    const float* entities; int ed;            // [E, T, ne, ed]
    const long long* actions; int n_actions;  // [E, T, na, 1] or null: last-action one-hot appended to agent rows (entity_last_action)
    const uint8_t* obs_mask; int mask_rows;   // [E, T, mask_rows, ne]; rows = ne (obs_mask) or na (gt_mask with gt_obs_mask)
    const uint8_t* entity_mask;               // [E, T, ne]
    const float *w1, *b1, *win, *wout, *bout, *w2, *b2;
    float* q;                                 // [E, na, A]
    int E, T, t, ne, na, ein, A;
    // optional fused epsilon-greedy selection (components/action_selectors.py:45-63; same rules as select_actions_kernel)
    const int32_t* avail;                     // [E, T, na, A] or null: no selection
    const float *u_pick, *u_act;              // [E, na] uniforms (null: greedy)
    const int32_t* est_flags;                 // [E] env flags (bit 1: still in the runner's live list) or null
    float eps_host; const float* eps_dev;     // epsilon: host value, overridden by the device scalar when given
    long long* actions_out;                   // the rollout's actions tensor [E, T, na, 1]: written at timestep t
};

// shared-memory strides: +1 float per transposed weight row, so that the transposing prologue stores (k fastest across a warp)
// spread over the banks instead of hitting one
#define FA_LD1 (FA_D + 1)

template <int NE>
__global__ void __launch_bounds__(FA_THREADS, 3) ff_agent_act_kernel(FfActArgs a) {
    constexpr int EPI = FA_R / NE;
    extern __shared__ __align__(16) float sm[];
    const int ein = a.ein, ne = a.ne, na = a.na, A = a.A;
    float* w1t = sm;                                  // [ein][65]
    float* woutt = w1t + FA_EIN_MAX * FA_LD1;         // [64][65]
    float* w2s = woutt + FA_D * FA_LD1;               // [A][64]
    float* bias = w2s + FA_A_MAX * FA_D;              // b1[64] | bout[64] | b2[16]
    float* xin = bias + 2 * FA_D + FA_A_MAX;          // [ein][R]   stage-1 input, transposed; later the stage-2 partial sums
    float* x1t = xin + FA_EIN_MAX * FA_R;             // [64][R]
    float* qkv = x1t + FA_D * FA_R;                   // [R][192]
    float* attt = qkv + FA_R * 3 * FA_D;              // [64][R]    attention output, transposed (rows = env * NE_MAX + agent)
    float* x2t = attt + FA_D * FA_R;                  // [64][R]
    __shared__ uint8_t s_obs[FA_R][NE], s_em[FA_R];     // rows = env * NE + entity slot
    __shared__ float s_q[FA_R][FA_A_MAX];               // utilities of the iteration's agents, for the fused selection
    const int tid = threadIdx.x;
    const int j = tid & 63, g = tid >> 6;              // GEMM stages: output column j, row group / reduction half g
    for (int f = tid; f < ein * FA_D; f += FA_THREADS) { const int jj = f / ein, k = f - jj * ein; w1t[k * FA_LD1 + jj] = __ldg(a.w1 + f); }
    for (int f = tid; f < FA_D * FA_D; f += FA_THREADS) { const int jj = f / FA_D, k = f - jj * FA_D; woutt[k * FA_LD1 + jj] = __ldg(a.wout + f); }
    for (int f = tid; f < A * FA_D; f += FA_THREADS) w2s[f] = __ldg(a.w2 + f);
    if (tid < FA_D) { bias[tid] = __ldg(a.b1 + tid); bias[FA_D + tid] = __ldg(a.bout + tid); }
    if (tid < A) bias[2 * FA_D + tid] = __ldg(a.b2 + tid);
    // in_trans (192 x 64 = 12288 weights = 96 per thread) lives in REGISTERS: thread (j, g) holds rows j, j + 64, j + 128 of W_in for
    // the reduction half k in [32 g, 32 g + 32) -- 48 KB less shared memory (three CTAs per SM instead of two) and 48 FMAs per 4
    // broadcast loads in the heaviest stage
    float win[3][32];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(a.win + (size_t)(c * FA_D + j) * FA_D + 32 * g + k));
            win[c][k] = v.x; win[c][k + 1] = v.y; win[c][k + 2] = v.z; win[c][k + 3] = v.w;
        }
    const float inv_scale = 0.25f;                     // 1 / sqrt(head dim 16)
    __syncthreads();
    for (int e0 = blockIdx.x * EPI; e0 < a.E; e0 += gridDim.x * EPI) {
        // ---- stage 0: inputs of FA_EPI environments, transposed; masks ------------------------------------------------------
        for (int f = tid; f < FA_R * ein; f += FA_THREADS) {
            const int r = f / ein, k = f - r * ein, ev = r / NE, en = r - ev * NE, e = e0 + ev;
            float v = 0.f;
            if (e < a.E && en < ne) {
                const size_t row = (size_t)e * a.T + a.t;
                if (k < a.ed) v = __ldg(a.entities + (row * ne + en) * a.ed + k);
                else if (a.actions && a.t > 0 && en < na)      // one-hot of the previous action (entity_controller.py:14-27)
                    v = (__ldg(a.actions + ((row - 1) * na + en)) == (long long)(k - a.ed)) ? 1.f : 0.f;
            }
            xin[k * FA_R + r] = v;
        }
        for (int f = tid; f < FA_R * NE; f += FA_THREADS) {
            const int r = f / NE, jj = f - r * NE, ev = r / NE, i = r - ev * NE, e = e0 + ev;
            uint8_t m = 1;
            if (e < a.E && i < na && jj < ne) m = a.obs_mask[(((size_t)e * a.T + a.t) * a.mask_rows + i) * ne + jj];
            s_obs[r][jj] = m;
        }
        if (tid < FA_R) {
            const int ev = tid / NE, en = tid - ev * NE, e = e0 + ev;
            s_em[tid] = (e < a.E && en < ne) ? a.entity_mask[((size_t)e * a.T + a.t) * ne + en] : 1;
        }
        __syncthreads();
        // ---- stage 1: x1 = relu(fc1(x))  [R, 64]; thread = (column j, rows 8 g .. 8 g + 7) -----------------------------------
        {
            float acc[8];
#pragma unroll
            for (int r = 0; r < 8; r++) acc[r] = bias[j];
            for (int k = 0; k < ein; k++) {
                const float w = w1t[k * FA_LD1 + j];
                const float4 x0 = *reinterpret_cast<const float4*>(xin + k * FA_R + 8 * g);
                const float4 x1 = *reinterpret_cast<const float4*>(xin + k * FA_R + 8 * g + 4);
                acc[0] = fmaf(x0.x, w, acc[0]); acc[1] = fmaf(x0.y, w, acc[1]); acc[2] = fmaf(x0.z, w, acc[2]); acc[3] = fmaf(x0.w, w, acc[3]);
                acc[4] = fmaf(x1.x, w, acc[4]); acc[5] = fmaf(x1.y, w, acc[5]); acc[6] = fmaf(x1.z, w, acc[6]); acc[7] = fmaf(x1.w, w, acc[7]);
            }
            float* o = x1t + j * FA_R + 8 * g;
            *reinterpret_cast<float4*>(o) = make_float4(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f), fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f));
            *reinterpret_cast<float4*>(o + 4) = make_float4(fmaxf(acc[4], 0.f), fmaxf(acc[5], 0.f), fmaxf(acc[6], 0.f), fmaxf(acc[7], 0.f));
        }
        __syncthreads();
        // ---- stage 2: QKV = in_trans(x1)  [R, 192], no bias; thread = (columns j, j + 64, j + 128; reduction half g), all 16 rows ----
        {
            float acc[3][FA_R];
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int r = 0; r < FA_R; r++) acc[c][r] = 0.f;
#pragma unroll
            for (int k = 0; k < 32; k++) {
                const float* xr = x1t + (32 * g + k) * FA_R;
                float xv[FA_R];
#pragma unroll
                for (int r4 = 0; r4 < FA_R; r4 += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(xr + r4);
                    xv[r4] = v.x; xv[r4 + 1] = v.y; xv[r4 + 2] = v.z; xv[r4 + 3] = v.w;
                }
#pragma unroll
                for (int r = 0; r < FA_R; r++) {
                    acc[0][r] = fmaf(xv[r], win[0][k], acc[0][r]);
                    acc[1][r] = fmaf(xv[r], win[1][k], acc[1][r]);
                    acc[2][r] = fmaf(xv[r], win[2][k], acc[2][r]);
                }
            }
            // the two reduction halves meet in shared memory: half 1 parks its partial sums (xin is dead by now), half 0 adds them
            if (g == 1) {
#pragma unroll
                for (int r = 0; r < FA_R; r++) {
                    float* o = qkv + r * 3 * FA_D + j;
                    o[0] = acc[0][r]; o[FA_D] = acc[1][r]; o[2 * FA_D] = acc[2][r];
                }
            }
            __syncthreads();
            if (g == 0) {
#pragma unroll
                for (int r = 0; r < FA_R; r++) {
                    float* o = qkv + r * 3 * FA_D + j;
                    o[0] += acc[0][r]; o[FA_D] += acc[1][r]; o[2 * FA_D] += acc[2][r];
                }
            }
        }
        __syncthreads();
        // ---- stage 3: masked multi-head attention; thread = (row = (env, agent slot), head h, half of the head dim) ----------------
        {
            const int half = tid & 1, h = (tid >> 1) & 3, r = tid >> 3, ev = r / NE;
            const int col = h * FA_HD + half * 8;
            const float* qr = qkv + r * 3 * FA_D + col;
            float qv[8];
#pragma unroll
            for (int c = 0; c < 8; c++) qv[c] = qr[c];
            float lg[NE], mx = -INFINITY;
#pragma unroll
            for (int jj = 0; jj < NE; jj++) {
                const float* kr = qkv + (ev * NE + jj) * 3 * FA_D + FA_D + col;
                float sdot = 0.f;
#pragma unroll
                for (int c = 0; c < 8; c++) sdot = fmaf(qv[c], kr[c], sdot);
                sdot += __shfl_xor_sync(0xffffffffu, sdot, 1);
                lg[jj] = sdot * inv_scale;
                if (!s_obs[r][jj]) mx = fmaxf(mx, lg[jj]);
            }
            float ssum = 0.f, o[8];
#pragma unroll
            for (int c = 0; c < 8; c++) o[c] = 0.f;
#pragma unroll
            for (int jj = 0; jj < NE; jj++) {
                const float ew = s_obs[r][jj] ? 0.f : expf(lg[jj] - mx);
                ssum += ew;
                const float* vr = qkv + (ev * NE + jj) * 3 * FA_D + 2 * FA_D + col;
#pragma unroll
                for (int c = 0; c < 8; c++) o[c] = fmaf(ew, vr[c], o[c]);
            }
            const float rn = ssum > 0.f ? 1.f / ssum : 0.f;        // all-masked row -> zeros (attention.py:58-60)
#pragma unroll
            for (int c = 0; c < 8; c++) attt[(col + c) * FA_R + r] = o[c] * rn;
        }
        __syncthreads();
        // ---- stage 4: x2 = relu(rowmask(out_trans(att)))  [R, 64] (rows of inactive agents zero, entity_ff_agent.py:44-46) -------
        {
            float acc[8];
#pragma unroll
            for (int r = 0; r < 8; r++) acc[r] = bias[FA_D + j];
#pragma unroll 4
            for (int k = 0; k < FA_D; k++) {
                const float w = woutt[k * FA_LD1 + j];
                const float4 x0 = *reinterpret_cast<const float4*>(attt + k * FA_R + 8 * g);
                const float4 x1 = *reinterpret_cast<const float4*>(attt + k * FA_R + 8 * g + 4);
                acc[0] = fmaf(x0.x, w, acc[0]); acc[1] = fmaf(x0.y, w, acc[1]); acc[2] = fmaf(x0.z, w, acc[2]); acc[3] = fmaf(x0.w, w, acc[3]);
                acc[4] = fmaf(x1.x, w, acc[4]); acc[5] = fmaf(x1.y, w, acc[5]); acc[6] = fmaf(x1.z, w, acc[6]); acc[7] = fmaf(x1.w, w, acc[7]);
            }
            float v[8];
#pragma unroll
            for (int r = 0; r < 8; r++) v[r] = s_em[8 * g + r] ? 0.f : fmaxf(acc[r], 0.f);
            float* o = x2t + j * FA_R + 8 * g;
            *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
        __syncthreads();
        // ---- stage 5: q = rowmask(fc2(x2))  [na, A] per environment -----------------------------------------------------------
        for (int f = tid; f < EPI * na * A; f += FA_THREADS) {
            const int ev = f / (na * A), i = (f / A) % na, ac = f % A, e = e0 + ev;
            if (e >= a.E) continue;
            float sq = bias[2 * FA_D + ac];
            const float* w = w2s + ac * FA_D;
            const int r = ev * NE + i;
#pragma unroll 8
            for (int k = 0; k < FA_D; k++) sq = fmaf(x2t[k * FA_R + r], w[k], sq);
            sq = s_em[r] ? 0.f : sq;
            a.q[((size_t)e * na + i) * A + ac] = sq;
            s_q[r][ac] = sq;
        }
        __syncthreads();
        // ---- stage 6 (optional): epsilon-greedy selection, one thread per (env, agent) ---------------------------------------
        if (a.avail && tid < EPI * na) {
            const int ev = tid / na, i = tid - ev * na, e = e0 + ev, r = ev * NE + i;
            if (e < a.E && !(a.est_flags && !(a.est_flags[e] & 2))) {
                const int32_t* av = a.avail + (((size_t)e * a.T + a.t) * na + i) * A;
                int best = 0, n_avail = 0;
                float bv = -INFINITY;
                bool have = false;
                for (int k = 0; k < A; k++) {               // unavailable -> -inf, FIRST maximum wins
                    const bool ok = av[k] != 0;
                    n_avail += ok;
                    const float v = ok ? s_q[r][k] : -INFINITY;
                    if (!have || v > bv) { best = k; bv = v; have = true; }
                }
                int pick = best;
                const float eps = a.eps_dev ? __ldg(a.eps_dev) : a.eps_host;
                const size_t idx = (size_t)e * na + i;
                if (a.u_pick && eps > 0.f && a.u_pick[idx] < eps && n_avail > 0) {
                    const int target = min((int)(a.u_act[idx] * (float)n_avail), n_avail - 1);
                    int seen = 0;
                    for (int k = 0; k < A; k++)
                        if (av[k] != 0) { if (seen == target) { pick = k; break; } seen++; }
                }
                a.actions_out[((size_t)e * a.T + a.t) * na + i] = pick;
            }
        }
        // (the next iteration's stage 0 does not touch s_q / s_em before its own barrier; s_em is rewritten after it)
        __syncthreads();
    }
}

extern "C" int refil_ff_agent_act_supported(int n_entities, int n_agents, int input_dim, int embed_dim, int n_heads,
                                            int n_actions) {
    return (embed_dim == FA_D && n_heads == FA_H && n_entities >= 1 && n_entities <= FA_NE_MAX && n_agents >= 1 &&
            n_agents <= n_entities && input_dim >= 1 && input_dim <= FA_EIN_MAX && n_actions >= 1 && n_actions <= FA_A_MAX) ? 1 : 0;
}

extern "C" int refil_ff_agent_act(const float* entities, int entity_dim, const long long* actions, int n_actions_onehot,
                                  const uint8_t* obs_mask, int mask_rows, const uint8_t* entity_mask, const float* fc1_w,
                                  const float* fc1_b, const float* in_trans_w, const float* out_trans_w,
                                  const float* out_trans_b, const float* fc2_w, const float* fc2_b, float* q, int n_envs,
                                  int T, int t, int n_entities, int n_agents, int n_actions, const int32_t* avail,
                                  const float* u_pick, const float* u_act, const int32_t* est_flags, float epsilon,
                                  const float* epsilon_dev, long long* actions_out, cudaStream_t stream) {
    const int ein = entity_dim + (actions ? n_actions_onehot : 0);
    REFIL_CHECK_ARG(entities && obs_mask && entity_mask && fc1_w && fc1_b && in_trans_w && out_trans_w && out_trans_b && fc2_w &&
                    fc2_b && q, "ff_agent_act: null pointer");
    REFIL_CHECK_ARG(refil_ff_agent_act_supported(n_entities, n_agents, ein, FA_D, FA_H, n_actions),
                    "ff_agent_act: unsupported shape (ne=%d na=%d ein=%d A=%d; d=64, 4 heads, <= 8 entities)", n_entities, n_agents, ein,
                    n_actions);
    REFIL_CHECK_ARG(n_envs > 0 && T > 0 && t >= 0 && t < T && (mask_rows == n_entities || mask_rows == n_agents),
                    "ff_agent_act: bad n_envs=%d T=%d t=%d mask_rows=%d", n_envs, T, t, mask_rows);
    REFIL_CHECK_ARG(!avail || actions_out, "ff_agent_act: selection (avail given) needs actions_out");
    REFIL_CHECK_ARG(!avail || ((epsilon <= 0.f && !epsilon_dev) || (u_pick && u_act)), "ff_agent_act: epsilon > 0 needs u_pick and u_act");
    FfActArgs a{entities, entity_dim, actions, n_actions_onehot, obs_mask, mask_rows, entity_mask, fc1_w, fc1_b, in_trans_w,
                out_trans_w, out_trans_b, fc2_w, fc2_b, q, n_envs, T, t, n_entities, n_agents, ein, n_actions,
                avail, u_pick, u_act, est_flags, epsilon, epsilon_dev, actions_out};
    const size_t smem = sizeof(float) * (FA_EIN_MAX * FA_LD1 + FA_D * FA_LD1 + FA_A_MAX * FA_D + 2 * FA_D + FA_A_MAX +
                                         FA_EIN_MAX * FA_R + FA_D * FA_R + FA_R * 3 * FA_D + 2 * FA_D * FA_R);
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(ff_agent_act_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ff_agent_act_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            refil_set_error("ff_agent_act: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
            return REFIL_ERR_CUDA;
        }
        attr = true;
    }
    const int NE = n_entities <= 4 ? 4 : 8;           // entity slots per environment in the 16-row stage tile
    int grid = refil_cdiv(n_envs, FA_R / NE);
    const int cap = 3 * refil_num_sms();              // three CTAs per SM (54 KB of shared memory, <= 170 registers per thread)
    if (grid > cap) grid = cap;
    if (NE == 4) ff_agent_act_kernel<4><<<grid, FA_THREADS, smem, stream>>>(a);
    else ff_agent_act_kernel<8><<<grid, FA_THREADS, smem, stream>>>(a);
    REFIL_CHECK_LAUNCH("ff_agent_act");
    return REFIL_OK;
}
