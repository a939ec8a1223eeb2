/* librefil_b200.so -- C ABI of the B200-native REFIL hot path.
 *
 * The reference (shariqiqbal2810/REFIL) is pure Python and has NO FFI of its own: its seams are Python classes
 * (SURVEY.md section 8b).  Each entry point below therefore cites the reference Python code whose arithmetic it
 * replaces (paths relative to /root/reference/src); the Python host classes under refil_b200/ mirror the reference
 * class API (controllers.BasicMAC/EntityMAC, learners.QLearner, envs.GroupMatching, ...) and call these through ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to a contiguous buffer owned by the caller (PyTorch tensors in practice),
 *     16-byte aligned; nothing is allocated or retained by the library;
 *   - all work is enqueued on `stream` and returns immediately (no internal synchronisation);
 *   - return value 0 = ok, <0 = error (REFIL_ERR_*), text via refil_last_error() (thread local);
 *   - masks: 1 = masked / inactive, dtype uint8;  "N" is the number of (b, t) rows, N = B*T;
 *   - row stacks [C, N, na, *] put the mask copy (plain / within / interact) outermost.
 *
 * refil_b200/_lib.py parses THIS file to build the ctypes signatures, so the declarations here are the single
 * source of truth (one declaration per statement, `int refil_xxx(...)` form).
 */
#ifndef REFIL_B200_H
#define REFIL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define REFIL_OK 0
#define REFIL_ERR_ARG (-1)
#define REFIL_ERR_CUDA (-2)
#define REFIL_ERR_UNSUPPORTED (-3)

/* attention mask modes (bit field, per copy) -- modules/agents/entity_rnn_agent.py:79-117, mixers/flex_qmix.py:43-46 */
#define REFIL_ATTN_PART_WITHIN 1   /* masked unless same random group and both present at t=0            */
#define REFIL_ATTN_PART_INTERACT 2 /* masked iff    same random group and both present at t=0            */
#define REFIL_ATTN_ACTIVE0 4       /* OR inactive0[i] | inactive0[j]   (mixer "noobs" masks)              */
#define REFIL_ATTN_DEFAULT 8       /* OR entity_mask[n,i] | entity_mask[n,j]  (hypernet default mask)     */

/* mixer kinds -- modules/mixers/flex_qmix.py:60 (FlexQMixer), :124 (LinearFlexQMixer), vdn.py:5 (VDNMixer) */
#define REFIL_MIX_FLEX 0
#define REFIL_MIX_LIN 1
#define REFIL_MIX_VDN 2

const char* refil_last_error(void);
int refil_abi_version(void);
int refil_device_sm_count(void);

/* ---- Group Matching environment (envs/group_matching/group_matching.py) --------------------------------------
 * State per env e of E (env-minor, so thread==env accesses coalesce):
 *   mt_key u32[624][E], mt_pos i32[E]  numpy legacy RandomState (MT19937)        group_matching.py:114-118
 *   loc i32[na][E], grp u32[ng][E] (member bitmasks), est i32[4][E] (prev_matches, t, flags, ep_len), ep_ret f64[E]
 * Rollout tensors use the EpisodeBatch layout [B, T, ...] of run.py:178-196; row of env e = env_offset + e. */
int refil_gm_env_seed(uint32_t* mt_key, int32_t* mt_pos, const uint32_t* seeds, int n_envs, cudaStream_t stream);
/* reset(): group_matching.py:91-106 (+ get_entities :66-73, get_masks :55-64, get_avail_actions :78-80 at ts=0) */
int refil_gm_env_reset(uint32_t* mt_key, int32_t* mt_pos, int32_t* loc, uint32_t* grp, int32_t* est, double* ep_ret,
                       float* entities, uint8_t* obs_mask, uint8_t* entity_mask, uint8_t* gt_mask,
                       int32_t* avail_actions, long long* filled, int n_envs, int n_agents, int n_entities,
                       int n_states, int n_groups, int fixed_scen, int episode_limit, int T, int env_offset,
                       cudaStream_t stream);
/* step(actions[:, ts]): group_matching.py:19-53 + the post/pre-transition writes of runners/parallel_runner.py:140-197 */
int refil_gm_env_step(uint32_t* mt_key, int32_t* mt_pos, int32_t* loc, uint32_t* grp, int32_t* est, double* ep_ret,
                      const long long* actions, float* entities, uint8_t* obs_mask, uint8_t* entity_mask,
                      uint8_t* gt_mask /* null: not rewritten after reset (ParallelRunner); else every step
                      (runners/episode_runner.py:66-67) */, int32_t* avail_actions, float* reward, uint8_t* terminated,
                      long long* filled,
                      unsigned long long* step_counter, int n_envs, int n_agents, int n_entities, int n_states,
                      int n_groups, double rand_trans, int episode_limit, int T, int ts, int env_offset,
                      cudaStream_t stream);

/* ---- acting: components/action_selectors.py:45-63 (EpsilonGreedyActionSelector.select_action) ---------------- */
int refil_select_actions(const float* q, long long q_stride_b, const int32_t* avail, long long avail_stride_b,
                         const float* u_pick, const float* u_act, const int32_t* est_flags, float epsilon,
                         const float* epsilon_dev /* optional device scalar that overrides `epsilon` (graph-captured rollouts) */,
                         long long* actions_out, long long out_stride_b, int B, int n_agents, int n_actions,
                         cudaStream_t stream);

/* ---- fused acting step of the feed-forward entity-attention agent: EntityAttentionFFAgent.forward
 *      (modules/agents/entity_ff_agent.py:30-57) on the inputs of EntityMAC._build_inputs (controllers/entity_controller.py:11-30),
 *      read in place from the EpisodeBatch tensors at timestep t ([E, T, ...] layouts).  q [E, na, A].  actions: null, or the
 *      [E, T, na, 1] tensor whose step t-1 is appended as a one-hot (entity_last_action).  mask_rows = ne (obs_mask) or na (gt_mask
 *      with gt_obs_mask).  Supported: embed 64, 4 heads, <= 8 entities, <= 32 input features, <= 16 actions. */
int refil_ff_agent_act_supported(int n_entities, int n_agents, int input_dim, int embed_dim, int n_heads, int n_actions);
int refil_ff_agent_act(const float* entities, int entity_dim, const long long* actions, int n_actions_onehot,
                       const uint8_t* obs_mask, int mask_rows, const uint8_t* entity_mask, const float* fc1_w,
                       const float* fc1_b, const float* in_trans_w, const float* out_trans_w, const float* out_trans_b,
                       const float* fc2_w, const float* fc2_b, float* q, int n_envs, int T, int t, int n_entities,
                       int n_agents, int n_actions,
                       const int32_t* avail /* optional fused epsilon-greedy selection (action_selectors.py:45-63): the
                       [E, T, na, A] availability tensor, read at t; null = utilities only */,
                       const float* u_pick, const float* u_act, const int32_t* est_flags, float epsilon,
                       const float* epsilon_dev, long long* actions_out /* [E, T, na, 1], written at t */,
                       cudaStream_t stream);

/* ---- dense layers (nn.Linear of modules/layers/attention.py:21-22, agents/entity_rnn_agent.py:12,23-25,
 *      mixers/flex_qmix.py:29,39).  C = [rowmask][relu](A W^T + b); row r of a [C, N, na] stack is zeroed when
 *      row_entity_mask[n, a] != 0 (attention.py:66-67, entity_rnn_agent.py:60, flex_qmix.py:50). */
int refil_linear_fwd(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc, int M,
                     int N, int K, int relu, const uint8_t* row_entity_mask, int na, int ne, int rows_per_copy,
                     cudaStream_t stream);
/* fc1 over the virtual concat [entities | onehot(last action)] of controllers/entity_controller.py:14-27 and
 * learners/q_learner.py:52-60 (last_action[row] = -1 -> zero one-hot) */
int refil_embed_fwd(const float* entities, int ed, const int32_t* last_action, int n_actions, const float* W,
                    const float* bias, float* C, int M, int N, int relu, cudaStream_t stream);
int refil_linear_bwd_data(const float* dC, int lddc, const float* relu_y, int ldy, const uint8_t* row_entity_mask,
                          int na, int ne, int rows_per_copy, const float* W, int ldw, float* dA, int ldda, int M,
                          int N, int K, cudaStream_t stream);
int refil_linear_bwd_weight(const float* dC, int lddc, const float* relu_y, int ldy, const uint8_t* row_entity_mask,
                            int na, int ne, int rows_per_copy, const float* A, int lda, float* dW, int ldw, float* db,
                            int M, int N, int K, cudaStream_t stream);
int refil_embed_bwd_weight(const float* dC, int lddc, const float* relu_y, int ldy, const float* entities, int ed,
                           const int32_t* last_action, int n_actions, float* dW, float* db, int M, int N,
                           cudaStream_t stream);
int refil_gru_bwd_weight_hh(const float* dGH, const float* HS, int n_agents, int T, float* dWhh, float* dbhh, int M,
                            int r, cudaStream_t stream);
/* padded network input [rows, padded_width] = [entities | onehot(last action) | 0] shared by fc1 of every network */
int refil_pack_inputs(const float* entities, int ed, const int32_t* last_action, int n_actions, float* out,
                      long long rows, int padded_width, cudaStream_t stream);
int refil_last_action_index(const long long* actions, int32_t* la, int B, int T, int n_agents, int n_entities,
                            cudaStream_t stream);

/* ---- the same dense layers on the tcgen05 tensor cores (3xTF32 split, fp32 accumulate in TMEM; fp32-grade accuracy):
 *      C[M,N] = [rowmask_c][relu]( g(A)[M,K] B[N,K]^T + bias ),  B[j,i] = B[j*b_stride_n + i*b_stride_k],
 *      g = relu'(relu_y) and/or row mask on A (backward-data use).  refil_tc_gemm_supported() != 0 iff the shape
 *      can run here (K % 32 == 0, N % 32 == 0, ...); otherwise use refil_linear_fwd / refil_linear_bwd_data.
 *      Operand path: the A operand of tcgen05.mma lives in tensor memory (TMA-fed, "ts"); REFIL_TC_MODE=ss in the environment
 *      selects the all-shared-memory predecessor for A/B comparisons.
 *      refil_tc_gemm_k_slices() > 1: the reduction is too long for a resident weight tile and is cut into k-slices whose
 *      partial tiles are reduce-added into a zeroed C -- only for a linear epilogue (no bias / relu / output row mask). */
/* launch-shape hint for the dense kernel: m-tiles every CTA should at least get (0 = default / REFIL_TC_MIN_TILES); returns the
 * previous override.  Small on a latency-critical chain (more CTAs, shorter kernel), large beside it (fewer SMs taken). */
int refil_tc_set_min_tiles(int min_tiles);
int refil_tc_gemm_supported(int M, int N, int K);
int refil_tc_gemm_k_slices(int N, int K);
int refil_tc_gemm_tn(const float* A, long long lda, const float* relu_y, long long ldy,
                     const uint8_t* a_row_entity_mask, int a_na, int a_ne, int a_rows_per_copy, const float* B,
                     long long b_stride_n, long long b_stride_k,
                     int b_k_valid /* B is taken as zero for reduction indices >= b_k_valid (0: K): a weight whose true fan-in is
                     not a multiple of 32 (fc1: 53, action head: 14) is read in place, no padded copy */,
                     const float* bias, int relu, const uint8_t* c_row_entity_mask, int c_na, int c_ne, int c_rows_per_copy,
                     float* C, long long ldc, int M, int N, int K, cudaStream_t stream);

/* Grouped launch: up to 8 independent problems of ONE (N, K) geometry -- the same layer of several networks (the eight
 * hypernetworks of the online and target mixers, the agent and its target) -- as a single kernel launch (blockIdx.y = problem).
 * Field meaning as the arguments of refil_tc_gemm_tn. */
typedef struct RefilGemmDesc {
    const float* A; long long lda;
    const float* relu_y; long long ldy;
    const uint8_t* a_row_entity_mask; int a_na, a_ne, a_rows_per_copy;
    const float* B; long long b_stride_n, b_stride_k; int b_k_valid;
    const float* bias; int relu;
    const uint8_t* c_row_entity_mask; int c_na, c_ne, c_rows_per_copy;
    float* C; long long ldc;
    int M;
    int row_group, row_group_stride;   /* > 0: row m of A and of C is physical row (m / row_group) * row_group_stride + m % row_group */
    int accumulate;                    /* C += A B^T instead of C = (linear epilogue only) */
    int n_cols;                        /* > 0: this problem has n_cols output columns instead of the group's N (same reduction length):
                                          in_trans as ONE launch of K|V for all entity rows (2d columns) + Q for the agent rows (d) */
    int k_len;                         /* > 0: this problem reduces over k_len instead of the group's K (a multiple of the same k-slice) */
} RefilGemmDesc;
int refil_tc_gemm_tn_group(const RefilGemmDesc* descs, int n_problems, int N, int K, cudaStream_t stream);

/* The same product on ROW GROUPS: only the first group_rows of every group_stride rows of A are read and only those rows of C are
 * written -- the agent rows of an [N, n_entities, d] entity tensor.  The attention layer of the reference computes queries for all
 * entities and keeps the agents' (modules/layers/attention.py:46-48, `query = query[:n_queries]` right after in_trans): here the Q
 * third of in_trans runs on the agent rows alone.  accumulate: C += (the matching piece of the backward). */
int refil_tc_gemm_tn_rows(const float* A, long long lda, const float* B, long long b_stride_n, long long b_stride_k,
                          float* C, long long ldc, int n_groups, int group_rows, int group_stride, int accumulate,
                          int N, int K, cudaStream_t stream);

/* weight gradient on the tensor cores: dW[P,Q] += g(X)[M,P]^T Y[M,Q]; db[P] += colsum g(X) (db may be null).
 * y_shift_rows > 0: Y row m is read from row m - y_shift_rows and is zero where (m / y_shift_rows) % y_period == 0
 * (the h_{t-1} view of the GRU state stack for dW_hh: shift = n_agents, period = T) */
int refil_tc_wgrad_supported(int M, int P, int Q);
int refil_tc_gemm_wgrad(const float* X, long long ldx, const float* relu_y, long long ldy,
                        const uint8_t* x_row_entity_mask, int na, int ne, int rows_per_copy, const float* Y,
                        long long ldyy, int y_shift_rows, int y_period, float* dW, long long lddw,
                        int q_valid /* dW has q_valid <= Q columns (0: Q); the rest of Y is zero padding */, float* db,
                        int M, int P, int Q, cudaStream_t stream);

/* grouped weight gradients: up to 8 problems of one (P, Q) geometry in a single launch; fields as refil_tc_gemm_wgrad */
typedef struct RefilWgradDesc {
    const float* X; long long ldx;
    const float* relu_y; long long ldy;
    const uint8_t* x_row_entity_mask; int na, ne, rows_per_copy;
    const float* Y; long long ldyy; int y_shift_rows, y_period;
    float* dW; long long lddw; int q_valid;
    float* db;
    int M;
    int p_cols;                        /* > 0: this problem has p_cols columns of X (rows of dW) instead of the group's P */
    int row_group, row_group_stride;   /* > 0: row m of X and Y is physical row (m / row_group) * row_group_stride + m % row_group */
} RefilWgradDesc;
int refil_tc_gemm_wgrad_group(const RefilWgradDesc* descs, int n_problems, int P, int Q, cudaStream_t stream);

/* ---- masked multi-head attention over entities: modules/layers/attention.py:43-64 with the mask algebra of
 *      agents/entity_rnn_agent.py:79-124 resolved on the fly.  QKV [N, ne, 3d]; OUT / dOUT [C, N, nq, d].
 *      Instantiated for head dim in {8, 16, 32} and heads in {1, 2, 4, 8}, ne <= 32; other shapes return REFIL_ERR_UNSUPPORTED. */
int refil_masked_attn_fwd(const float* qkv, float* out, const uint8_t* mask0, const uint8_t* mask1,
                          const uint8_t* mask2, long long mask_stride0, long long mask_stride1,
                          long long mask_stride2, int mode0, int mode1, int mode2, const uint8_t* group_bits,
                          const uint8_t* entity_mask, int N, int T, int n_entities, int n_queries, int embed_dim,
                          int n_heads, int n_copies, cudaStream_t stream);
int refil_masked_attn_bwd(const float* qkv, const float* dout, float* dqkv, const uint8_t* mask0,
                          const uint8_t* mask1, const uint8_t* mask2, long long mask_stride0, long long mask_stride1,
                          long long mask_stride2, int mode0, int mode1, int mode2, const uint8_t* group_bits,
                          const uint8_t* entity_mask, int N, int T, int n_entities, int n_queries, int embed_dim,
                          int n_heads, int n_copies, cudaStream_t stream);

/* grouped attention: up to 8 problems of one geometry (N, T, ne, nq, d, heads) -- the attention of several networks over the same
 * (b, t) units -- in a single launch; fields as the arguments of refil_masked_attn_fwd / bwd (out: forward; dout, dqkv: backward) */
typedef struct RefilAttnDesc {
    const float* qkv; float* out; const float* dout; float* dqkv;
    const uint8_t* mask0; const uint8_t* mask1; const uint8_t* mask2;
    long long mask_stride0, mask_stride1, mask_stride2;
    int mode0, mode1, mode2;
    const uint8_t* group_bits; const uint8_t* entity_mask;
    int n_copies;
} RefilAttnDesc;
int refil_masked_attn_fwd_group(const RefilAttnDesc* descs, int n_problems, int N, int T, int n_entities, int n_queries,
                                int embed_dim, int n_heads, cudaStream_t stream);
int refil_masked_attn_bwd_group(const RefilAttnDesc* descs, int n_problems, int N, int T, int n_entities, int n_queries,
                                int embed_dim, int n_heads, cudaStream_t stream);

/* ---- EntityPoolingLayer (`pooling_type: mean | max`): modules/layers/attention.py:82-132, same mask interface as the attention
 *      kernels.  E [N, ne, d] = in_trans(x1); OUT / dOUT [C, N, nq, d]; dE [N, ne, d] (overwritten).  pool_type 0 = mean (masked
 *      entities count as zeros, divisor ne), 1 = max (zeros of masked entities take part; gradient to the first arg-max). */
int refil_entity_pool_fwd(const float* E, float* out, const uint8_t* mask0, const uint8_t* mask1, const uint8_t* mask2,
                          long long mask_stride0, long long mask_stride1, long long mask_stride2, int mode0, int mode1,
                          int mode2, const uint8_t* group_bits, const uint8_t* entity_mask, int N, int T, int n_entities,
                          int n_queries, int embed_dim, int n_copies, int pool_type, cudaStream_t stream);
int refil_entity_pool_bwd(const float* E, const float* dout, float* dE, const uint8_t* mask0, const uint8_t* mask1,
                          const uint8_t* mask2, long long mask_stride0, long long mask_stride1, long long mask_stride2,
                          int mode0, int mode1, int mode2, const uint8_t* group_bits, const uint8_t* entity_mask, int N,
                          int T, int n_entities, int n_queries, int embed_dim, int n_copies, int pool_type,
                          cudaStream_t stream);

/* ---- GRU scan: the `for t` loop of agents/entity_rnn_agent.py:51-55 (torch.nn.GRUCell, gates r,z,n) ------------
 * rows = (seq-batch, t, agent); GI [R, 3r] = x W_ih^T + b_ih; gates [R, 4r] (saved for BPTT, may be null). */
int refil_gru_scan_fwd(const float* GI, const float* Whh, const float* bhh, const float* h0, float* HS, float* gates,
                       int n_seq, int T, int n_agents, int r, cudaStream_t stream);
int refil_gru_scan_bwd(const float* dHS, const float* gates, const float* HS, const float* h0, const float* Whh,
                       float* dGI, float* dGH, int n_seq, int T, int n_agents, int r, cudaStream_t stream);

/* ---- mixers: mixers/flex_qmix.py:79-121 (flex), :136-172 (lin_flex), vdn.py:9-10 ------------------------------ */
int refil_mixer_fwd(int kind, const float* W1, const float* B1, const float* WF, const float* V, const float* q,
                    const float* qW, const float* qI, float* qtot, float* qtot_im, float* ingroup_out, int N,
                    int n_agents, int mixing_embed, int w1_copies, int imagine, int softmax_weights, int tanh_nonlin,
                    cudaStream_t stream);
int refil_mixer_bwd(int kind, const float* W1, const float* B1, const float* WF, const float* V, const float* q,
                    const float* qW, const float* qI, const float* g_plain, const float* g_im, float* dW1, float* dB1,
                    float* dWF, float* dV, float* dq, float* dqW, float* dqI, int N, int n_agents, int mixing_embed,
                    int w1_copies, int imagine, int softmax_weights, int tanh_nonlin, cudaStream_t stream);

/* ---- TD target / loss / optimiser: learners/q_learner.py:68-72,92-96,111-126,157-178 --------------------------- */
int refil_gather_chosen(const float* Q, const long long* actions, float* chosen, int copies, long long rows_per_copy,
                        int n_actions, cudaStream_t stream);
int refil_scatter_dq(const float* dchosen, const long long* actions, float* dQ, int copies, long long rows_per_copy,
                     int n_actions, int T, int n_agents, cudaStream_t stream);
int refil_target_max(const float* q_online, const float* q_target, const int32_t* avail, float* tgt,
                     long long* cur_max_actions, long long rows, int n_actions, int double_q, cudaStream_t stream);
/* stats f64[>=8]: 0 sum(mask) 1 sum((mask td)^2) 2 sum((mask td_im)^2) 3 sum|mask td| 4 sum(q_tot mask) 5 sum(targets mask) */
int refil_td_loss(const float* qtot, const float* qtot_im, const float* tgt_tot, const float* reward,
                  const uint8_t* terminated, const long long* filled, float* g_plain, float* g_im, float* targets_out,
                  double* stats, int B, int T, float gamma, float lmbda, cudaStream_t stream);
int refil_grad_sumsq(const float* grads, long long n_params, double* out, cudaStream_t stream);
/* tail[k] = (float) stats[k]: loss statistics ride behind the flat gradient so one all-reduce covers both */
int refil_pack_stats(const double* stats, float* tail, int n, cudaStream_t stream);
/* g <- g / *mask_sum; clip_grad_norm_(grad_clip) with norm = sqrt(*sumsq) / *mask_sum; torch.optim.RMSprop step */
int refil_clip_rmsprop_step(float* params, float* grads, float* square_avg, long long n_params, const float* mask_sum,
                            const double* sumsq, float* grad_norm_out, float grad_clip, float lr, float alpha,
                            float eps, float weight_decay, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* REFIL_B200_H */
