// C-ABI entry points for the dense layers of the learner (kernels K1/K3/K6 GEMM parts).
#include "gemm.cuh"

using namespace rg;

static RowMask make_rm(const uint8_t* em, int na, int ne, int mper) { return RowMask{em, na > 0 ? na : 1, ne, mper > 0 ? mper : 1}; }

// ---------------------------------------------------------------------------------------------
// forward:  C[M,N] = [rowmask][relu](A[M,K] W[N,K]^T + bias)
// ---------------------------------------------------------------------------------------------
extern "C" int refil_linear_fwd(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc,
                                int M, int N, int K, int relu, const uint8_t* row_entity_mask, int na, int ne,
                                int rows_per_copy, cudaStream_t stream) {
    REFIL_CHECK_ARG(A && W && C && M > 0 && N > 0 && K > 0, "linear_fwd: bad arguments (M=%d N=%d K=%d)", M, N, K);
    MatPlain a{A, lda, M, K}, w{W, ldw, N, K};
    EpiStore e{C, ldc, bias, relu, make_rm(row_entity_mask, na, ne, rows_per_copy)};
    launch_sgemm<MatPlain, true, MatPlain, true, EpiStore>(a, w, e, M, N, K, 1, stream);
    REFIL_CHECK_LAUNCH("linear_fwd");
    return REFIL_OK;
}

// x1 = relu([entities | onehot(last_action)] W1^T + b1)  -- the concat of entity_controller.py:14-27 /
// q_learner.py:52-60 is virtual (column gather of W1), never written to HBM
extern "C" int refil_embed_fwd(const float* entities, int ed, const int32_t* last_action, int n_actions,
                               const float* W, const float* bias, float* C, int M, int N, int relu,
                               cudaStream_t stream) {
    REFIL_CHECK_ARG(entities && W && C && M > 0 && N > 0 && ed > 0, "embed_fwd: bad arguments");
    int A = last_action ? n_actions : 0;
    MatConcat a{entities, ed, last_action, A, 0, M, ed + A};
    MatPlain w{W, ed + A, N, ed + A};
    EpiStore e{C, N, bias, relu, make_rm(nullptr, 1, 1, 1)};
    launch_sgemm<MatConcat, true, MatPlain, true, EpiStore>(a, w, e, M, N, ed + A, 1, stream);
    REFIL_CHECK_LAUNCH("embed_fwd");
    return REFIL_OK;
}

// ---------------------------------------------------------------------------------------------
// backward-data:  dA[M,K] = g(dC)[M,N] W[N,K],  g = relu'(relu_y) and/or row mask
// ---------------------------------------------------------------------------------------------
extern "C" int refil_linear_bwd_data(const float* dC, int lddc, const float* relu_y, int ldy,
                                     const uint8_t* row_entity_mask, int na, int ne, int rows_per_copy, const float* W,
                                     int ldw, float* dA, int ldda, int M, int N, int K, cudaStream_t stream) {
    REFIL_CHECK_ARG(dC && W && dA && M > 0 && N > 0 && K > 0, "linear_bwd_data: bad arguments");
    MatGrad g{dC, lddc, relu_y, ldy, make_rm(row_entity_mask, na, ne, rows_per_copy), M, N};
    MatPlain w{W, ldw, N, K};
    EpiStore e{dA, ldda, nullptr, 0, make_rm(nullptr, 1, 1, 1)};
    // output [M, K]; reduction over N
    launch_sgemm<MatGrad, true, MatPlain, false, EpiStore>(g, w, e, M, K, N, 1, stream);
    REFIL_CHECK_LAUNCH("linear_bwd_data");
    return REFIL_OK;
}

// ---------------------------------------------------------------------------------------------
// backward-weight:  dW[N,K] += g(dC)[M,N]^T A[M,K] ;  db[N] += colsum g(dC)
// ---------------------------------------------------------------------------------------------
template <class Mat>
__global__ void colsum_kernel(Mat g, float* __restrict__ db, int M, int N, int rows_per_block) {
    __shared__ float red[8][33];
    int n = blockIdx.x * 32 + threadIdx.x;
    int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
    float s = 0.f;
    if (n < N)
        for (int r = r0 + threadIdx.y; r < r1; r += 8) s += g.at(r, n);
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) t += red[i][threadIdx.x];
        atomicAdd(db + n, t);
    }
}

static int weight_splits(int M, int tiles) {
    int target = 2 * refil_num_sms();
    int s = (target + tiles - 1) / tiles;
    int maxs = (M + 4 * BK - 1) / (4 * BK);
    if (s > maxs) s = maxs;
    return s < 1 ? 1 : s;
}

template <class AM>
static int bwd_weight_impl(const MatGrad& g, const AM& a, float* dW, int ldw, int ncol_w, float* db_from_ones,
                           int M, int N, int Kcols, cudaStream_t stream) {
    EpiAtomic e{dW, ldw, ncol_w, db_from_ones};
    int tile_n = Kcols > 64 ? 128 : (Kcols > 32 ? 64 : (Kcols > 16 ? 32 : 16));
    int tiles = refil_cdiv(N, 128) * refil_cdiv(Kcols, tile_n);
    launch_sgemm<MatGrad, false, AM, false, EpiAtomic>(g, a, e, N, Kcols, M, weight_splits(M, tiles), stream);
    return 0;
}

extern "C" int refil_linear_bwd_weight(const float* dC, int lddc, const float* relu_y, int ldy,
                                       const uint8_t* row_entity_mask, int na, int ne, int rows_per_copy,
                                       const float* A, int lda, float* dW, int ldw, float* db, int M, int N, int K,
                                       cudaStream_t stream) {
    REFIL_CHECK_ARG(dC && A && dW && M > 0 && N > 0 && K > 0, "linear_bwd_weight: bad arguments");
    MatGrad g{dC, lddc, relu_y, ldy, make_rm(row_entity_mask, na, ne, rows_per_copy), M, N};
    MatPlain a{A, lda, M, K};
    bwd_weight_impl(g, a, dW, ldw, K, nullptr, M, N, K, stream);
    REFIL_CHECK_LAUNCH("linear_bwd_weight");
    if (db) {
        int nby = refil_cdiv(M, 2048);
        if (nby > 4 * refil_num_sms()) nby = 4 * refil_num_sms();
        int rpb = refil_cdiv(M, nby);
        colsum_kernel<MatGrad><<<dim3(refil_cdiv(N, 32), refil_cdiv(M, rpb)), dim3(32, 8), 0, stream>>>(g, db, M, N, rpb);
        REFIL_CHECK_LAUNCH("linear_bwd_weight(colsum)");
    }
    return REFIL_OK;
}

// fc1 of agent / hypernets: A = [entities | onehot | 1] so that db is the last column of the same GEMM
extern "C" int refil_embed_bwd_weight(const float* dC, int lddc, const float* relu_y, int ldy, const float* entities,
                                      int ed, const int32_t* last_action, int n_actions, float* dW, float* db, int M,
                                      int N, cudaStream_t stream) {
    REFIL_CHECK_ARG(dC && entities && dW && db && M > 0 && N > 0, "embed_bwd_weight: bad arguments");
    int A = last_action ? n_actions : 0;
    MatGrad g{dC, lddc, relu_y, ldy, make_rm(nullptr, 1, 1, 1), M, N};
    MatConcat a{entities, ed, last_action, A, 1, M, ed + A + 1};
    bwd_weight_impl(g, a, dW, ed + A, ed + A, db, M, N, ed + A + 1, stream);
    REFIL_CHECK_LAUNCH("embed_bwd_weight");
    return REFIL_OK;
}

// GRU recurrent weights: dW_hh[3r, r] += dGH^T H_prev where H_prev(row at t) = HS(row at t-1), zero at t=0
extern "C" int refil_gru_bwd_weight_hh(const float* dGH, const float* HS, int n_agents, int T, float* dWhh,
                                       float* dbhh, int M, int r, cudaStream_t stream) {
    REFIL_CHECK_ARG(dGH && HS && dWhh && M > 0 && r > 0, "gru_bwd_weight_hh: bad arguments");
    MatGrad g{dGH, 3 * r, nullptr, 0, make_rm(nullptr, 1, 1, 1), M, 3 * r};
    MatPrevT a{HS, r, n_agents, T, M, r};
    bwd_weight_impl(g, a, dWhh, r, r, nullptr, M, 3 * r, r, stream);
    REFIL_CHECK_LAUNCH("gru_bwd_weight_hh");
    if (dbhh) {
        int nby = refil_cdiv(M, 2048);
        int rpb = refil_cdiv(M, nby);
        colsum_kernel<MatGrad><<<dim3(refil_cdiv(3 * r, 32), refil_cdiv(M, rpb)), dim3(32, 8), 0, stream>>>(g, dbhh, M, 3 * r, rpb);
        REFIL_CHECK_LAUNCH("gru_bwd_weight_hh(colsum)");
    }
    return REFIL_OK;
}

// last-action index per (b, t, entity): actions[b, t-1, e] for agents at t > 0, else -1
// (entity_controller.py:16-24: one-hot rows are zero at t=0 and for non-agent entities)
__global__ void last_action_kernel(const long long* __restrict__ actions, int32_t* __restrict__ la, int B, int T,
                                   int na, int ne) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * T * ne) return;
    int e = idx % ne, t = (idx / ne) % T, b = idx / (ne * T);
    la[idx] = (t > 0 && e < na) ? (int32_t)actions[((size_t)b * T + (t - 1)) * na + e] : -1;
}

extern "C" int refil_last_action_index(const long long* actions, int32_t* la, int B, int T, int n_agents,
                                       int n_entities, cudaStream_t stream) {
    REFIL_CHECK_ARG(actions && la && B > 0 && T > 0, "last_action_index: bad arguments");
    int n = B * T * n_entities;
    last_action_kernel<<<refil_cdiv(n, 256), 256, 0, stream>>>(actions, la, B, T, n_agents, n_entities);
    REFIL_CHECK_LAUNCH("last_action_index");
    return REFIL_OK;
}

// Materialise the padded network input once per step: out[r, :] = [entities[r, :ed] | onehot(last_action[r], A) | 0...]
// (row width Kp = multiple of 32) so that fc1 of all ten networks (agent, 4 hypernets, and their targets) runs as a plain
// K-major tensor-core GEMM over the same buffer (controllers/entity_controller.py:14-27, learners/q_learner.py:52-60).
__global__ void pack_inputs_kernel(const float* __restrict__ ents, const int32_t* __restrict__ la, float* __restrict__ out,
                                   long long R, int ed, int A, int Kp) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * Kp) return;
    long long r = idx / Kp;
    int c = (int)(idx - r * Kp);
    float v = 0.f;
    if (c < ed) v = __ldg(ents + r * ed + c);
    else if (la && c < ed + A) v = (la[r] == c - ed) ? 1.f : 0.f;
    out[idx] = v;
}

extern "C" int refil_pack_inputs(const float* entities, int ed, const int32_t* last_action, int n_actions, float* out,
                                 long long rows, int padded_width, cudaStream_t stream) {
    REFIL_CHECK_ARG(entities && out && rows > 0 && ed > 0, "pack_inputs: bad arguments");
    REFIL_CHECK_ARG(padded_width >= ed + (last_action ? n_actions : 0), "pack_inputs: padded width %d too small", padded_width);
    long long n = rows * padded_width;
    pack_inputs_kernel<<<refil_cdiv(n, 256), 256, 0, stream>>>(entities, last_action, out, rows, ed, n_actions, padded_width);
    REFIL_CHECK_LAUNCH("pack_inputs");
    return REFIL_OK;
}
