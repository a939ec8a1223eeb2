// GRU scan over the episode axis (kernel K4).
//
// Replaces the python `for t in range(ts): h = self.rnn(x3[:, t], h)` loop of
//   /root/reference/src/modules/agents/entity_rnn_agent.py:51-55
// (torch.nn.GRUCell, gate order r, z, n) by ONE persistent launch per direction: the input projection
// GI = x3 W_ih^T + b_ih is a dense GEMM done beforehand for all T; this kernel keeps W_hh resident in shared
// memory and walks t = 0..T-1 (forward) or T-1..0 (backward, BPTT) for a tile of sequences per CTA.
//
// Row layout: rows are (seq-batch cb, t, agent a) -> row = (cb*T + t)*na + a ; a "sequence" is (cb, a).
//   GI [R, 3r]  HS [R, r]  GATES [R, 4r] = (r | z | n | W_hn h + b_hn)   dHS [R, r]  dGI [R, 3r]  dGH [R, 3r]
// Thread mapping: 128 threads = (128/r) groups x r features; every thread owns feature j of 4 sequences, so the
// recurrent mat-vec runs 12 FMAs per 3 conflict-free weight loads + 1 broadcast 128-bit state load.
#include "common.cuh"

#define GRU_THREADS 128
#define GRU_SEQ 4

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// smem: wt [r][3r] (k-major transpose of weight_hh) | hbuf [2][r][S_TILE]
__global__ void __launch_bounds__(GRU_THREADS) gru_scan_fwd_kernel(const float* __restrict__ GI,
                                                                  const float* __restrict__ Whh,
                                                                  const float* __restrict__ bhh,
                                                                  const float* __restrict__ h0, float* __restrict__ HS,
                                                                  float* __restrict__ GATES, int n_seq, int T, int na,
                                                                  int r) {
    extern __shared__ __align__(16) float smem[];
    const int ngrp = GRU_THREADS / r, S_TILE = ngrp * GRU_SEQ;
    float* wt = smem;                 // [r][3r]
    float* hbuf = wt + 3 * r * r;     // [2][r][S_TILE]
    const int tid = threadIdx.x, j = tid % r, grp = tid / r;
    for (int idx = tid; idx < 3 * r * r; idx += GRU_THREADS) {
        int gj = idx / r, k = idx - gj * r;  // Whh[gj][k]
        wt[k * 3 * r + gj] = Whh[idx];
    }
    const int s0 = blockIdx.x * S_TILE + grp * GRU_SEQ;
    long long row0[GRU_SEQ];
    bool valid[GRU_SEQ];
    float h[GRU_SEQ];
#pragma unroll
    for (int q = 0; q < GRU_SEQ; q++) {
        int s = s0 + q;
        valid[q] = s < n_seq;
        int cb = valid[q] ? s / na : 0, a = valid[q] ? s - cb * na : 0;
        row0[q] = ((long long)cb * T) * na + a;
        h[q] = (valid[q] && h0) ? h0[(size_t)s * r + j] : 0.f;
        hbuf[(0 * r + j) * S_TILE + grp * GRU_SEQ + q] = h[q];
    }
    const float br = bhh[j], bz = bhh[r + j], bn = bhh[2 * r + j];
    __syncthreads();
    int cur = 0;
    for (int t = 0; t < T; t++) {
        float gi[3][GRU_SEQ];
#pragma unroll
        for (int q = 0; q < GRU_SEQ; q++) {
            const float* g = GI + (size_t)(row0[q] + (long long)t * na) * 3 * r;
#pragma unroll
            for (int gg = 0; gg < 3; gg++) gi[gg][q] = valid[q] ? __ldg(g + gg * r + j) : 0.f;
        }
        float ar[GRU_SEQ], az[GRU_SEQ], an[GRU_SEQ];
#pragma unroll
        for (int q = 0; q < GRU_SEQ; q++) { ar[q] = br; az[q] = bz; an[q] = bn; }
        const float* hb = hbuf + cur * r * S_TILE + grp * GRU_SEQ;
        for (int k = 0; k < r; k++) {
            const float wr = wt[k * 3 * r + j], wz = wt[k * 3 * r + r + j], wn = wt[k * 3 * r + 2 * r + j];
            const float4 hv = *reinterpret_cast<const float4*>(hb + k * S_TILE);
            ar[0] = fmaf(hv.x, wr, ar[0]); ar[1] = fmaf(hv.y, wr, ar[1]); ar[2] = fmaf(hv.z, wr, ar[2]); ar[3] = fmaf(hv.w, wr, ar[3]);
            az[0] = fmaf(hv.x, wz, az[0]); az[1] = fmaf(hv.y, wz, az[1]); az[2] = fmaf(hv.z, wz, az[2]); az[3] = fmaf(hv.w, wz, az[3]);
            an[0] = fmaf(hv.x, wn, an[0]); an[1] = fmaf(hv.y, wn, an[1]); an[2] = fmaf(hv.z, wn, an[2]); an[3] = fmaf(hv.w, wn, an[3]);
        }
        float* hnext = hbuf + (cur ^ 1) * r * S_TILE;
#pragma unroll
        for (int q = 0; q < GRU_SEQ; q++) {
            const float rg = sigmoidf_(gi[0][q] + ar[q]);
            const float zg = sigmoidf_(gi[1][q] + az[q]);
            const float ng = tanhf(gi[2][q] + rg * an[q]);
            const float hn = (h[q] - ng) * zg + ng;
            h[q] = hn;
            hnext[j * S_TILE + grp * GRU_SEQ + q] = hn;
            if (valid[q]) {
                const size_t row = (size_t)(row0[q] + (long long)t * na);
                HS[row * r + j] = hn;
                if (GATES) {
                    float* g = GATES + row * 4 * r;
                    g[j] = rg; g[r + j] = zg; g[2 * r + j] = ng; g[3 * r + j] = an[q];
                }
            }
        }
        __syncthreads();
        cur ^= 1;
    }
}

// smem: w [3r][r] (weight_hh as stored) | dgh [3r][S_TILE]
__global__ void __launch_bounds__(GRU_THREADS) gru_scan_bwd_kernel(const float* __restrict__ dHS,
                                                                  const float* __restrict__ GATES,
                                                                  const float* __restrict__ HS,
                                                                  const float* __restrict__ h0,
                                                                  const float* __restrict__ Whh, float* __restrict__ dGI,
                                                                  float* __restrict__ dGH, int n_seq, int T, int na,
                                                                  int r) {
    extern __shared__ __align__(16) float smem[];
    const int ngrp = GRU_THREADS / r, S_TILE = ngrp * GRU_SEQ;
    float* w = smem;               // [3r][r]
    float* sg = w + 3 * r * r;     // [3r][S_TILE]
    const int tid = threadIdx.x, j = tid % r, grp = tid / r;
    for (int idx = tid; idx < 3 * r * r; idx += GRU_THREADS) w[idx] = Whh[idx];
    const int s0 = blockIdx.x * S_TILE + grp * GRU_SEQ;
    long long row0[GRU_SEQ];
    bool valid[GRU_SEQ];
    float dh[GRU_SEQ];
#pragma unroll
    for (int q = 0; q < GRU_SEQ; q++) {
        int s = s0 + q;
        valid[q] = s < n_seq;
        int cb = valid[q] ? s / na : 0, a = valid[q] ? s - cb * na : 0;
        row0[q] = ((long long)cb * T) * na + a;
        dh[q] = 0.f;
    }
    __syncthreads();
    for (int t = T - 1; t >= 0; t--) {
        float keep[GRU_SEQ];
#pragma unroll
        for (int q = 0; q < GRU_SEQ; q++) {
            float d_r = 0.f, d_z = 0.f, d_n = 0.f, d_nh = 0.f;
            keep[q] = 0.f;
            if (valid[q]) {
                const size_t row = (size_t)(row0[q] + (long long)t * na);
                const float* g = GATES + row * 4 * r;
                const float rg = g[j], zg = g[r + j], ng = g[2 * r + j], hn = g[3 * r + j];
                float hp;
                if (t > 0) hp = HS[(row - na) * r + j];
                else hp = h0 ? h0[(size_t)(s0 + q) * r + j] : 0.f;
                const float d = dh[q] + dHS[row * r + j];
                // h' = (hp - n) z + n
                const float dz = d * (hp - ng);
                const float dn = d * (1.f - zg);
                keep[q] = d * zg;
                d_n = dn * (1.f - ng * ng);
                d_nh = d_n * rg;
                d_r = d_n * hn * rg * (1.f - rg);
                d_z = dz * zg * (1.f - zg);
                float* o = dGI + row * 3 * r;
                o[j] = d_r; o[r + j] = d_z; o[2 * r + j] = d_n;
                float* o2 = dGH + row * 3 * r;
                o2[j] = d_r; o2[r + j] = d_z; o2[2 * r + j] = d_nh;
            }
            sg[(0 * r + j) * S_TILE + grp * GRU_SEQ + q] = d_r;
            sg[(1 * r + j) * S_TILE + grp * GRU_SEQ + q] = d_z;
            sg[(2 * r + j) * S_TILE + grp * GRU_SEQ + q] = d_nh;
        }
        __syncthreads();
        // dh_prev[k = j] = keep + sum_gj dGH[gj] * Whh[gj][k]
        float acc[GRU_SEQ];
#pragma unroll
        for (int q = 0; q < GRU_SEQ; q++) acc[q] = keep[q];
        const float* sgp = sg + grp * GRU_SEQ;
        for (int gj = 0; gj < 3 * r; gj++) {
            const float wv = w[gj * r + j];
            const float4 gv = *reinterpret_cast<const float4*>(sgp + gj * S_TILE);
            acc[0] = fmaf(gv.x, wv, acc[0]);
            acc[1] = fmaf(gv.y, wv, acc[1]);
            acc[2] = fmaf(gv.z, wv, acc[2]);
            acc[3] = fmaf(gv.w, wv, acc[3]);
        }
#pragma unroll
        for (int q = 0; q < GRU_SEQ; q++) dh[q] = acc[q];
        __syncthreads();
    }
}

static int gru_check(const char* name, int n_seq, int T, int na, int r, size_t* smem, size_t extra_rows) {
    REFIL_CHECK_ARG(n_seq > 0 && T > 0 && na > 0 && n_seq % na == 0, "%s: bad n_seq=%d T=%d na=%d", name, n_seq, T, na);
    REFIL_CHECK_ARG(r >= 8 && r <= GRU_THREADS && (r & (r - 1)) == 0, "%s: rnn_hidden_dim %d must be a power of two in [8,128]", name, r);
    const int S_TILE = (GRU_THREADS / r) * GRU_SEQ;
    *smem = ((size_t)3 * r * r + extra_rows * r * S_TILE) * sizeof(float);
    return REFIL_OK;
}

template <class K>
static int gru_set_smem(K kernel, size_t smem, const char* name) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            refil_set_error("%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
            return REFIL_ERR_CUDA;
        }
    }
    return REFIL_OK;
}

extern "C" int refil_gru_scan_fwd(const float* GI, const float* Whh, const float* bhh, const float* h0, float* HS,
                                  float* gates, int n_seq, int T, int n_agents, int r, cudaStream_t stream) {
    size_t smem = 0;
    int rc = gru_check("gru_scan_fwd", n_seq, T, n_agents, r, &smem, 2);
    if (rc) return rc;
    REFIL_CHECK_ARG(GI && Whh && bhh && HS, "gru_scan_fwd: null pointer");
    rc = gru_set_smem(gru_scan_fwd_kernel, smem, "gru_scan_fwd");
    if (rc) return rc;
    const int S_TILE = (GRU_THREADS / r) * GRU_SEQ;
    gru_scan_fwd_kernel<<<refil_cdiv(n_seq, S_TILE), GRU_THREADS, smem, stream>>>(GI, Whh, bhh, h0, HS, gates, n_seq, T,
                                                                                n_agents, r);
    REFIL_CHECK_LAUNCH("gru_scan_fwd");
    return REFIL_OK;
}

extern "C" int refil_gru_scan_bwd(const float* dHS, const float* gates, const float* HS, const float* h0,
                                  const float* Whh, float* dGI, float* dGH, int n_seq, int T, int n_agents, int r,
                                  cudaStream_t stream) {
    size_t smem = 0;
    int rc = gru_check("gru_scan_bwd", n_seq, T, n_agents, r, &smem, 3);
    if (rc) return rc;
    REFIL_CHECK_ARG(dHS && gates && HS && Whh && dGI && dGH, "gru_scan_bwd: null pointer");
    rc = gru_set_smem(gru_scan_bwd_kernel, smem, "gru_scan_bwd");
    if (rc) return rc;
    const int S_TILE = (GRU_THREADS / r) * GRU_SEQ;
    gru_scan_bwd_kernel<<<refil_cdiv(n_seq, S_TILE), GRU_THREADS, smem, stream>>>(dHS, gates, HS, h0, Whh, dGI, dGH,
                                                                                n_seq, T, n_agents, r);
    REFIL_CHECK_LAUNCH("gru_scan_bwd");
    return REFIL_OK;
}
