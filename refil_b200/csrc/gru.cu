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

// ---- specialised scans: hidden size R is a template parameter (unrolled mat-vec, immediate offsets), 8 sequences per
// thread (24 FMAs per 3 weight loads + 2 broadcast 128-bit state loads), CTAs of R threads so that the grid (n_seq / 8 CTAs)
// spreads evenly over the SMs, and the global inputs of step t+1 are requested before step t is computed.
#define GRU8 8

template <int R, int SEQ, int NG>
__global__ void __launch_bounds__(R * NG) gru_scan_fwdt_kernel(const float* __restrict__ GI, const float* __restrict__ Whh,
                                                              const float* __restrict__ bhh, const float* __restrict__ h0,
                                                              float* __restrict__ HS, float* __restrict__ GATES, int n_seq,
                                                              int T, int na) {
    static_assert(SEQ == 4 || SEQ == 8, "sequences per thread");
    constexpr int S_TILE = NG * SEQ, NT = R * NG;
    extern __shared__ __align__(16) float smem[];
    float* wt = smem;                 // [R][3R]  (k-major transpose of weight_hh)
    float* hbuf = wt + 3 * R * R;     // [2][R][S_TILE]
    const int tid = threadIdx.x, j = tid % R, grp = tid / R;
    for (int idx = tid; idx < 3 * R * R; idx += NT) {
        const int gj = idx / R, k = idx - gj * R;  // Whh[gj][k]
        wt[k * 3 * R + gj] = Whh[idx];
    }
    const int s0 = blockIdx.x * S_TILE + grp * SEQ;
    long long row0[SEQ];
    bool valid[SEQ];
    float h[SEQ];
#pragma unroll
    for (int q = 0; q < SEQ; q++) {
        const int s = s0 + q;
        valid[q] = s < n_seq;
        const int cb = valid[q] ? s / na : 0, a = valid[q] ? s - cb * na : 0;
        row0[q] = ((long long)cb * T) * na + a;
        h[q] = (valid[q] && h0) ? h0[(size_t)s * R + j] : 0.f;
        hbuf[j * S_TILE + grp * SEQ + q] = h[q];
    }
    const float br = bhh[j], bz = bhh[R + j], bn = bhh[2 * R + j];
    float gn[3][SEQ];
#pragma unroll
    for (int q = 0; q < SEQ; q++) {
        const float* g = GI + (size_t)row0[q] * 3 * R;
#pragma unroll
        for (int gg = 0; gg < 3; gg++) gn[gg][q] = valid[q] ? __ldg(g + gg * R + j) : 0.f;
    }
    __syncthreads();
    int cur = 0;
    for (int t = 0; t < T; t++) {
        float gi[3][SEQ];
#pragma unroll
        for (int q = 0; q < SEQ; q++) {
#pragma unroll
            for (int gg = 0; gg < 3; gg++) gi[gg][q] = gn[gg][q];
        }
        if (t + 1 < T) {                              // next step's input projections in flight during this step
#pragma unroll
            for (int q = 0; q < SEQ; q++) {
                const float* g = GI + (size_t)(row0[q] + (long long)(t + 1) * na) * 3 * R;
#pragma unroll
                for (int gg = 0; gg < 3; gg++) gn[gg][q] = valid[q] ? __ldg(g + gg * R + j) : 0.f;
            }
        }
        float ar[SEQ], az[SEQ], an[SEQ];
#pragma unroll
        for (int q = 0; q < SEQ; q++) { ar[q] = br; az[q] = bz; an[q] = bn; }
        const float* hb = hbuf + cur * R * S_TILE + grp * SEQ;
#pragma unroll 8
        for (int k = 0; k < R; k++) {
            const float wr = wt[k * 3 * R + j], wz = wt[k * 3 * R + R + j], wn = wt[k * 3 * R + 2 * R + j];
            float hv[SEQ];
            const float4 h0v = *reinterpret_cast<const float4*>(hb + k * S_TILE);
            hv[0] = h0v.x; hv[1] = h0v.y; hv[2] = h0v.z; hv[3] = h0v.w;
            if (SEQ == 8) {
                const float4 h1v = *reinterpret_cast<const float4*>(hb + k * S_TILE + 4);
                hv[SEQ - 4] = h1v.x; hv[SEQ - 3] = h1v.y; hv[SEQ - 2] = h1v.z; hv[SEQ - 1] = h1v.w;
            }
#pragma unroll
            for (int q = 0; q < SEQ; q++) {
                ar[q] = fmaf(hv[q], wr, ar[q]);
                az[q] = fmaf(hv[q], wz, az[q]);
                an[q] = fmaf(hv[q], wn, an[q]);
            }
        }
        float* hnext = hbuf + (cur ^ 1) * R * S_TILE;
#pragma unroll
        for (int q = 0; q < SEQ; q++) {
            const float rg = sigmoidf_(gi[0][q] + ar[q]);
            const float zg = sigmoidf_(gi[1][q] + az[q]);
            const float ng = tanhf(gi[2][q] + rg * an[q]);
            const float hn = (h[q] - ng) * zg + ng;
            h[q] = hn;
            hnext[j * S_TILE + grp * SEQ + q] = hn;
            if (valid[q]) {
                const size_t row = (size_t)(row0[q] + (long long)t * na);
                HS[row * R + j] = hn;
                if (GATES) {
                    float* g = GATES + row * 4 * R;
                    g[j] = rg; g[R + j] = zg; g[2 * R + j] = ng; g[3 * R + j] = an[q];
                }
            }
        }
        __syncthreads();
        cur ^= 1;
    }
}

template <int R>
__global__ void __launch_bounds__(R) gru_scan_bwd8_kernel(const float* __restrict__ dHS, const float* __restrict__ GATES,
                                                         const float* __restrict__ HS, const float* __restrict__ h0,
                                                         const float* __restrict__ Whh, float* __restrict__ dGI,
                                                         float* __restrict__ dGH, int n_seq, int T, int na) {
    extern __shared__ __align__(16) float smem[];
    float* w = smem;               // [3R][R]  (weight_hh as stored)
    float* sg = w + 3 * R * R;     // [3R][8]
    const int j = threadIdx.x;
    for (int idx = j; idx < 3 * R * R; idx += R) w[idx] = Whh[idx];
    const int s0 = blockIdx.x * GRU8;
    long long row0[GRU8];
    bool valid[GRU8];
    float dh[GRU8];
#pragma unroll
    for (int q = 0; q < GRU8; q++) {
        const int s = s0 + q;
        valid[q] = s < n_seq;
        const int cb = valid[q] ? s / na : 0, a = valid[q] ? s - cb * na : 0;
        row0[q] = ((long long)cb * T) * na + a;
        dh[q] = 0.f;
    }
    // inputs of one step: gates (r, z, n, W_hn h + b_hn), previous state, upstream gradient
    float nr[GRU8], nz[GRU8], nn[GRU8], nh[GRU8], np[GRU8], nd[GRU8];
    auto load_step = [&](int t) {
#pragma unroll
        for (int q = 0; q < GRU8; q++) {
            nr[q] = nz[q] = nn[q] = nh[q] = np[q] = nd[q] = 0.f;
            if (valid[q]) {
                const size_t row = (size_t)(row0[q] + (long long)t * na);
                const float* g = GATES + row * 4 * R;
                nr[q] = __ldg(g + j); nz[q] = __ldg(g + R + j); nn[q] = __ldg(g + 2 * R + j); nh[q] = __ldg(g + 3 * R + j);
                if (t > 0) np[q] = __ldg(HS + (row - na) * R + j);
                else np[q] = h0 ? __ldg(h0 + (size_t)(s0 + q) * R + j) : 0.f;
                nd[q] = __ldg(dHS + row * R + j);
            }
        }
    };
    load_step(T - 1);
    __syncthreads();
    for (int t = T - 1; t >= 0; t--) {
        float keep[GRU8];
        float d_r[GRU8], d_z[GRU8], d_nh[GRU8], d_n[GRU8];
#pragma unroll
        for (int q = 0; q < GRU8; q++) {
            const float rg = nr[q], zg = nz[q], ng = nn[q], hn = nh[q], hp = np[q];
            const float d = dh[q] + nd[q];
            // h' = (hp - n) z + n
            const float dz = d * (hp - ng);
            const float dn = d * (1.f - zg);
            keep[q] = valid[q] ? d * zg : 0.f;
            d_n[q] = dn * (1.f - ng * ng);
            d_nh[q] = d_n[q] * rg;
            d_r[q] = d_n[q] * hn * rg * (1.f - rg);
            d_z[q] = dz * zg * (1.f - zg);
            if (!valid[q]) { d_n[q] = 0.f; d_nh[q] = 0.f; d_r[q] = 0.f; d_z[q] = 0.f; }
        }
        if (t > 0) load_step(t - 1);                  // next step's inputs in flight during the mat-vec below
#pragma unroll
        for (int q = 0; q < GRU8; q++) {
            if (valid[q]) {
                const size_t row = (size_t)(row0[q] + (long long)t * na);
                float* o = dGI + row * 3 * R;
                o[j] = d_r[q]; o[R + j] = d_z[q]; o[2 * R + j] = d_n[q];
                float* o2 = dGH + row * 3 * R;
                o2[j] = d_r[q]; o2[R + j] = d_z[q]; o2[2 * R + j] = d_nh[q];
            }
            sg[(0 * R + j) * GRU8 + q] = d_r[q];
            sg[(1 * R + j) * GRU8 + q] = d_z[q];
            sg[(2 * R + j) * GRU8 + q] = d_nh[q];
        }
        __syncthreads();
        // dh_prev[k = j] = keep + sum_gj dGH[gj] * Whh[gj][k]
        float acc[GRU8];
#pragma unroll
        for (int q = 0; q < GRU8; q++) acc[q] = keep[q];
#pragma unroll 8
        for (int gj = 0; gj < 3 * R; gj++) {
            const float wv = w[gj * R + j];
            const float4 g0 = *reinterpret_cast<const float4*>(sg + gj * GRU8);
            const float4 g1 = *reinterpret_cast<const float4*>(sg + gj * GRU8 + 4);
            acc[0] = fmaf(g0.x, wv, acc[0]); acc[1] = fmaf(g0.y, wv, acc[1]);
            acc[2] = fmaf(g0.z, wv, acc[2]); acc[3] = fmaf(g0.w, wv, acc[3]);
            acc[4] = fmaf(g1.x, wv, acc[4]); acc[5] = fmaf(g1.y, wv, acc[5]);
            acc[6] = fmaf(g1.z, wv, acc[6]); acc[7] = fmaf(g1.w, wv, acc[7]);
        }
#pragma unroll
        for (int q = 0; q < GRU8; q++) dh[q] = acc[q];
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
    if (r == 64 || r == 32 || r == 128) {           // specialised scan: 128 threads, 4 sequences per thread
#define GRU_FWDT(RV, NGV)                                                                                                  \
    {                                                                                                                      \
        const size_t smt = ((size_t)3 * RV * RV + 2 * (size_t)RV * NGV * 4) * sizeof(float);                               \
        rc = gru_set_smem(gru_scan_fwdt_kernel<RV, 4, NGV>, smt, "gru_scan_fwd");                                          \
        if (rc) return rc;                                                                                                 \
        gru_scan_fwdt_kernel<RV, 4, NGV><<<refil_cdiv(n_seq, NGV * 4), RV * NGV, smt, stream>>>(GI, Whh, bhh, h0, HS, gates, \
                                                                                             n_seq, T, n_agents);          \
    }
        if (r == 64) GRU_FWDT(64, 2) else if (r == 32) GRU_FWDT(32, 4) else GRU_FWDT(128, 1)
#undef GRU_FWDT
        REFIL_CHECK_LAUNCH("gru_scan_fwd");
        return REFIL_OK;
    }
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
    if (r == 64 || r == 32 || r == 128) {
        const size_t sm8 = ((size_t)3 * r * r + 3 * (size_t)r * GRU8) * sizeof(float);
        const int grid8 = refil_cdiv(n_seq, GRU8);
#define GRU_BWD8(RV)                                                                                              \
    {                                                                                                             \
        rc = gru_set_smem(gru_scan_bwd8_kernel<RV>, sm8, "gru_scan_bwd");                                         \
        if (rc) return rc;                                                                                        \
        gru_scan_bwd8_kernel<RV><<<grid8, RV, sm8, stream>>>(dHS, gates, HS, h0, Whh, dGI, dGH, n_seq, T, n_agents); \
    }
        if (r == 64) GRU_BWD8(64) else if (r == 32) GRU_BWD8(32) else GRU_BWD8(128)
#undef GRU_BWD8
        REFIL_CHECK_LAUNCH("gru_scan_bwd");
        return REFIL_OK;
    }
    rc = gru_set_smem(gru_scan_bwd_kernel, smem, "gru_scan_bwd");
    if (rc) return rc;
    const int S_TILE = (GRU_THREADS / r) * GRU_SEQ;
    gru_scan_bwd_kernel<<<refil_cdiv(n_seq, S_TILE), GRU_THREADS, smem, stream>>>(dHS, gates, HS, h0, Whh, dGI, dGH,
                                                                                n_seq, T, n_agents, r);
    REFIL_CHECK_LAUNCH("gru_scan_bwd");
    return REFIL_OK;
}
