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
#include <stdlib.h>

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

// REFIL_GRU_MODE: "v3" (default) register-resident FFMA scans, "mma" the mma.sync 3xTF32 scans, "ffma" the first-generation scans
static int gru_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("REFIL_GRU_MODE");
        mode = (e && e[0] == 'f') ? 0 : (e && e[0] == 'm') ? 1 : 2;
    }
    return mode;
}
static bool gru_use_mma() { return gru_mode() == 1; }

static int gru_set_smem_attr(const void* kernel, size_t smem, const char* name) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            refil_set_error("%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
            return REFIL_ERR_CUDA;
        }
    }
    return REFIL_OK;
}

// =====================================================================================================================
// Tensor-core scans (hidden size 32 / 64): the per-step recurrent product is a small GEMM -- [16 sequences x R] x [R x 3R]
// forward, [16 x 3R] x [3R x R] backward -- issued as warp-level mma.sync m16n8k8 TF32 with the same 3xTF32 hi/lo split as the
// dense layers (fp32-grade: ~2^-21 relative per product).  The FFMA scans above spend ~780 warp instructions per (sequence,
// step) and ~3 us per step however few sequences a CTA holds (ncu r2d: one warp per scheduler, stalled on shared-memory
// latency); here
//   * W_hh lives in REGISTERS as pre-split B fragments (warp w owns features [8w, 8w+8) of all three gates: 96 registers at
//     R = 64), loaded once per launch;
//   * the state h (forward) / the gate gradients (backward) of the CTA's 16 x MT sequences sit in a double-buffered
//     shared-memory tile, padded so that the A-fragment reads of a warp hit 32 distinct banks; ONE __syncthreads per step;
//   * the accumulator fragment of a thread -- rows gid, gid + 8, features 8w + 2 tig, + 1 -- is exactly the set of
//     (sequence, feature) pairs whose gate math it does, so the MMA result never leaves registers;
//   * the global inputs of step t + 1 are requested before step t is computed.
// 72 x MT mma per warp and step; a 16-episode shard (384 sequences) runs on 24 CTAs in ~0.5 us per step.
// =====================================================================================================================
__device__ __forceinline__ void gru_mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                             uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void gru_split(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
// acc += A (4 fp32 fragment values) x B (pre-split fragment), three TF32 passes
__device__ __forceinline__ void gru_mma3(float (&c)[4], const float (&a)[4], const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
    uint32_t ah[4], al[4];
#pragma unroll
    for (int i = 0; i < 4; i++) gru_split(a[i], ah[i], al[i]);
    gru_mma_tf32(c, al[0], al[1], al[2], al[3], bh[0], bh[1]);
    gru_mma_tf32(c, ah[0], ah[1], ah[2], ah[3], bl[0], bl[1]);
    gru_mma_tf32(c, ah[0], ah[1], ah[2], ah[3], bh[0], bh[1]);
}

template <int R, int MT>
__global__ void __launch_bounds__(R * 4, 1) gru_scan_fwd_mma_kernel(const float* __restrict__ GI, const float* __restrict__ Whh,
                                                                    const float* __restrict__ bhh, const float* __restrict__ h0,
                                                                    float* __restrict__ HS, float* __restrict__ GATES, int n_seq,
                                                                    int T, int na) {
    constexpr int KT = R / 8, LD = R + 4, NT = R * 4;      // R / 8 warps, one per 8 features
    __shared__ float hsm[2][MT * 16][LD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
    const int f0 = 8 * warp + 2 * tig;                     // my two features f0, f0 + 1 (accumulator columns 2 tig, 2 tig + 1)
    // B fragments: B[k][n] = Whh[g * R + 8 warp + n][k]; b0: k = 8 kt + tig, b1: k + 4; n = gid
    uint32_t bh[KT][3][2], bl[KT][3][2];
#pragma unroll
    for (int kt = 0; kt < KT; kt++)
#pragma unroll
        for (int g = 0; g < 3; g++)
#pragma unroll
            for (int h = 0; h < 2; h++)
                gru_split(__ldg(Whh + (size_t)(g * R + 8 * warp + gid) * R + 8 * kt + tig + 4 * h), bh[kt][g][h], bl[kt][g][h]);
    float bias[3][2];
#pragma unroll
    for (int g = 0; g < 3; g++) { bias[g][0] = __ldg(bhh + g * R + f0); bias[g][1] = __ldg(bhh + g * R + f0 + 1); }
    // my sequences: rows gid and gid + 8 of every 16-row m-tile
    long long row0[MT][2];
    bool valid[MT][2];
    float hreg[MT][2][2];
    const int sbase = blockIdx.x * (MT * 16);
#pragma unroll
    for (int m = 0; m < MT; m++)
#pragma unroll
        for (int hr = 0; hr < 2; hr++) {
            const int s = sbase + m * 16 + gid + 8 * hr;
            valid[m][hr] = s < n_seq;
            const int cb = valid[m][hr] ? s / na : 0, a = valid[m][hr] ? s - cb * na : 0;
            row0[m][hr] = ((long long)cb * T) * na + a;
            float2 hv = make_float2(0.f, 0.f);
            if (valid[m][hr] && h0) hv = __ldg(reinterpret_cast<const float2*>(h0 + (size_t)s * R + f0));
            hreg[m][hr][0] = hv.x; hreg[m][hr][1] = hv.y;
            hsm[0][m * 16 + gid + 8 * hr][f0] = hv.x;
            hsm[0][m * 16 + gid + 8 * hr][f0 + 1] = hv.y;
        }
    float2 gn[MT][2][3];                                   // input projections of the next step
    auto load_gi = [&](int t) {
#pragma unroll
        for (int m = 0; m < MT; m++)
#pragma unroll
            for (int hr = 0; hr < 2; hr++) {
                const float* g = GI + (size_t)(row0[m][hr] + (long long)t * na) * 3 * R + f0;
#pragma unroll
                for (int gg = 0; gg < 3; gg++)
                    gn[m][hr][gg] = valid[m][hr] ? __ldg(reinterpret_cast<const float2*>(g + gg * R)) : make_float2(0.f, 0.f);
            }
    };
    load_gi(0);
    __syncthreads();
    int cur = 0;
    for (int t = 0; t < T; t++) {
        float2 gi[MT][2][3];
#pragma unroll
        for (int m = 0; m < MT; m++)
#pragma unroll
            for (int hr = 0; hr < 2; hr++)
#pragma unroll
                for (int gg = 0; gg < 3; gg++) gi[m][hr][gg] = gn[m][hr][gg];
        if (t + 1 < T) load_gi(t + 1);
        float acc[MT][3][4];
#pragma unroll
        for (int m = 0; m < MT; m++)
#pragma unroll
            for (int g = 0; g < 3; g++) { acc[m][g][0] = bias[g][0]; acc[m][g][1] = bias[g][1]; acc[m][g][2] = bias[g][0]; acc[m][g][3] = bias[g][1]; }
#pragma unroll
        for (int kt = 0; kt < KT; kt++) {
#pragma unroll
            for (int m = 0; m < MT; m++) {
                const float* hb = &hsm[cur][m * 16][0];
                float a[4];
                a[0] = hb[gid * LD + 8 * kt + tig];
                a[1] = hb[(gid + 8) * LD + 8 * kt + tig];
                a[2] = hb[gid * LD + 8 * kt + tig + 4];
                a[3] = hb[(gid + 8) * LD + 8 * kt + tig + 4];
#pragma unroll
                for (int g = 0; g < 3; g++) gru_mma3(acc[m][g], a, bh[kt][g], bl[kt][g]);
            }
        }
#pragma unroll
        for (int m = 0; m < MT; m++)
#pragma unroll
            for (int hr = 0; hr < 2; hr++) {
                float hn[2], rgv[2], zgv[2], ngv[2], anv[2];
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const float ar = acc[m][0][2 * hr + e], az = acc[m][1][2 * hr + e], an = acc[m][2][2 * hr + e];
                    const float gir = e ? gi[m][hr][0].y : gi[m][hr][0].x, giz = e ? gi[m][hr][1].y : gi[m][hr][1].x,
                                gin = e ? gi[m][hr][2].y : gi[m][hr][2].x;
                    const float rg = sigmoidf_(gir + ar), zg = sigmoidf_(giz + az);
                    const float ng = tanhf(gin + rg * an);
                    hn[e] = (hreg[m][hr][e] - ng) * zg + ng;
                    hreg[m][hr][e] = hn[e];
                    rgv[e] = rg; zgv[e] = zg; ngv[e] = ng; anv[e] = an;
                }
                float* hw = &hsm[cur ^ 1][m * 16 + gid + 8 * hr][f0];
                hw[0] = hn[0]; hw[1] = hn[1];
                if (valid[m][hr]) {
                    const size_t row = (size_t)(row0[m][hr] + (long long)t * na);
                    *reinterpret_cast<float2*>(HS + row * R + f0) = make_float2(hn[0], hn[1]);
                    if (GATES) {
                        float* g = GATES + row * 4 * R + f0;
                        *reinterpret_cast<float2*>(g) = make_float2(rgv[0], rgv[1]);
                        *reinterpret_cast<float2*>(g + R) = make_float2(zgv[0], zgv[1]);
                        *reinterpret_cast<float2*>(g + 2 * R) = make_float2(ngv[0], ngv[1]);
                        *reinterpret_cast<float2*>(g + 3 * R) = make_float2(anv[0], anv[1]);
                    }
                }
            }
        __syncthreads();
        cur ^= 1;
    }
    (void)NT;
}

template <int R, int MT>
__global__ void __launch_bounds__(R * 4, 1) gru_scan_bwd_mma_kernel(const float* __restrict__ dHS, const float* __restrict__ GATES,
                                                                    const float* __restrict__ HS, const float* __restrict__ h0,
                                                                    const float* __restrict__ Whh, float* __restrict__ dGI,
                                                                    float* __restrict__ dGH, int n_seq, int T, int na) {
    constexpr int KT = 3 * R / 8, LD = 3 * R + 4;          // reduction over the 3R gate rows of W_hh
    extern __shared__ __align__(16) float smem_g[];        // [2][MT * 16][LD]: (d_r | d_z | d_nh) of the CTA's sequences
    float (*sg)[LD] = reinterpret_cast<float (*)[LD]>(smem_g);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
    const int f0 = 8 * warp + 2 * tig;
    // B fragments: B[k = gj][n] = Whh[gj][8 warp + n]; b0: gj = 8 kt + tig, b1: gj + 4; n = gid
    uint32_t bh[KT][2], bl[KT][2];
#pragma unroll
    for (int kt = 0; kt < KT; kt++)
#pragma unroll
        for (int h = 0; h < 2; h++)
            gru_split(__ldg(Whh + (size_t)(8 * kt + tig + 4 * h) * R + 8 * warp + gid), bh[kt][h], bl[kt][h]);
    long long row0[MT][2];
    bool valid[MT][2];
    int sidx[MT][2];
    float dh[MT][2][2];
    const int sbase = blockIdx.x * (MT * 16);
#pragma unroll
    for (int m = 0; m < MT; m++)
#pragma unroll
        for (int hr = 0; hr < 2; hr++) {
            const int s = sbase + m * 16 + gid + 8 * hr;
            valid[m][hr] = s < n_seq;
            sidx[m][hr] = s;
            const int cb = valid[m][hr] ? s / na : 0, a = valid[m][hr] ? s - cb * na : 0;
            row0[m][hr] = ((long long)cb * T) * na + a;
            dh[m][hr][0] = 0.f; dh[m][hr][1] = 0.f;
        }
    // inputs of one step: gates (r, z, n, W_hn h + b_hn), previous state, upstream gradient
    float2 nx[MT][2][6];
    auto load_step = [&](int t) {
#pragma unroll
        for (int m = 0; m < MT; m++)
#pragma unroll
            for (int hr = 0; hr < 2; hr++) {
#pragma unroll
                for (int q = 0; q < 6; q++) nx[m][hr][q] = make_float2(0.f, 0.f);
                if (valid[m][hr]) {
                    const size_t row = (size_t)(row0[m][hr] + (long long)t * na);
                    const float* g = GATES + row * 4 * R + f0;
#pragma unroll
                    for (int q = 0; q < 4; q++) nx[m][hr][q] = __ldg(reinterpret_cast<const float2*>(g + q * R));
                    if (t > 0) nx[m][hr][4] = __ldg(reinterpret_cast<const float2*>(HS + (row - na) * R + f0));
                    else if (h0) nx[m][hr][4] = __ldg(reinterpret_cast<const float2*>(h0 + (size_t)sidx[m][hr] * R + f0));
                    nx[m][hr][5] = __ldg(reinterpret_cast<const float2*>(dHS + row * R + f0));
                }
            }
    };
    load_step(T - 1);
    int cur = 0;
    for (int t = T - 1; t >= 0; t--) {
        float keep[MT][2][2];
        float (*sc)[LD] = sg + cur * (MT * 16);
#pragma unroll
        for (int m = 0; m < MT; m++)
#pragma unroll
            for (int hr = 0; hr < 2; hr++) {
                float dr[2], dz_[2], dn[2], dnh[2];
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const float rg = e ? nx[m][hr][0].y : nx[m][hr][0].x, zg = e ? nx[m][hr][1].y : nx[m][hr][1].x,
                                ng = e ? nx[m][hr][2].y : nx[m][hr][2].x, hn = e ? nx[m][hr][3].y : nx[m][hr][3].x,
                                hp = e ? nx[m][hr][4].y : nx[m][hr][4].x, du = e ? nx[m][hr][5].y : nx[m][hr][5].x;
                    const float d = dh[m][hr][e] + du;
                    // h' = (hp - n) z + n
                    const float dz = d * (hp - ng);
                    const float dnn = d * (1.f - zg);
                    keep[m][hr][e] = valid[m][hr] ? d * zg : 0.f;
                    dn[e] = valid[m][hr] ? dnn * (1.f - ng * ng) : 0.f;
                    dnh[e] = dn[e] * rg;
                    dr[e] = dn[e] * hn * rg * (1.f - rg);
                    dz_[e] = valid[m][hr] ? dz * zg * (1.f - zg) : 0.f;
                }
                float* srow = &sc[m * 16 + gid + 8 * hr][f0];
                srow[0] = dr[0]; srow[1] = dr[1];
                srow[R] = dz_[0]; srow[R + 1] = dz_[1];
                srow[2 * R] = dnh[0]; srow[2 * R + 1] = dnh[1];
                if (valid[m][hr]) {
                    const size_t row = (size_t)(row0[m][hr] + (long long)t * na);
                    float* o = dGI + row * 3 * R + f0;
                    *reinterpret_cast<float2*>(o) = make_float2(dr[0], dr[1]);
                    *reinterpret_cast<float2*>(o + R) = make_float2(dz_[0], dz_[1]);
                    *reinterpret_cast<float2*>(o + 2 * R) = make_float2(dn[0], dn[1]);
                    float* o2 = dGH + row * 3 * R + f0;
                    *reinterpret_cast<float2*>(o2) = make_float2(dr[0], dr[1]);
                    *reinterpret_cast<float2*>(o2 + R) = make_float2(dz_[0], dz_[1]);
                    *reinterpret_cast<float2*>(o2 + 2 * R) = make_float2(dnh[0], dnh[1]);
                }
            }
        if (t > 0) load_step(t - 1);                  // next step's inputs in flight during the product below
        __syncthreads();
        // dh_prev[seq][k] = keep + sum_gj (d_r | d_z | d_nh)[seq][gj] * Whh[gj][k]; three independent accumulator chains
        float acc[MT][3][4];
#pragma unroll
        for (int m = 0; m < MT; m++) {
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int e = 0; e < 4; e++) acc[m][c][e] = 0.f;
            acc[m][0][0] = keep[m][0][0]; acc[m][0][1] = keep[m][0][1]; acc[m][0][2] = keep[m][1][0]; acc[m][0][3] = keep[m][1][1];
        }
#pragma unroll
        for (int kt = 0; kt < KT; kt++) {
#pragma unroll
            for (int m = 0; m < MT; m++) {
                const float* ab = &sc[m * 16][0];
                float a[4];
                a[0] = ab[gid * LD + 8 * kt + tig];
                a[1] = ab[(gid + 8) * LD + 8 * kt + tig];
                a[2] = ab[gid * LD + 8 * kt + tig + 4];
                a[3] = ab[(gid + 8) * LD + 8 * kt + tig + 4];
                gru_mma3(acc[m][kt % 3], a, bh[kt], bl[kt]);
            }
        }
#pragma unroll
        for (int m = 0; m < MT; m++) {
            dh[m][0][0] = acc[m][0][0] + acc[m][1][0] + acc[m][2][0];
            dh[m][0][1] = acc[m][0][1] + acc[m][1][1] + acc[m][2][1];
            dh[m][1][0] = acc[m][0][2] + acc[m][1][2] + acc[m][2][2];
            dh[m][1][1] = acc[m][0][3] + acc[m][1][3] + acc[m][2][3];
        }
        cur ^= 1;                                     // the other buffer is written next step: one barrier per step
    }
}

// =====================================================================================================================
// Register-resident FFMA scans (default for hidden size 32 / 64).  Measured on B200 (profiles/r2f): the legacy mma.sync
// TF32 path retires one m16n8k8 per ~3.2 cycles per SM, i.e. ~107 MAC/clk/SM after the three passes of the 3xTF32 split --
// below the 128 FMA/clk/SM of the fp32 pipe -- so the recurrence gains nothing from it; what it needs is a SHORT step on as
// many SMs as there are sequences to spread.  Here
//   * W_hh never leaves the registers: lane = (feature j of a 16-feature group, half kh of the reduction), 96 weights per
//     thread at R = 64, loaded once per launch -- the inner loop is 12 FMAs per 128-bit broadcast load of the state;
//   * the two halves of a reduction meet with ONE shuffle per (gate, sequence), no shared-memory partials;
//   * the state (forward) / gate gradients (backward) of the CTA's sequences are double-buffered in shared memory: one
//     __syncthreads per step; the global inputs of step t + 1 are requested before step t is computed;
//   * a CTA takes S = 2 x SB x NB sequences, chosen so that the grid is a single wave: 4 sequences per CTA for a 16-episode shard
//     (384 sequences on 96 SMs, ~0.5 us per step), 24 for the full north-star batch (3072 sequences on 128 SMs).
// =====================================================================================================================
__device__ __forceinline__ float gru_sel(bool hi, float a, float b) { return hi ? b : a; }

template <int R, int SB, int NB>
__global__ void __launch_bounds__(R * 4, 1) gru_scan_fwd_v3_kernel(const float* __restrict__ GI, const float* __restrict__ Whh,
                                                                   const float* __restrict__ bhh, const float* __restrict__ h0,
                                                                   float* __restrict__ HS, float* __restrict__ GATES, int n_seq,
                                                                   int T, int na) {
    constexpr int KH = R / 2, FG = R / 16, S = 2 * SB * NB, PP = SB / 2;
    __shared__ __align__(16) float hbuf[2][S][R];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, jl = lane & 15, kh = lane >> 4;
    const int fg = warp % FG, sgp = warp / FG, j = 16 * fg + jl;
    float w[3][KH];
#pragma unroll
    for (int g = 0; g < 3; g++)
#pragma unroll
        for (int kk = 0; kk < KH; kk += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(Whh + (size_t)(g * R + j) * R + KH * kh + kk));
            w[g][kk] = v.x; w[g][kk + 1] = v.y; w[g][kk + 2] = v.z; w[g][kk + 3] = v.w;
        }
    const float br = __ldg(bhh + j), bz = __ldg(bhh + R + j), bn = __ldg(bhh + 2 * R + j);
    pdl_launch_dependents();
    pdl_wait();                        // the 96 weight registers above were loaded beside the previous kernel's tail
    // my (sequence, feature j) pairs: block b = sgp * NB + nb holds local sequences [b SB, b SB + SB); mine are kh + 2 p
    long long row0[NB][PP];
    bool valid[NB][PP];
    float hreg[NB][PP];
    const int sbase = blockIdx.x * S;
#pragma unroll
    for (int nb = 0; nb < NB; nb++)
#pragma unroll
        for (int p = 0; p < PP; p++) {
            const int ls = (sgp * NB + nb) * SB + kh + 2 * p, sq = sbase + ls;
            valid[nb][p] = sq < n_seq;
            const int cb = valid[nb][p] ? sq / na : 0, a = valid[nb][p] ? sq - cb * na : 0;
            row0[nb][p] = ((long long)cb * T) * na + a;
            hreg[nb][p] = (valid[nb][p] && h0) ? __ldg(h0 + (size_t)sq * R + j) : 0.f;
            hbuf[0][ls][j] = hreg[nb][p];
        }
    float gn[NB][PP][3];
    auto load_gi = [&](int t) {
#pragma unroll
        for (int nb = 0; nb < NB; nb++)
#pragma unroll
            for (int p = 0; p < PP; p++) {
                const float* g = GI + (size_t)(row0[nb][p] + (long long)t * na) * 3 * R + j;
#pragma unroll
                for (int gg = 0; gg < 3; gg++) gn[nb][p][gg] = valid[nb][p] ? __ldg(g + gg * R) : 0.f;
            }
    };
    load_gi(0);
    __syncthreads();
    int cur = 0;
    for (int t = 0; t < T; t++) {
        float gi[NB][PP][3];
#pragma unroll
        for (int nb = 0; nb < NB; nb++)
#pragma unroll
            for (int p = 0; p < PP; p++)
#pragma unroll
                for (int gg = 0; gg < 3; gg++) gi[nb][p][gg] = gn[nb][p][gg];
        if (t + 1 < T) load_gi(t + 1);
#pragma unroll
        for (int nb = 0; nb < NB; nb++) {
            float acc[3][SB];
#pragma unroll
            for (int g = 0; g < 3; g++)
#pragma unroll
                for (int q = 0; q < SB; q++) acc[g][q] = 0.f;
            const float* hb = &hbuf[cur][(sgp * NB + nb) * SB][KH * kh];
#pragma unroll
            for (int kk = 0; kk < KH; kk += 4) {
#pragma unroll
                for (int q = 0; q < SB; q++) {
                    const float4 hv = *reinterpret_cast<const float4*>(hb + q * R + kk);
#pragma unroll
                    for (int g = 0; g < 3; g++) {
                        acc[g][q] = fmaf(hv.x, w[g][kk], acc[g][q]);
                        acc[g][q] = fmaf(hv.y, w[g][kk + 1], acc[g][q]);
                        acc[g][q] = fmaf(hv.z, w[g][kk + 2], acc[g][q]);
                        acc[g][q] = fmaf(hv.w, w[g][kk + 3], acc[g][q]);
                    }
                }
            }
#pragma unroll
            for (int p = 0; p < PP; p++) {
                // my pair is sequence kh + 2p of the block: I send my partner the partial sums of ITS sequence and add the
                // partner's partial sums of mine
                float sum[3];
#pragma unroll
                for (int g = 0; g < 3; g++) {
                    const float send = gru_sel(kh != 0, acc[g][2 * p + 1], acc[g][2 * p]);
                    const float mine = gru_sel(kh != 0, acc[g][2 * p], acc[g][2 * p + 1]);
                    sum[g] = mine + __shfl_xor_sync(0xffffffffu, send, 16);
                }
                const float an = sum[2] + bn;
                const float rg = sigmoidf_(gi[nb][p][0] + sum[0] + br);
                const float zg = sigmoidf_(gi[nb][p][1] + sum[1] + bz);
                const float ng = tanhf(gi[nb][p][2] + rg * an);
                const float hn = (hreg[nb][p] - ng) * zg + ng;
                hreg[nb][p] = hn;
                hbuf[cur ^ 1][(sgp * NB + nb) * SB + kh + 2 * p][j] = hn;
                if (valid[nb][p]) {
                    const size_t row = (size_t)(row0[nb][p] + (long long)t * na);
                    HS[row * R + j] = hn;
                    if (GATES) {
                        float* g = GATES + row * 4 * R;
                        g[j] = rg; g[R + j] = zg; g[2 * R + j] = ng; g[3 * R + j] = an;
                    }
                }
            }
        }
        __syncthreads();
        cur ^= 1;
    }
}

template <int R, int SB, int NB>
__global__ void __launch_bounds__(R * 4, 1) gru_scan_bwd_v3_kernel(const float* __restrict__ dHS, const float* __restrict__ GATES,
                                                                   const float* __restrict__ HS, const float* __restrict__ h0,
                                                                   const float* __restrict__ Whh, float* __restrict__ dGI,
                                                                   float* __restrict__ dGH, int n_seq, int T, int na) {
    constexpr int GH = 3 * R / 2, FG = R / 16, S = 2 * SB * NB, PP = SB / 2;
    __shared__ __align__(16) float abuf[2][S][3 * R];          // (d_r | d_z | d_nh) of the CTA's sequences
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, jl = lane & 15, kh = lane >> 4;
    const int fg = warp % FG, sgp = warp / FG, j = 16 * fg + jl;
    float w[GH];                                               // Whh[gj][j] for my half of the 3R gate rows
#pragma unroll
    for (int i = 0; i < GH; i++) w[i] = __ldg(Whh + (size_t)(GH * kh + i) * R + j);
    pdl_launch_dependents();
    pdl_wait();
    long long row0[NB][PP];
    bool valid[NB][PP];
    int sidx[NB][PP];
    float dh[NB][PP];
    const int sbase = blockIdx.x * S;
#pragma unroll
    for (int nb = 0; nb < NB; nb++)
#pragma unroll
        for (int p = 0; p < PP; p++) {
            const int ls = (sgp * NB + nb) * SB + kh + 2 * p, sq = sbase + ls;
            valid[nb][p] = sq < n_seq;
            sidx[nb][p] = sq;
            const int cb = valid[nb][p] ? sq / na : 0, a = valid[nb][p] ? sq - cb * na : 0;
            row0[nb][p] = ((long long)cb * T) * na + a;
            dh[nb][p] = 0.f;
        }
    float nx[NB][PP][6];                                       // r, z, n, W_hn h + b_hn, h_{t-1}, dHS of the next step
    auto load_step = [&](int t) {
#pragma unroll
        for (int nb = 0; nb < NB; nb++)
#pragma unroll
            for (int p = 0; p < PP; p++) {
#pragma unroll
                for (int q = 0; q < 6; q++) nx[nb][p][q] = 0.f;
                if (valid[nb][p]) {
                    const size_t row = (size_t)(row0[nb][p] + (long long)t * na);
                    const float* g = GATES + row * 4 * R + j;
#pragma unroll
                    for (int q = 0; q < 4; q++) nx[nb][p][q] = __ldg(g + q * R);
                    if (t > 0) nx[nb][p][4] = __ldg(HS + (row - na) * R + j);
                    else if (h0) nx[nb][p][4] = __ldg(h0 + (size_t)sidx[nb][p] * R + j);
                    nx[nb][p][5] = __ldg(dHS + row * R + j);
                }
            }
    };
    load_step(T - 1);
    int cur = 0;
    for (int t = T - 1; t >= 0; t--) {
        float keep[NB][PP];
#pragma unroll
        for (int nb = 0; nb < NB; nb++)
#pragma unroll
            for (int p = 0; p < PP; p++) {
                const float rg = nx[nb][p][0], zg = nx[nb][p][1], ng = nx[nb][p][2], hn = nx[nb][p][3], hp = nx[nb][p][4];
                const float d = dh[nb][p] + nx[nb][p][5];
                // h' = (hp - n) z + n
                const float dz = d * (hp - ng);
                const float dnn = d * (1.f - zg);
                const bool ok = valid[nb][p];
                keep[nb][p] = ok ? d * zg : 0.f;
                const float d_n = ok ? dnn * (1.f - ng * ng) : 0.f;
                const float d_nh = d_n * rg;
                const float d_r = d_n * hn * rg * (1.f - rg);
                const float d_z = ok ? dz * zg * (1.f - zg) : 0.f;
                float* ar = &abuf[cur][(sgp * NB + nb) * SB + kh + 2 * p][j];
                ar[0] = d_r; ar[R] = d_z; ar[2 * R] = d_nh;
                if (ok) {
                    const size_t row = (size_t)(row0[nb][p] + (long long)t * na);
                    float* o = dGI + row * 3 * R;
                    o[j] = d_r; o[R + j] = d_z; o[2 * R + j] = d_n;
                    float* o2 = dGH + row * 3 * R;
                    o2[j] = d_r; o2[R + j] = d_z; o2[2 * R + j] = d_nh;
                }
            }
        if (t > 0) load_step(t - 1);                  // next step's inputs in flight during the product below
        __syncthreads();
        // dh_prev[seq][j] = keep + sum_gj (d_r | d_z | d_nh)[seq][gj] * Whh[gj][j]
#pragma unroll
        for (int nb = 0; nb < NB; nb++) {
            float acc[SB];
#pragma unroll
            for (int q = 0; q < SB; q++) acc[q] = 0.f;
            const float* ab = &abuf[cur][(sgp * NB + nb) * SB][GH * kh];
#pragma unroll
            for (int i = 0; i < GH; i += 4) {
#pragma unroll
                for (int q = 0; q < SB; q++) {
                    const float4 av = *reinterpret_cast<const float4*>(ab + q * 3 * R + i);
                    acc[q] = fmaf(av.x, w[i], acc[q]);
                    acc[q] = fmaf(av.y, w[i + 1], acc[q]);
                    acc[q] = fmaf(av.z, w[i + 2], acc[q]);
                    acc[q] = fmaf(av.w, w[i + 3], acc[q]);
                }
            }
#pragma unroll
            for (int p = 0; p < PP; p++) {
                const float send = gru_sel(kh != 0, acc[2 * p + 1], acc[2 * p]);
                const float mine = gru_sel(kh != 0, acc[2 * p], acc[2 * p + 1]);
                dh[nb][p] = keep[nb][p] + mine + __shfl_xor_sync(0xffffffffu, send, 16);
            }
        }
        cur ^= 1;                                     // the other buffer is written next step: one barrier per step
    }
}

// sequences per CTA S = 2 SB NB in {4, 8, 16, 24}: the smallest that keeps the grid within one wave of CTAs
template <int R>
static void gru_launch_v3(bool fwd, const float* a0, const float* a1, const float* a2, const float* a3, const float* a4,
                          float* o0, float* o1, int n_seq, int T, int na, cudaStream_t stream) {
    const int sms = refil_num_sms();
    int S = 24;
    if (refil_cdiv(n_seq, 4) <= sms) S = 4;
    else if (refil_cdiv(n_seq, 8) <= sms) S = 8;
    else if (refil_cdiv(n_seq, 16) <= sms) S = 16;
    const int grid = refil_cdiv(n_seq, S);
#define GRU_V3(SBV, NBV)                                                                                                 \
    {                                                                                                                    \
        if (fwd) refil_launch(gru_scan_fwd_v3_kernel<R, SBV, NBV>, dim3(grid), dim3(R * 4), 0, stream, true, a0, a1, a2, a3, o0, o1,  \
                              n_seq, T, na);                                                                              \
        else refil_launch(gru_scan_bwd_v3_kernel<R, SBV, NBV>, dim3(grid), dim3(R * 4), 0, stream, true, a0, a1, a2, a3, a4, o0, o1, \
                          n_seq, T, na);                                                                                  \
    }
    if (S == 4) GRU_V3(2, 1) else if (S == 8) GRU_V3(4, 1) else if (S == 16) GRU_V3(4, 2) else GRU_V3(4, 3)
#undef GRU_V3
}

// MT = m16 tiles (16 sequences each) per CTA: one while the sequences fit one wave of CTAs, else two
template <int R>
static int gru_launch_mma(bool fwd, const float* a0, const float* a1, const float* a2, const float* a3, const float* a4,
                          float* o0, float* o1, int n_seq, int T, int na, cudaStream_t stream) {
    const int mt = refil_cdiv(n_seq, 16) > refil_num_sms() ? 2 : 1;
    const int grid = refil_cdiv(n_seq, 16 * mt);
    if (fwd) {
        if (mt == 1) gru_scan_fwd_mma_kernel<R, 1><<<grid, R * 4, 0, stream>>>(a0, a1, a2, a3, o0, o1, n_seq, T, na);
        else gru_scan_fwd_mma_kernel<R, 2><<<grid, R * 4, 0, stream>>>(a0, a1, a2, a3, o0, o1, n_seq, T, na);
        return REFIL_OK;
    }
    const size_t smem = (size_t)2 * mt * 16 * (3 * R + 4) * sizeof(float);
    if (mt == 1) {
        int rc = gru_set_smem_attr((const void*)gru_scan_bwd_mma_kernel<R, 1>, smem, "gru_scan_bwd");
        if (rc) return rc;
        gru_scan_bwd_mma_kernel<R, 1><<<grid, R * 4, smem, stream>>>(a0, a1, a2, a3, a4, o0, o1, n_seq, T, na);
    } else {
        int rc = gru_set_smem_attr((const void*)gru_scan_bwd_mma_kernel<R, 2>, smem, "gru_scan_bwd");
        if (rc) return rc;
        gru_scan_bwd_mma_kernel<R, 2><<<grid, R * 4, smem, stream>>>(a0, a1, a2, a3, a4, o0, o1, n_seq, T, na);
    }
    return REFIL_OK;
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
    if ((r == 64 || r == 32) && gru_mode() == 2) {
        if (r == 64) gru_launch_v3<64>(true, GI, Whh, bhh, h0, nullptr, HS, gates, n_seq, T, n_agents, stream);
        else gru_launch_v3<32>(true, GI, Whh, bhh, h0, nullptr, HS, gates, n_seq, T, n_agents, stream);
        REFIL_CHECK_LAUNCH("gru_scan_fwd (v3)");
        return REFIL_OK;
    }
    if ((r == 64 || r == 32) && gru_use_mma()) {    // tensor-core scan (REFIL_GRU_MODE=ffma selects the FFMA scans below)
        rc = r == 64 ? gru_launch_mma<64>(true, GI, Whh, bhh, h0, nullptr, HS, gates, n_seq, T, n_agents, stream)
                     : gru_launch_mma<32>(true, GI, Whh, bhh, h0, nullptr, HS, gates, n_seq, T, n_agents, stream);
        if (rc) return rc;
        REFIL_CHECK_LAUNCH("gru_scan_fwd (mma)");
        return REFIL_OK;
    }
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
    if ((r == 64 || r == 32) && gru_mode() == 2) {
        if (r == 64) gru_launch_v3<64>(false, dHS, gates, HS, h0, Whh, dGI, dGH, n_seq, T, n_agents, stream);
        else gru_launch_v3<32>(false, dHS, gates, HS, h0, Whh, dGI, dGH, n_seq, T, n_agents, stream);
        REFIL_CHECK_LAUNCH("gru_scan_bwd (v3)");
        return REFIL_OK;
    }
    if ((r == 64 || r == 32) && gru_use_mma()) {
        rc = r == 64 ? gru_launch_mma<64>(false, dHS, gates, HS, h0, Whh, dGI, dGH, n_seq, T, n_agents, stream)
                     : gru_launch_mma<32>(false, dHS, gates, HS, h0, Whh, dGI, dGH, n_seq, T, n_agents, stream);
        if (rc) return rc;
        REFIL_CHECK_LAUNCH("gru_scan_bwd (mma)");
        return REFIL_OK;
    }
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
