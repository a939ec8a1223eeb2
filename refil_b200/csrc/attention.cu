// Randomized-partition masked multi-head attention over entities (kernel K2).
//
// Replaces the bmm / masked_fill / softmax / NaN->0 / bmm sequence of
//   /root/reference/src/modules/layers/attention.py:43-64
// and the mask algebra of
//   /root/reference/src/modules/agents/entity_rnn_agent.py:79-124  (imagine: within / interact copies)
//   /root/reference/src/modules/mixers/flex_qmix.py:43-46          (hypernet default mask)
// for up to 3 mask "copies" that share one QKV tensor (the reference repeats the entities 3x and
// recomputes fc1 + in_trans for each copy; here QKV is read once per (b, t) and every copy is resolved
// against it in shared memory).
//
// Layout in HBM:
//   QKV  f32 [N, ne, 3d]   (N = B*T rows of (b, t); columns Q | K | V, heads = contiguous hd slices)
//   OUT  f32 [C, N, nq, d] (copy-major; nq = n_agents)
// Masks (1 = masked), per copy c:   masked(n,i,j) = explicit_c[n, i, j]            (optional u8 tensor)
//                                               | part(mode_c; group bits of episode b, inactive0)
//                                               | ((mode_c & 4) ? inactive0_i | inactive0_j : 0)
//                                               | ((mode_c & 8) ? em[n, i] | em[n, j]       : 0)
//   part W (mode&3 == 1): masked unless i, j are in the same random group and both present at t = 0
//   part I (mode&3 == 2): masked iff    i, j are in the same random group and both present at t = 0
// A row whose entities are all masked yields zeros (attention.py:58-60).
//
// Work decomposition: one CTA per (b, t); warp w owns head h = w (+ k*nwarps).  Phase A maps lanes to
// entities j (logits, warp-shuffle softmax), phase B maps lanes to the head feature k (weighted sum of V).
// The QKV tile (ne x 3d floats) is staged once in shared memory with rows padded by 4 floats so the
// 128-bit row reads of phase A are conflict-free.
#include "common.cuh"

#define ATT_THREADS 128
#define ATT_MAX_NE 32
#define ATT_MAX_COPIES 3

struct AttnArgs {
    const float* qkv;
    float* out;          // fwd: OUT ; bwd: unused
    const float* dout;   // bwd: dOUT [C, N, nq, d]
    float* dqkv;         // bwd: dQKV [N, ne, 3d]
    const uint8_t* mask[ATT_MAX_COPIES];
    long long mask_stride_n[ATT_MAX_COPIES];
    int mode[ATT_MAX_COPIES];
    const uint8_t* group_bits;   // [B, ne] or null
    const uint8_t* entity_mask;  // [N, ne] (N = B*T) or null
    int N, T, ne, nq, d, H, C;
};

__device__ __forceinline__ bool att_masked(const AttnArgs& a, int c, int n, int i, int j, int gi, int gj, int ina_i,
                                           int ina_j, int em_i, int em_j) {
    bool m = false;
    if (a.mask[c]) m = a.mask[c][(size_t)n * a.mask_stride_n[c] + (size_t)i * a.ne + j] != 0;
    const int mode = a.mode[c];
    const int part = mode & 3;
    if (part) {
        bool same = (gi == gj) && !ina_i && !ina_j;
        m = m || (part == 1 ? !same : same);
    }
    if (mode & 4) m = m || ina_i || ina_j;
    if (mode & 8) m = m || em_i || em_j;
    return m;
}

// stage the [ne, 3d] tile of unit n into smem rows of stride ld (floats)
__device__ __forceinline__ void att_load_tile(float* tile, const float* __restrict__ src, int ne, int w, int ld) {
    const int n4 = w >> 2;
    for (int f = threadIdx.x; f < ne * n4; f += ATT_THREADS) {
        int r = f / n4, c4 = f - r * n4;
        float4 v = __ldg(reinterpret_cast<const float4*>(src + (size_t)r * w) + c4);
        *reinterpret_cast<float4*>(tile + r * ld + c4 * 4) = v;
    }
}

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attn_fwd_kernel(AttnArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int d = a.d, ne = a.ne, nq = a.nq, ld = 3 * d + 4;
    float* tile = smem;                       // [ne][ld]
    float* sw = tile + ne * ld;               // [nwarps][32]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = ATT_THREADS / 32;
    const int n = blockIdx.x;
    att_load_tile(tile, a.qkv + (size_t)n * ne * 3 * d, ne, 3 * d, ld);
    // per-lane entity attributes (lane = entity j)
    const int b = n / a.T;
    int g_j = 0, ina_j = 0, em_j = 0;
    if (lane < ne) {
        if (a.group_bits) g_j = a.group_bits[(size_t)b * ne + lane];
        if (a.entity_mask) {
            ina_j = a.entity_mask[((size_t)b * a.T) * ne + lane];
            em_j = a.entity_mask[(size_t)n * ne + lane];
        }
    }
    __syncthreads();
    const float scale = sqrtf((float)HD);
    float* myw = sw + warp * 32;
    for (int h = warp; h < a.H; h += nwarps) {
        // lane j: K row of head h ; lane k: V column k of head h
        float kr[HD];
        if (lane < ne) {
#pragma unroll
            for (int k = 0; k < HD; k += 4) {
                float4 v = *reinterpret_cast<const float4*>(tile + lane * ld + d + h * HD + k);
                kr[k] = v.x; kr[k + 1] = v.y; kr[k + 2] = v.z; kr[k + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < HD; k++) kr[k] = 0.f;
        }
        float vc[ATT_MAX_NE];
#pragma unroll
        for (int j = 0; j < ATT_MAX_NE; j++) vc[j] = (j < ne && lane < HD) ? tile[j * ld + 2 * d + h * HD + lane] : 0.f;

        for (int i = 0; i < nq; i++) {
            float dot = 0.f;
#pragma unroll
            for (int k = 0; k < HD; k += 4) {
                float4 q = *reinterpret_cast<const float4*>(tile + i * ld + h * HD + k);
                dot = fmaf(q.x, kr[k], dot);
                dot = fmaf(q.y, kr[k + 1], dot);
                dot = fmaf(q.z, kr[k + 2], dot);
                dot = fmaf(q.w, kr[k + 3], dot);
            }
            const float logit0 = dot / scale;
            const int g_i = __shfl_sync(0xffffffffu, g_j, i), ina_i = __shfl_sync(0xffffffffu, ina_j, i),
                      em_i = __shfl_sync(0xffffffffu, em_j, i);
            for (int c = 0; c < a.C; c++) {
                float logit = -INFINITY;
                if (lane < ne && !att_masked(a, c, n, i, lane, g_i, g_j, ina_i, ina_j, em_i, em_j)) logit = logit0;
                const float m = warp_max(logit);
                const float p = (logit == -INFINITY) ? 0.f : expf(logit - m);
                const float s = warp_sum(p);
                myw[lane] = (s > 0.f) ? p / s : 0.f;
                __syncwarp();
                if (lane < HD) {
                    float acc = 0.f;
#pragma unroll
                    for (int j = 0; j < ATT_MAX_NE; j++)
                        if (j < ne) acc = fmaf(myw[j], vc[j], acc);
                    a.out[(((size_t)c * a.N + n) * nq + i) * d + h * HD + lane] = acc;
                }
                __syncwarp();
            }
        }
    }
}

// Backward: dQKV[n] = d/dQKV sum_c <dOUT[c, n], attn_c(QKV[n])>, accumulated over the copies in registers / smem.
template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_kernel(AttnArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int d = a.d, ne = a.ne, nq = a.nq, ld = 3 * d + 4, ldo = d + 4;
    float* tile = smem;                         // [ne][ld]   QKV, overwritten in place by dQKV
    float* sdo = tile + ne * ld;                // [C][nq][ldo]
    float* sw = sdo + a.C * nq * ldo;           // [nwarps][32]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = ATT_THREADS / 32;
    const int n = blockIdx.x;
    att_load_tile(tile, a.qkv + (size_t)n * ne * 3 * d, ne, 3 * d, ld);
    for (int c = 0; c < a.C; c++)
        att_load_tile(sdo + c * nq * ldo, a.dout + ((size_t)c * a.N + n) * nq * d, nq, d, ldo);
    const int b = n / a.T;
    int g_j = 0, ina_j = 0, em_j = 0;
    if (lane < ne) {
        if (a.group_bits) g_j = a.group_bits[(size_t)b * ne + lane];
        if (a.entity_mask) {
            ina_j = a.entity_mask[((size_t)b * a.T) * ne + lane];
            em_j = a.entity_mask[(size_t)n * ne + lane];
        }
    }
    __syncthreads();
    const float scale = sqrtf((float)HD), inv_scale = 1.f / scale;
    float* myw = sw + warp * 32;
    for (int h = warp; h < a.H; h += nwarps) {
        float kr[HD], vr[HD], dk[HD], dv[HD];
#pragma unroll
        for (int k = 0; k < HD; k++) { kr[k] = 0.f; vr[k] = 0.f; dk[k] = 0.f; dv[k] = 0.f; }
        if (lane < ne) {
#pragma unroll
            for (int k = 0; k < HD; k += 4) {
                float4 v = *reinterpret_cast<const float4*>(tile + lane * ld + d + h * HD + k);
                kr[k] = v.x; kr[k + 1] = v.y; kr[k + 2] = v.z; kr[k + 3] = v.w;
                float4 u = *reinterpret_cast<const float4*>(tile + lane * ld + 2 * d + h * HD + k);
                vr[k] = u.x; vr[k + 1] = u.y; vr[k + 2] = u.z; vr[k + 3] = u.w;
            }
        }
        for (int i = 0; i < nq; i++) {
            float qr[HD];
            float dot = 0.f;
#pragma unroll
            for (int k = 0; k < HD; k += 4) {
                float4 q = *reinterpret_cast<const float4*>(tile + i * ld + h * HD + k);
                qr[k] = q.x; qr[k + 1] = q.y; qr[k + 2] = q.z; qr[k + 3] = q.w;
                dot = fmaf(q.x, kr[k], dot);
                dot = fmaf(q.y, kr[k + 1], dot);
                dot = fmaf(q.z, kr[k + 2], dot);
                dot = fmaf(q.w, kr[k + 3], dot);
            }
            const float logit0 = dot / scale;
            const int g_i = __shfl_sync(0xffffffffu, g_j, i), ina_i = __shfl_sync(0xffffffffu, ina_j, i),
                      em_i = __shfl_sync(0xffffffffu, em_j, i);
            float dq = 0.f;  // lane k: dQ[i, h*HD + k] summed over the copies
            for (int c = 0; c < a.C; c++) {
                float logit = -INFINITY;
                if (lane < ne && !att_masked(a, c, n, i, lane, g_i, g_j, ina_i, ina_j, em_i, em_j)) logit = logit0;
                const float m = warp_max(logit);
                const float p = (logit == -INFINITY) ? 0.f : expf(logit - m);
                const float s = warp_sum(p);
                const float w = (s > 0.f) ? p / s : 0.f;
                const float* dor = sdo + (c * nq + i) * ldo + h * HD;
                float dw = 0.f;
#pragma unroll
                for (int k = 0; k < HD; k += 4) {
                    float4 g = *reinterpret_cast<const float4*>(dor + k);
                    dw = fmaf(g.x, vr[k], dw);
                    dw = fmaf(g.y, vr[k + 1], dw);
                    dw = fmaf(g.z, vr[k + 2], dw);
                    dw = fmaf(g.w, vr[k + 3], dw);
                    dv[k] = fmaf(w, g.x, dv[k]);
                    dv[k + 1] = fmaf(w, g.y, dv[k + 1]);
                    dv[k + 2] = fmaf(w, g.z, dv[k + 2]);
                    dv[k + 3] = fmaf(w, g.w, dv[k + 3]);
                }
                const float tsum = warp_sum(w * dw);
                const float dl = w * (dw - tsum) * inv_scale;
#pragma unroll
                for (int k = 0; k < HD; k++) dk[k] = fmaf(dl, qr[k], dk[k]);
                myw[lane] = dl;
                __syncwarp();
                if (lane < HD) {
                    for (int j = 0; j < ne; j++) dq = fmaf(myw[j], tile[j * ld + d + h * HD + lane], dq);
                }
                __syncwarp();
            }
            // Q_i of this head is dead for this warp from here on: overwrite with dQ_i
            if (lane < HD) tile[i * ld + h * HD + lane] = dq;
            __syncwarp();
        }
        // K/V of this head are dead for this warp: overwrite rows with dK / dV, zero the unused Q rows
        __syncwarp();
        if (lane < ne) {
#pragma unroll
            for (int k = 0; k < HD; k += 4) {
                *reinterpret_cast<float4*>(tile + lane * ld + d + h * HD + k) = make_float4(dk[k], dk[k + 1], dk[k + 2], dk[k + 3]);
                *reinterpret_cast<float4*>(tile + lane * ld + 2 * d + h * HD + k) = make_float4(dv[k], dv[k + 1], dv[k + 2], dv[k + 3]);
            }
            if (lane >= nq) {
#pragma unroll
                for (int k = 0; k < HD; k += 4)
                    *reinterpret_cast<float4*>(tile + lane * ld + h * HD + k) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    __syncthreads();
    // coalesced write-back of the dQKV tile
    {
        const int w = 3 * d, n4 = w >> 2;
        float* dst = a.dqkv + (size_t)n * ne * w;
        for (int f = threadIdx.x; f < ne * n4; f += ATT_THREADS) {
            int r = f / n4, c4 = f - r * n4;
            float4 v = *reinterpret_cast<const float4*>(tile + r * ld + c4 * 4);
            *(reinterpret_cast<float4*>(dst + (size_t)r * w) + c4) = v;
        }
    }
}

static int attn_fill(AttnArgs& a, const float* qkv, const uint8_t* m0, const uint8_t* m1, const uint8_t* m2,
                     long long s0, long long s1, long long s2, int mode0, int mode1, int mode2,
                     const uint8_t* group_bits, const uint8_t* entity_mask, int N, int T, int ne, int nq, int d, int H,
                     int C) {
    REFIL_CHECK_ARG(qkv && N > 0 && T > 0 && N % T == 0, "masked_attn: bad N=%d T=%d", N, T);
    REFIL_CHECK_ARG(ne >= 1 && ne <= ATT_MAX_NE, "masked_attn: n_entities %d outside [1,%d]", ne, ATT_MAX_NE);
    REFIL_CHECK_ARG(nq >= 1 && nq <= ne, "masked_attn: n_queries %d outside [1,%d]", nq, ne);
    REFIL_CHECK_ARG(C >= 1 && C <= ATT_MAX_COPIES, "masked_attn: copies %d outside [1,%d]", C, ATT_MAX_COPIES);
    REFIL_CHECK_ARG(H >= 1 && d % H == 0 && d % 4 == 0, "masked_attn: embed %d / heads %d", d, H);
    const int hd = d / H;
    REFIL_CHECK_ARG(hd == 8 || hd == 16 || hd == 32, "masked_attn: head dim %d not in {8,16,32}", hd);
    int modes[3] = {mode0, mode1, mode2};
    for (int c = 0; c < C; c++) {
        REFIL_CHECK_ARG((modes[c] & ~15) == 0 && (modes[c] & 3) != 3, "masked_attn: bad mode %d", modes[c]);
        REFIL_CHECK_ARG(!(modes[c] & 3) || group_bits, "masked_attn: partition mode needs group_bits");
        REFIL_CHECK_ARG(!(modes[c] & 15) || entity_mask, "masked_attn: mode %d needs entity_mask", modes[c]);
    }
    a.qkv = qkv;
    a.mask[0] = m0; a.mask[1] = m1; a.mask[2] = m2;
    a.mask_stride_n[0] = s0; a.mask_stride_n[1] = s1; a.mask_stride_n[2] = s2;
    a.mode[0] = mode0; a.mode[1] = mode1; a.mode[2] = mode2;
    a.group_bits = group_bits; a.entity_mask = entity_mask;
    a.N = N; a.T = T; a.ne = ne; a.nq = nq; a.d = d; a.H = H; a.C = C;
    return REFIL_OK;
}

template <class K>
static int attn_launch(K kernel, const AttnArgs& a, size_t smem, cudaStream_t stream, const char* name) {
    if (smem > 227 * 1024) {
        refil_set_error("%s: tile needs %zu bytes of shared memory (> 227 KB)", name, smem);
        return REFIL_ERR_UNSUPPORTED;
    }
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            refil_set_error("%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
            return REFIL_ERR_CUDA;
        }
    }
    kernel<<<a.N, ATT_THREADS, smem, stream>>>(a);
    REFIL_CHECK_LAUNCH(name);
    return REFIL_OK;
}

extern "C" int refil_masked_attn_fwd(const float* qkv, float* out, const uint8_t* mask0, const uint8_t* mask1,
                                     const uint8_t* mask2, long long mask_stride0, long long mask_stride1,
                                     long long mask_stride2, int mode0, int mode1, int mode2,
                                     const uint8_t* group_bits, const uint8_t* entity_mask, int N, int T,
                                     int n_entities, int n_queries, int embed_dim, int n_heads, int n_copies,
                                     cudaStream_t stream) {
    AttnArgs a{};
    int rc = attn_fill(a, qkv, mask0, mask1, mask2, mask_stride0, mask_stride1, mask_stride2, mode0, mode1, mode2,
                       group_bits, entity_mask, N, T, n_entities, n_queries, embed_dim, n_heads, n_copies);
    if (rc) return rc;
    REFIL_CHECK_ARG(out != nullptr, "masked_attn_fwd: out is null");
    a.out = out;
    size_t smem = ((size_t)n_entities * (3 * embed_dim + 4) + (ATT_THREADS / 32) * 32) * sizeof(float);
    switch (embed_dim / n_heads) {
        case 8: return attn_launch(attn_fwd_kernel<8>, a, smem, stream, "masked_attn_fwd");
        case 16: return attn_launch(attn_fwd_kernel<16>, a, smem, stream, "masked_attn_fwd");
        default: return attn_launch(attn_fwd_kernel<32>, a, smem, stream, "masked_attn_fwd");
    }
}

extern "C" int refil_masked_attn_bwd(const float* qkv, const float* dout, float* dqkv, const uint8_t* mask0,
                                     const uint8_t* mask1, const uint8_t* mask2, long long mask_stride0,
                                     long long mask_stride1, long long mask_stride2, int mode0, int mode1, int mode2,
                                     const uint8_t* group_bits, const uint8_t* entity_mask, int N, int T,
                                     int n_entities, int n_queries, int embed_dim, int n_heads, int n_copies,
                                     cudaStream_t stream) {
    AttnArgs a{};
    int rc = attn_fill(a, qkv, mask0, mask1, mask2, mask_stride0, mask_stride1, mask_stride2, mode0, mode1, mode2,
                       group_bits, entity_mask, N, T, n_entities, n_queries, embed_dim, n_heads, n_copies);
    if (rc) return rc;
    REFIL_CHECK_ARG(dout && dqkv, "masked_attn_bwd: dout / dqkv is null");
    a.dout = dout;
    a.dqkv = dqkv;
    size_t smem = ((size_t)n_entities * (3 * embed_dim + 4) + (size_t)n_copies * n_queries * (embed_dim + 4) +
                   (ATT_THREADS / 32) * 32) * sizeof(float);
    switch (embed_dim / n_heads) {
        case 8: return attn_launch(attn_bwd_kernel<8>, a, smem, stream, "masked_attn_bwd");
        case 16: return attn_launch(attn_bwd_kernel<16>, a, smem, stream, "masked_attn_bwd");
        default: return attn_launch(attn_bwd_kernel<32>, a, smem, stream, "masked_attn_bwd");
    }
}
