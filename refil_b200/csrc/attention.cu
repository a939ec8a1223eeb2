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

// ---- TMA (bulk async copy) staging of one (b, t) unit: ne row copies of 3d floats into padded smem rows ----------
__device__ __forceinline__ uint32_t att_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void att_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(att_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void att_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = att_smem_u32(bar), done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
// called by one full warp: lane 0 arms the barrier with the byte count, lanes < rows issue one row copy each
__device__ __forceinline__ void att_tma_load_rows(float* tile, const float* src, int rows, int w, int ld, uint64_t* bar,
                                                  int lane) {
    const uint32_t bytes = (uint32_t)w * 4u;
    if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(att_smem_u32(bar)), "r"(bytes * (uint32_t)rows)
                     : "memory");
    }
    __syncwarp();
    if (lane < rows) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         att_smem_u32(tile + lane * ld)),
                     "l"(src + (size_t)lane * w), "r"(bytes), "r"(att_smem_u32(bar))
                     : "memory");
    }
}

// mask bits of lane j for every query row i and copy c (bit i set = masked); the masks do not depend on the head
__device__ __forceinline__ void att_mask_bits(const AttnArgs& a, int n, int lane, int g_j, int ina_j, int em_j,
                                              uint32_t mb[ATT_MAX_COPIES]) {
#pragma unroll
    for (int c = 0; c < ATT_MAX_COPIES; c++) mb[c] = 0;
    for (int i = 0; i < a.nq; i++) {
        const int g_i = __shfl_sync(0xffffffffu, g_j, i), ina_i = __shfl_sync(0xffffffffu, ina_j, i),
                  em_i = __shfl_sync(0xffffffffu, em_j, i);
#pragma unroll
        for (int c = 0; c < ATT_MAX_COPIES; c++) {
            if (c < a.C) {
                bool m = lane >= a.ne || att_masked(a, c, n, i, lane, g_i, g_j, ina_i, ina_j, em_i, em_j);
                mb[c] |= (m ? 1u : 0u) << i;
            }
        }
    }
}

// softmax of 8 query rows at once: lane = (row i = lane >> 2, quarter s = lane & 3), 8 logits per lane; rows whose
// entities are all masked produce zeros (attention.py:58-60)
__device__ __forceinline__ void att_softmax8(float* sl, int lane) {
    float* row = sl + (lane >> 2) * 36 + (lane & 3) * 8;
    float4 v0 = *reinterpret_cast<const float4*>(row), v1 = *reinterpret_cast<const float4*>(row + 4);
    float m = fmaxf(fmaxf(fmaxf(v0.x, v0.y), fmaxf(v0.z, v0.w)), fmaxf(fmaxf(v1.x, v1.y), fmaxf(v1.z, v1.w)));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    const float mm = (m == -INFINITY) ? 0.f : m;          // exp(-inf - 0) = 0 for masked entries
    v0.x = __expf(v0.x - mm); v0.y = __expf(v0.y - mm); v0.z = __expf(v0.z - mm); v0.w = __expf(v0.w - mm);
    v1.x = __expf(v1.x - mm); v1.y = __expf(v1.y - mm); v1.z = __expf(v1.z - mm); v1.w = __expf(v1.w - mm);
    float sum = ((v0.x + v0.y) + (v0.z + v0.w)) + ((v1.x + v1.y) + (v1.z + v1.w));
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float r = sum > 0.f ? 1.f / sum : 0.f;
    v0.x *= r; v0.y *= r; v0.z *= r; v0.w *= r; v1.x *= r; v1.y *= r; v1.z *= r; v1.w *= r;
    *reinterpret_cast<float4*>(row) = v0;
    *reinterpret_cast<float4*>(row + 4) = v1;
}

// Persistent CTAs walk the (b, t) units; the QKV tile of the next unit is in flight (TMA bulk copies + mbarrier) while
// the current one is processed.  Per head: phase A lanes = entities (Q.K dot products, 8 query rows per pass),
// phase S lanes = (row, quarter) softmax of the 8 rows at once, phase B lanes = head features (weights x V).
template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attn_fwd_kernel(AttnArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int d = a.d, ne = a.ne, nq = a.nq, ld = 3 * d + 4;
    float* tiles = smem;                              // [2][ne][ld]
    float* slog = tiles + 2 * ne * ld;                // [nwarps][8][36]
    __shared__ uint64_t full_bar[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = ATT_THREADS / 32;
    if (threadIdx.x == 0) {
        att_mbar_init(&full_bar[0], 1);
        att_mbar_init(&full_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0 && (int)blockIdx.x < a.N)
        att_tma_load_rows(tiles, a.qkv + (size_t)blockIdx.x * ne * 3 * d, ne, 3 * d, ld, &full_bar[0], lane);
    const float inv_scale = 1.f / sqrtf((float)HD);
    float* sl = slog + warp * (8 * 36);
    int it = 0;
    for (int n = blockIdx.x; n < a.N; n += gridDim.x, it++) {
        const int st = it & 1;
        const int n_next = n + gridDim.x;
        if (warp == 0 && n_next < a.N)               // stage st^1 was released by the __syncthreads of the last unit
            att_tma_load_rows(tiles + (st ^ 1) * ne * ld, a.qkv + (size_t)n_next * ne * 3 * d, ne, 3 * d, ld,
                              &full_bar[st ^ 1], lane);
        const int b = n / a.T;
        int g_j = 0, ina_j = 0, em_j = 0;
        if (lane < ne) {
            if (a.group_bits) g_j = a.group_bits[(size_t)b * ne + lane];
            if (a.entity_mask) {
                ina_j = a.entity_mask[((size_t)b * a.T) * ne + lane];
                em_j = a.entity_mask[(size_t)n * ne + lane];
            }
        }
        uint32_t mb[ATT_MAX_COPIES];
        att_mask_bits(a, n, lane, g_j, ina_j, em_j, mb);
        att_mbar_wait(&full_bar[st], (it >> 1) & 1);
        const float* tile = tiles + st * ne * ld;
        for (int h = warp; h < a.H; h += nwarps) {
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
            for (int ib = 0; ib < nq; ib += 8) {
                float lg[8];
#pragma unroll
                for (int ii = 0; ii < 8; ii++) {
                    float dot = 0.f;
                    if (ib + ii < nq) {
                        const float* qrow = tile + (ib + ii) * ld + h * HD;
#pragma unroll
                        for (int k = 0; k < HD; k += 4) {
                            float4 q = *reinterpret_cast<const float4*>(qrow + k);
                            dot = fmaf(q.x, kr[k], dot);
                            dot = fmaf(q.y, kr[k + 1], dot);
                            dot = fmaf(q.z, kr[k + 2], dot);
                            dot = fmaf(q.w, kr[k + 3], dot);
                        }
                    }
                    lg[ii] = dot * inv_scale;
                }
                for (int c = 0; c < a.C; c++) {
                    const uint32_t bits = mb[c] >> ib;
#pragma unroll
                    for (int ii = 0; ii < 8; ii++) sl[ii * 36 + lane] = ((bits >> ii) & 1u) ? -INFINITY : lg[ii];
                    __syncwarp();
                    att_softmax8(sl, lane);
                    __syncwarp();
                    if (lane < HD) {
#pragma unroll
                        for (int ii = 0; ii < 8; ii++) {
                            if (ib + ii < nq) {
                                float acc = 0.f;
#pragma unroll
                                for (int j4 = 0; j4 < ATT_MAX_NE / 4; j4++) {
                                    if (4 * j4 < ne) {
                                        const float4 w4 = *reinterpret_cast<const float4*>(sl + ii * 36 + 4 * j4);
                                        acc = fmaf(w4.x, vc[4 * j4], acc);
                                        acc = fmaf(w4.y, vc[4 * j4 + 1], acc);
                                        acc = fmaf(w4.z, vc[4 * j4 + 2], acc);
                                        acc = fmaf(w4.w, vc[4 * j4 + 3], acc);
                                    }
                                }
                                a.out[(((size_t)c * a.N + n) * nq + ib + ii) * d + h * HD + lane] = acc;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
        }
        __syncthreads();      // every warp is done with stage st before it is refilled two units later
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
static int attn_launch(K kernel, const AttnArgs& a, size_t smem, int grid, cudaStream_t stream, const char* name) {
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
    kernel<<<grid, ATT_THREADS, smem, stream>>>(a);
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
    REFIL_CHECK_ARG(((3 * embed_dim * 4) % 16) == 0 && ((uintptr_t)qkv % 16) == 0, "masked_attn_fwd: QKV rows must be 16-byte aligned");
    size_t smem = ((size_t)2 * n_entities * (3 * embed_dim + 4) + (ATT_THREADS / 32) * 8 * 36) * sizeof(float);
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    int grid = per_sm * refil_num_sms();
    if (grid > N) grid = N;
    switch (embed_dim / n_heads) {
        case 8: return attn_launch(attn_fwd_kernel<8>, a, smem, grid, stream, "masked_attn_fwd");
        case 16: return attn_launch(attn_fwd_kernel<16>, a, smem, grid, stream, "masked_attn_fwd");
        default: return attn_launch(attn_fwd_kernel<32>, a, smem, grid, stream, "masked_attn_fwd");
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
        case 8: return attn_launch(attn_bwd_kernel<8>, a, smem, N, stream, "masked_attn_bwd");
        case 16: return attn_launch(attn_bwd_kernel<16>, a, smem, N, stream, "masked_attn_bwd");
        default: return attn_launch(attn_bwd_kernel<32>, a, smem, N, stream, "masked_attn_bwd");
    }
}
