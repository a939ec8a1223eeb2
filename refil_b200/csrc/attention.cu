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
// Work decomposition: every WARP is an independent persistent worker that walks (b, t) units.  The K and V rows of a unit
// (ne x d floats each) are staged in the warp's shared-memory tile by TMA tensor copies (cp.async.bulk.tensor + mbarrier);
// in the forward kernel K and V are separate transactions so that the next unit's K lands while this unit's softmax / PV
// phase runs and its V while the next logits run.  The Q row and the mask bytes of the next unit are requested one unit
// ahead into registers.  Lane = (agent i, head h): each thread owns one attention row, so logits, the masked softmax and
// the weighted sum over V never leave its registers (no shuffles, no shared-memory round trips); the masks are resolved
// once per unit as entity sets with warp ballots (lane = entity j).  Threads of different heads read K/V chunks in a
// head-rotated order so that the four distinct broadcast addresses of a warp load hit different banks.  The tile pointer
// is derived by pointer arithmetic on the __shared__ symbol and the head count is a template parameter, so every tile
// access is an LDS.128 with an immediate offset.
#include <cuda.h>   // CUtensorMap (the encode function is fetched through cudaGetDriverEntryPoint, libcuda is not linked)

#include <stdlib.h>
#include <type_traits>
#include <string.h>

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


#define ATT_MAX_WARPS 8
#define ATT_MAX_GROUP 8
// a launch carries up to ATT_MAX_GROUP independent problems of one geometry (the attention of several networks over the same
// (b, t) units): blockIdx.y selects the problem
struct AttnGroup { AttnArgs a[ATT_MAX_GROUP]; int cta_begin[ATT_MAX_GROUP + 1]; };
struct AttnMaps { CUtensorMap m[ATT_MAX_GROUP]; };

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
// one full warp: lane 0 arms the barrier with the byte count, lanes < rows issue one bulk row copy each (TMA, 1-D)
__device__ __forceinline__ void att_tma_load_rows(float* tile, const float* src, long long src_stride, int rows, int w,
                                                  int ld, uint64_t* bar, int lane) {
    const uint32_t bytes = (uint32_t)w * 4u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic accesses of the tile are ordered first
    __syncwarp();
    if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(att_smem_u32(bar)), "r"(bytes * (uint32_t)rows)
                     : "memory");
    }
    __syncwarp();
    if (lane < rows) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         att_smem_u32(tile + lane * ld)),
                     "l"(src + (size_t)lane * src_stride), "r"(bytes), "r"(att_smem_u32(bar))
                     : "memory");
    }
}

// one TMA tensor copy of the whole K|V tile of unit n: box (2d, ne, 1) of the [N][ne][3d] tensor at column d
__device__ __forceinline__ void att_tma_load_tile(float* tile, const CUtensorMap* tmap, int col, int ne, long long n,
                                                  uint64_t* bar, int lane, int box_cols) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        const uint32_t bytes = (uint32_t)(ne * box_cols) * 4u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(att_smem_u32(bar)), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"(att_smem_u32(tile)), "l"(tmap), "r"(col), "r"(0), "r"((int)n), "r"(att_smem_u32(bar))
            : "memory");
    }
}

// Mask resolution, split in two so that the global bytes of the NEXT unit can be requested a whole unit of compute ahead:
//   att_meta_load    : lane = entity j requests its group bit, t=0 / current inactive flags and the explicit-mask bytes of
//                      the 8 query rows of the pass (copies that share copy 0's tensor load nothing);
//   att_meta_resolve : warp ballots turn them into entity SETS (inactive at t=0, masked now, active members of group 0 / 1,
//                      one word per explicit row); every lane then assembles the mask word of ITS query row for each copy
//                      with plain bit algebra:  W (mode&3 == 1): everything outside my group, I (== 2): my group,
//                      ACTIVE0 (4) / DEFAULT (8): the inactive sets, all-ones if my own row is inactive; bit j set = masked.
struct AttMeta {
    uint8_t v[ATT_MAX_COPIES][8];
    int g, ina, em;
};

__device__ __forceinline__ bool att_owns_mask(const AttnArgs& a, int c) {
    return c < a.C && a.mask[c] != nullptr &&
           !(c > 0 && a.mask[c] == a.mask[0] && a.mask_stride_n[c] == a.mask_stride_n[0]);
}

__device__ __forceinline__ void att_meta_load(const AttnArgs& a, long long n, int lane, int ib, AttMeta& m) {
    m.g = 0; m.ina = 0; m.em = 0;
#pragma unroll
    for (int c = 0; c < ATT_MAX_COPIES; c++)
#pragma unroll
        for (int r = 0; r < 8; r++) m.v[c][r] = 0;
    if (lane < a.ne) {
        const long long b = n / a.T;
        if (a.group_bits) m.g = a.group_bits[(size_t)b * a.ne + lane];
        if (a.entity_mask) {
            m.ina = a.entity_mask[((size_t)b * a.T) * a.ne + lane];
            m.em = a.entity_mask[(size_t)n * a.ne + lane];
        }
#pragma unroll
        for (int c = 0; c < ATT_MAX_COPIES; c++) {
            if (att_owns_mask(a, c)) {
                const uint8_t* mp = a.mask[c] + (size_t)n * a.mask_stride_n[c] + (size_t)ib * a.ne + lane;
#pragma unroll
                for (int r = 0; r < 8; r++)
                    if (ib + r < a.nq) m.v[c][r] = mp[(size_t)r * a.ne];
            }
        }
    }
}

__device__ __forceinline__ void att_meta_resolve(const AttnArgs& a, long long n, int lane, int ib, int ipp, int il,
                                                 const AttMeta& m, uint32_t mb[ATT_MAX_COPIES]) {
    const bool in_range = lane < a.ne;
    const uint32_t PAD = ~__ballot_sync(0xffffffffu, in_range);
    const uint32_t INA = __ballot_sync(0xffffffffu, m.ina != 0), EM = __ballot_sync(0xffffffffu, m.em != 0);
    const uint32_t G1 = __ballot_sync(0xffffffffu, in_range && m.g != 0 && m.ina == 0);
    const uint32_t G0 = __ballot_sync(0xffffffffu, in_range && m.g == 0 && m.ina == 0);
    const int i = ib + il;                                        // my query row (< 32)
    const bool ina_i = (INA >> i) & 1u, em_i = (EM >> i) & 1u;
    const uint32_t same = ((G1 >> i) & 1u) ? G1 : (((G0 >> i) & 1u) ? G0 : 0u);   // active entities of my row's group
    uint32_t exw[ATT_MAX_COPIES];                                 // explicit mask word of my row
#pragma unroll
    for (int c = 0; c < ATT_MAX_COPIES; c++) {
        exw[c] = 0;
        if (att_owns_mask(a, c)) {
#pragma unroll
            for (int r = 0; r < 8; r++) {
                if (r < ipp) {                                    // warp-uniform
                    const uint32_t w = __ballot_sync(0xffffffffu, m.v[c][r] != 0);
                    if (il == r) exw[c] = w;
                }
            }
            for (int r = 8; r < ipp; r++) {                       // heads <= 2: more than 8 rows per pass (rare)
                const bool bit = in_range && ib + r < a.nq &&
                                 a.mask[c][(size_t)n * a.mask_stride_n[c] + (size_t)(ib + r) * a.ne + lane] != 0;
                const uint32_t w = __ballot_sync(0xffffffffu, bit);
                if (il == r) exw[c] = w;
            }
        } else if (c > 0 && c < a.C && a.mask[c] != nullptr) {
            exw[c] = exw[0];
        }
    }
#pragma unroll
    for (int c = 0; c < ATT_MAX_COPIES; c++) {
        uint32_t word = 0xffffffffu;
        if (c < a.C && i < a.nq) {
            const int mode = a.mode[c], part = mode & 3;
            word = PAD | exw[c];
            if (part == 1) word |= ~same;
            if (part == 2) word |= same;
            if (mode & 4) word |= INA | (ina_i ? 0xffffffffu : 0u);
            if (mode & 8) word |= EM | (em_i ? 0xffffffffu : 0u);
        }
        mb[c] = word;
    }
}

// head dim <= 16 (d = 64 configs: small units, latency-bound): registers capped so that two CTAs share an SM
template <int HD, int H, bool GROUPED>
__global__ void __launch_bounds__(32 * ATT_MAX_WARPS, (HD <= 16) ? 2 : 1) attn_fwd_kernel(const __grid_constant__ AttnGroup grp, int NEB, int tile_floats, int warp_floats, const __grid_constant__ AttnMaps maps, int use_tmap) {
    // grouped launch: the CTAs are dealt out to the problems in proportion to their work (grp.cta_begin); a single-problem launch
    // reads its arguments at fixed parameter offsets
    int prob = 0, cta = (int)blockIdx.x, ncta = (int)gridDim.x;
    if (GROUPED) {
        while (prob + 1 < ATT_MAX_GROUP && (int)blockIdx.x >= grp.cta_begin[prob + 1]) prob++;
        cta = (int)blockIdx.x - grp.cta_begin[prob];
        ncta = grp.cta_begin[prob + 1] - grp.cta_begin[prob];
    }
    const AttnArgs& a = grp.a[prob];
    const CUtensorMap& tmap = maps.m[prob];
    extern __shared__ __align__(128) float smem_raw_[];
    // 128-byte aligned base, derived by pointer arithmetic on the __shared__ symbol so that every access below is
    // provably shared memory (LDS / STS, not generic LD / ST)
    float* smem = smem_raw_ + (((128u - (att_smem_u32(smem_raw_) & 127u)) & 127u) >> 2);
    constexpr int NCH = HD / 4, d = HD * H;               // compile-time row stride: tile offsets become immediates
    const int ne = a.ne, nq = a.nq;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    // The K tile and the V tile of a unit are separate TMA transactions with separate barriers: K of the NEXT unit is
    // requested as soon as the logits of this unit are done (it lands during the softmax / PV phase), V of the next unit at
    // the end (it lands during the next unit's logits) -- the tile latency is hidden without a second tile buffer.
    float* kt = smem + (size_t)warp * warp_floats;                       // [NEB][d]: K rows of the current unit
    float* vt = kt + NEB * d;                                            // [NEB][d]: V rows
    float* lgs = kt + tile_floats + lane;                                // [NEB][32]: my row of logits, lgs[j * 32]
    uint64_t* bark = reinterpret_cast<uint64_t*>(smem + (size_t)wpc * warp_floats) + 2 * warp;
    uint64_t* barv = bark + 1;
    if (lane == 0) {
        att_mbar_init(bark, 1);
        att_mbar_init(barv, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int f = ne * d + lane; f < NEB * d; f += 32) { kt[f] = 0.f; vt[f] = 0.f; }   // padding rows stay zero
    __syncwarp();
    pdl_launch_dependents();
    pdl_wait();                                          // QKV is the previous kernel's output
    auto load_k = [&](long long n) {
        if (use_tmap) att_tma_load_tile(kt, &tmap, d, ne, n, bark, lane, d);
        else att_tma_load_rows(kt, a.qkv + ((size_t)n * ne) * 3 * d + d, 3 * d, ne, d, d, bark, lane);
    };
    auto load_v = [&](long long n) {
        if (use_tmap) att_tma_load_tile(vt, &tmap, 2 * d, ne, n, barv, lane, d);
        else att_tma_load_rows(vt, a.qkv + ((size_t)n * ne) * 3 * d + 2 * d, 3 * d, ne, d, d, barv, lane);
    };
    constexpr int ipp = 32 / H;
    const int h = lane % H, il = lane / H;
    int rot[NCH];                                        // float offset of my kc-th chunk inside a row (head-rotated)
#pragma unroll
    for (int kc = 0; kc < NCH; kc++) rot[kc] = h * HD + 4 * ((kc + h) & (NCH - 1));
    const float inv_scale = 1.f / sqrtf((float)HD);
    const long long gw = (long long)cta * wpc + warp, GW = (long long)ncta * wpc;
    uint32_t parity = 0;
    // the Q row and the mask bytes of the NEXT unit are requested before the current one is computed
    auto load_q = [&](long long n, int i, float4 (&dst)[NCH]) {
        const float* qsrc = a.qkv + ((size_t)n * ne + i) * 3 * d;
#pragma unroll
        for (int kc = 0; kc < NCH; kc++) dst[kc] = __ldg(reinterpret_cast<const float4*>(qsrc + rot[kc]));
    };
    AttMeta mn;
    float4 qn[NCH];
    if (gw < a.N) {
        load_k(gw);
        load_v(gw);
        att_meta_load(a, gw, lane, 0, mn);
        load_q(gw, il < nq ? il : 0, qn);
    }
    for (long long n = gw; n < a.N; n += GW) {
        const AttMeta mc = mn;
        float4 qc[NCH];
#pragma unroll
        for (int kc = 0; kc < NCH; kc++) qc[kc] = qn[kc];
        if (n + GW < a.N) {
            att_meta_load(a, n + GW, lane, 0, mn);
            load_q(n + GW, il < nq ? il : 0, qn);
        }
        bool waited = false;
        for (int ib = 0; ib < nq; ib += ipp) {
            const int i = ib + il;
            const bool active = i < nq;
            uint32_t mb[ATT_MAX_COPIES];
            float q[HD];
            if (ib == 0) {
                att_meta_resolve(a, n, lane, 0, ipp, il, mc, mb);
#pragma unroll
                for (int kc = 0; kc < NCH; kc++) {
                    q[4 * kc] = qc[kc].x; q[4 * kc + 1] = qc[kc].y; q[4 * kc + 2] = qc[kc].z; q[4 * kc + 3] = qc[kc].w;
                }
            } else {                                               // further passes (nq > 32 / heads): loaded in place
                AttMeta mt;
                att_meta_load(a, n, lane, ib, mt);
                float4 qt[NCH];
                load_q(n, active ? i : 0, qt);
                att_meta_resolve(a, n, lane, ib, ipp, il, mt, mb);
#pragma unroll
                for (int kc = 0; kc < NCH; kc++) {
                    q[4 * kc] = qt[kc].x; q[4 * kc + 1] = qt[kc].y; q[4 * kc + 2] = qt[kc].z; q[4 * kc + 3] = qt[kc].w;
                }
            }
            const bool last_pass = ib + ipp >= nq;
            if (!waited) att_mbar_wait(bark, parity);
            const float* kbase = kt;
            const float* vbase = vt;
            float mx[ATT_MAX_COPIES] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll 4
            for (int j = 0; j < NEB; j++) {
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                for (int kc = 0; kc < NCH; kc++) {
                    const float4 k4 = *reinterpret_cast<const float4*>(kbase + j * d + rot[kc]);
                    s0 = fmaf(q[4 * kc], k4.x, s0);
                    s1 = fmaf(q[4 * kc + 1], k4.y, s1);
                    s2 = fmaf(q[4 * kc + 2], k4.z, s2);
                    s3 = fmaf(q[4 * kc + 3], k4.w, s3);
                }
                const float lg = ((s0 + s1) + (s2 + s3)) * inv_scale;
                lgs[j * 32] = lg;
#pragma unroll
                for (int c = 0; c < ATT_MAX_COPIES; c++)
                    if (!((mb[c] >> j) & 1u)) mx[c] = fmaxf(mx[c], lg);
            }
            if (last_pass) {                      // every lane is done with K: the next unit's K tile may land
                __syncwarp();
                if (n + GW < a.N) load_k(n + GW);
            }
            if (!waited) { att_mbar_wait(barv, parity); parity ^= 1; waited = true; }
#pragma unroll
            for (int c = 0; c < ATT_MAX_COPIES; c++) {
                if (c < a.C) {
                    const uint32_t bits = mb[c];
                    const float m = mx[c];
                    float acc[HD];
#pragma unroll
                    for (int k = 0; k < HD; k++) acc[k] = 0.f;
                    float ssum = 0.f;
#pragma unroll 4
                    for (int j = 0; j < NEB; j++) {
                        const float e = ((bits >> j) & 1u) ? 0.f : __expf(lgs[j * 32] - m);
                        ssum += e;
#pragma unroll
                        for (int kc = 0; kc < NCH; kc++) {
                            const float4 v4 = *reinterpret_cast<const float4*>(vbase + j * d + rot[kc]);
                            acc[4 * kc] = fmaf(e, v4.x, acc[4 * kc]);
                            acc[4 * kc + 1] = fmaf(e, v4.y, acc[4 * kc + 1]);
                            acc[4 * kc + 2] = fmaf(e, v4.z, acc[4 * kc + 2]);
                            acc[4 * kc + 3] = fmaf(e, v4.w, acc[4 * kc + 3]);
                        }
                    }
                    const float r = ssum > 0.f ? 1.f / ssum : 0.f;    // all-masked row -> zeros (attention.py:58-60)
                    if (active) {
                        float* dst = a.out + (((size_t)c * a.N + n) * nq + i) * d;
#pragma unroll
                        for (int kc = 0; kc < NCH; kc++)
                            *reinterpret_cast<float4*>(dst + rot[kc]) =
                                make_float4(acc[4 * kc] * r, acc[4 * kc + 1] * r, acc[4 * kc + 2] * r, acc[4 * kc + 3] * r);
                    }
                }
            }
        }
        __syncwarp();          // every lane is done with V before the next unit's copy lands in it
        if (n + GW < a.N) load_v(n + GW);
    }
}

// =====================================================================================================================
// Forward, specialised for the benchmark's geometry (4 heads of 32, <= 8 query rows): TWO agents per lane, ALL mask copies in one
// sweep over V.  ncu on attn_fwd_kernel<32, 4> (capture r3d): the shared-memory pipe is the busiest unit (57 % of peak wavefronts
// with one copy, 70 % with three), DRAM ~50 %.  lane = (agent pair ip, head h, half dh of the head dim) as in attn_bwd_h4_kernel:
// every 128-bit load of a K / V chunk feeds two agents, and the softmax numerators of the copies (own row: computed, partner row:
// one shuffle) multiply the same V chunk -- 8 instead of 16 (one copy) / 32 (three copies) 128-bit loads per entity row.
// Tile prefetch (K of the next unit during the PV sweep, V during the next logits) as in the generic kernel.
// =====================================================================================================================
template <bool GROUPED>
__global__ void __launch_bounds__(32 * ATT_MAX_WARPS, 1) attn_fwd_h4_kernel(const __grid_constant__ AttnGroup grp, int NEB, int tile_floats, int warp_floats, const __grid_constant__ AttnMaps maps, int use_tmap) {
    int prob = 0, cta = (int)blockIdx.x, ncta = (int)gridDim.x;
    if (GROUPED) {
        while (prob + 1 < ATT_MAX_GROUP && (int)blockIdx.x >= grp.cta_begin[prob + 1]) prob++;
        cta = (int)blockIdx.x - grp.cta_begin[prob];
        ncta = grp.cta_begin[prob + 1] - grp.cta_begin[prob];
    }
    const AttnArgs& a = grp.a[prob];
    const CUtensorMap& tmap = maps.m[prob];
    extern __shared__ __align__(128) float smem_raw_[];
    float* smem = smem_raw_ + (((128u - (att_smem_u32(smem_raw_) & 127u)) & 127u) >> 2);
    constexpr int HD = 32, H = 4, d = HD * H;
    const int ne = a.ne, nq = a.nq;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    const int dh = lane & 1, h = (lane >> 1) & 3, ip = lane >> 3;          // half of the head dim, head, agent pair (ip, ip + 4)
    const int i0 = ip, i1 = ip + 4, io = dh ? i1 : i0;                      // io: the row whose scalars this lane owns
    const bool act0 = i0 < nq, act1 = i1 < nq;
    int rot[4];                                                             // float offset of my kc-th chunk inside a row
#pragma unroll
    for (int kc = 0; kc < 4; kc++) rot[kc] = h * HD + dh * 16 + 4 * ((kc + h) & 3);
    float* kt = smem + (size_t)warp * warp_floats;                       // [NEB][d]: K rows of the current unit
    float* vt = kt + NEB * d;                                            // [NEB][d]: V rows
    float* lgs = kt + tile_floats + lane;                                // [NEB][32]: logits of my own row, lgs[j * 32]
    uint64_t* bark = reinterpret_cast<uint64_t*>(smem + (size_t)wpc * warp_floats) + 2 * warp;
    uint64_t* barv = bark + 1;
    if (lane == 0) {
        att_mbar_init(bark, 1);
        att_mbar_init(barv, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int f = ne * d + lane; f < NEB * d; f += 32) { kt[f] = 0.f; vt[f] = 0.f; }   // padding rows stay zero
    __syncwarp();
    pdl_launch_dependents();
    pdl_wait();                                          // QKV is the previous kernel's output
    auto load_k = [&](long long n) {
        if (use_tmap) att_tma_load_tile(kt, &tmap, d, ne, n, bark, lane, d);
        else att_tma_load_rows(kt, a.qkv + ((size_t)n * ne) * 3 * d + d, 3 * d, ne, d, d, bark, lane);
    };
    auto load_v = [&](long long n) {
        if (use_tmap) att_tma_load_tile(vt, &tmap, 2 * d, ne, n, barv, lane, d);
        else att_tma_load_rows(vt, a.qkv + ((size_t)n * ne) * 3 * d + 2 * d, 3 * d, ne, d, d, barv, lane);
    };
    const float inv_scale = 1.f / sqrtf((float)HD);
    const long long gw = (long long)cta * wpc + warp, GW = (long long)ncta * wpc;
    uint32_t parity = 0;
    // the Q rows and the mask bytes of the NEXT unit are requested before the current one is computed
    auto load_q = [&](long long n, float4 (&dst)[2][4]) {
        const float* q0 = a.qkv + ((size_t)n * ne + (act0 ? i0 : 0)) * 3 * d;
        const float* q1 = a.qkv + ((size_t)n * ne + (act1 ? i1 : 0)) * 3 * d;
#pragma unroll
        for (int kc = 0; kc < 4; kc++) {
            dst[0][kc] = __ldg(reinterpret_cast<const float4*>(q0 + rot[kc]));
            dst[1][kc] = __ldg(reinterpret_cast<const float4*>(q1 + rot[kc]));
        }
    };
    AttMeta mn;
    float4 qn[2][4];
    if (gw < a.N) {
        load_k(gw);
        load_v(gw);
        att_meta_load(a, gw, lane, 0, mn);
        load_q(gw, qn);
    }
    for (long long n = gw; n < a.N; n += GW) {
        const AttMeta mc = mn;
        float q0[16], q1[16];
#pragma unroll
        for (int kc = 0; kc < 4; kc++) {
            q0[4 * kc] = act0 ? qn[0][kc].x : 0.f; q0[4 * kc + 1] = act0 ? qn[0][kc].y : 0.f;
            q0[4 * kc + 2] = act0 ? qn[0][kc].z : 0.f; q0[4 * kc + 3] = act0 ? qn[0][kc].w : 0.f;
            q1[4 * kc] = act1 ? qn[1][kc].x : 0.f; q1[4 * kc + 1] = act1 ? qn[1][kc].y : 0.f;
            q1[4 * kc + 2] = act1 ? qn[1][kc].z : 0.f; q1[4 * kc + 3] = act1 ? qn[1][kc].w : 0.f;
        }
        if (n + GW < a.N) {
            att_meta_load(a, n + GW, lane, 0, mn);
            load_q(n + GW, qn);
        }
        uint32_t mb[ATT_MAX_COPIES];                                       // mask words of MY OWN row io
        att_meta_resolve(a, n, lane, 0, 8, io, mc, mb);
        att_mbar_wait(bark, parity);
        float mx[ATT_MAX_COPIES] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll 4
        for (int j = 0; j < NEB; j++) {
            float p0a = 0.f, p0b = 0.f, p1a = 0.f, p1b = 0.f;
#pragma unroll
            for (int kc = 0; kc < 4; kc++) {
                const float4 k4 = *reinterpret_cast<const float4*>(kt + j * d + rot[kc]);
                p0a = fmaf(q0[4 * kc], k4.x, p0a); p0b = fmaf(q0[4 * kc + 1], k4.y, p0b);
                p0a = fmaf(q0[4 * kc + 2], k4.z, p0a); p0b = fmaf(q0[4 * kc + 3], k4.w, p0b);
                p1a = fmaf(q1[4 * kc], k4.x, p1a); p1b = fmaf(q1[4 * kc + 1], k4.y, p1b);
                p1a = fmaf(q1[4 * kc + 2], k4.z, p1a); p1b = fmaf(q1[4 * kc + 3], k4.w, p1b);
            }
            float p0 = p0a + p0b, p1 = p1a + p1b;
            p0 += __shfl_xor_sync(0xffffffffu, p0, 1);
            p1 += __shfl_xor_sync(0xffffffffu, p1, 1);
            const float lg = (dh ? p1 : p0) * inv_scale;
            lgs[j * 32] = lg;
#pragma unroll
            for (int c = 0; c < ATT_MAX_COPIES; c++)
                if (!((mb[c] >> j) & 1u)) mx[c] = fmaxf(mx[c], lg);
        }
        __syncwarp();                             // every lane is done with K: the next unit's K tile may land
        if (n + GW < a.N) load_k(n + GW);
        att_mbar_wait(barv, parity);
        parity ^= 1;
        auto pv_pass = [&](auto nc_tag) {
            constexpr int NC = decltype(nc_tag)::value;
            float acc[NC][2][16];
            float ssum[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) {
                ssum[c] = 0.f;
#pragma unroll
                for (int k = 0; k < 16; k++) { acc[c][0][k] = 0.f; acc[c][1][k] = 0.f; }
            }
#pragma unroll 2
            for (int j = 0; j < NEB; j++) {
                const float lg = lgs[j * 32];
                float e0[NC], e1[NC];
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    const float e = ((mb[c] >> j) & 1u) ? 0.f : __expf(lg - mx[c]);   // numerator of my own row
                    ssum[c] += e;
                    const float eo = __shfl_xor_sync(0xffffffffu, e, 1);              // ... and of my partner's row
                    e0[c] = dh ? eo : e;
                    e1[c] = dh ? e : eo;
                }
#pragma unroll
                for (int kc = 0; kc < 4; kc++) {
                    const float4 v4 = *reinterpret_cast<const float4*>(vt + j * d + rot[kc]);
#pragma unroll
                    for (int c = 0; c < NC; c++) {
                        acc[c][0][4 * kc] = fmaf(e0[c], v4.x, acc[c][0][4 * kc]); acc[c][0][4 * kc + 1] = fmaf(e0[c], v4.y, acc[c][0][4 * kc + 1]);
                        acc[c][0][4 * kc + 2] = fmaf(e0[c], v4.z, acc[c][0][4 * kc + 2]); acc[c][0][4 * kc + 3] = fmaf(e0[c], v4.w, acc[c][0][4 * kc + 3]);
                        acc[c][1][4 * kc] = fmaf(e1[c], v4.x, acc[c][1][4 * kc]); acc[c][1][4 * kc + 1] = fmaf(e1[c], v4.y, acc[c][1][4 * kc + 1]);
                        acc[c][1][4 * kc + 2] = fmaf(e1[c], v4.z, acc[c][1][4 * kc + 2]); acc[c][1][4 * kc + 3] = fmaf(e1[c], v4.w, acc[c][1][4 * kc + 3]);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const float r = ssum[c] > 0.f ? 1.f / ssum[c] : 0.f;        // all-masked row -> zeros (attention.py:58-60)
                const float ro = __shfl_xor_sync(0xffffffffu, r, 1);
                const float r0 = dh ? ro : r, r1 = dh ? r : ro;
                if (act0) {
                    float* dst = a.out + (((size_t)c * a.N + n) * nq + i0) * d;
#pragma unroll
                    for (int kc = 0; kc < 4; kc++)
                        *reinterpret_cast<float4*>(dst + rot[kc]) = make_float4(acc[c][0][4 * kc] * r0, acc[c][0][4 * kc + 1] * r0,
                                                                                acc[c][0][4 * kc + 2] * r0, acc[c][0][4 * kc + 3] * r0);
                }
                if (act1) {
                    float* dst = a.out + (((size_t)c * a.N + n) * nq + i1) * d;
#pragma unroll
                    for (int kc = 0; kc < 4; kc++)
                        *reinterpret_cast<float4*>(dst + rot[kc]) = make_float4(acc[c][1][4 * kc] * r1, acc[c][1][4 * kc + 1] * r1,
                                                                                acc[c][1][4 * kc + 2] * r1, acc[c][1][4 * kc + 3] * r1);
                }
            }
        };
        if (a.C == 1) pv_pass(std::integral_constant<int, 1>{});
        else if (a.C == 2) pv_pass(std::integral_constant<int, 2>{});
        else pv_pass(std::integral_constant<int, 3>{});
        __syncwarp();          // every lane is done with V before the next unit's copy lands in it
        if (n + GW < a.N) load_v(n + GW);
    }
}

// Backward: dQKV[n] = d/dQKV sum_c <dOUT[c, n], attn_c(QKV[n])>.
//   phase 1, lane = (agent i, head h): recompute the row softmax, dw_j = <dO_i, V_j>, dlogit_j = w_j (dw_j - sum_j' w_j' dw_j')
//            / sqrt(hd); dQ_i accumulates in registers over the copies; w and dlogit go to the warp's smem scratch;
//   phase 2, lane = (entity j, head h): dK_j = sum_{c,i} dlogit_ij Q_i, dV_j = sum_{c,i} w_ij dO_i (Q / dO rows come
//            back through L1), written over the K|V tile, which is then streamed out as the K|V columns of dQKV.
template <int HD, int H, bool GROUPED>
__global__ void __launch_bounds__(32 * ATT_MAX_WARPS, (HD <= 16) ? 2 : 1) attn_bwd_kernel(const __grid_constant__ AttnGroup grp, int NEB, int tile_floats, int warp_floats, const __grid_constant__ AttnMaps maps, int use_tmap) {
    // grouped launch: the CTAs are dealt out to the problems in proportion to their work (grp.cta_begin); a single-problem launch
    // reads its arguments at fixed parameter offsets
    int prob = 0, cta = (int)blockIdx.x, ncta = (int)gridDim.x;
    if (GROUPED) {
        while (prob + 1 < ATT_MAX_GROUP && (int)blockIdx.x >= grp.cta_begin[prob + 1]) prob++;
        cta = (int)blockIdx.x - grp.cta_begin[prob];
        ncta = grp.cta_begin[prob + 1] - grp.cta_begin[prob];
    }
    const AttnArgs& a = grp.a[prob];
    const CUtensorMap& tmap = maps.m[prob];
    extern __shared__ __align__(128) float smem_raw_[];
    // 128-byte aligned base, derived by pointer arithmetic on the __shared__ symbol so that every access below is
    // provably shared memory (LDS / STS, not generic LD / ST)
    float* smem = smem_raw_ + (((128u - (att_smem_u32(smem_raw_) & 127u)) & 127u) >> 2);
    constexpr int NCH = HD / 4, d = HD * H, ldk = 2 * d;   // compile-time row strides: tile offsets become immediates
    const int ne = a.ne, nq = a.nq;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    constexpr int ipp = 32 / H;
    const int h = lane % H, il = lane / H;
    int rot[NCH];                                        // float offset of my kc-th chunk inside a row (head-rotated)
#pragma unroll
    for (int kc = 0; kc < NCH; kc++) rot[kc] = h * HD + 4 * ((kc + h) & (NCH - 1));
    const int nqp = (nq + ipp - 1) / ipp * ipp;                          // agent rows padded to whole passes
    float* kv = smem + (size_t)warp * warp_floats;                       // [NEB][ldk]
    const int sstr = H * nqp;                                            // scratch stride between entities j
    float* sw = kv + tile_floats;                                        // [C][NEB][H][nqp]  softmax weights
    float* sdl = sw + a.C * NEB * sstr;                                  // [C][NEB][H][nqp]  dw, then dlogits
    float* lgs = sdl + a.C * NEB * sstr + lane;                          // [NEB][32]         my row of logits
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)wpc * warp_floats) + warp;
    if (lane == 0) {
        att_mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    pdl_launch_dependents();
    pdl_wait();                                          // dOUT (and QKV) come from earlier kernels of the step
    const float inv_scale = 1.f / sqrtf((float)HD);
    const long long gw = (long long)cta * wpc + warp, GW = (long long)ncta * wpc;
    uint32_t parity = 0;
    // the Q row and the mask bytes of the NEXT unit are requested before the current one is computed
    auto load_q = [&](long long n, int i, float4 (&dst)[NCH]) {
        const float* qsrc = a.qkv + ((size_t)n * ne + i) * 3 * d;
#pragma unroll
        for (int kc = 0; kc < NCH; kc++) dst[kc] = __ldg(reinterpret_cast<const float4*>(qsrc + rot[kc]));
    };
    AttMeta mn;
    float4 qn[NCH];
    if (gw < a.N) {
        att_meta_load(a, gw, lane, 0, mn);
        load_q(gw, il < nq ? il : 0, qn);
    }
    for (long long n = gw; n < a.N; n += GW) {
        for (int f = ne * ldk + lane; f < NEB * ldk; f += 32) kv[f] = 0.f;   // padding rows (the tile is reused for dK|dV)
        if (use_tmap) att_tma_load_tile(kv, &tmap, d, ne, n, bar, lane, 2 * d);
        else att_tma_load_rows(kv, a.qkv + ((size_t)n * ne) * 3 * d + d, 3 * d, ne, 2 * d, ldk, bar, lane);
        const AttMeta mc = mn;
        float4 qc[NCH];
#pragma unroll
        for (int kc = 0; kc < NCH; kc++) qc[kc] = qn[kc];
        if (n + GW < a.N) {
            att_meta_load(a, n + GW, lane, 0, mn);
            load_q(n + GW, il < nq ? il : 0, qn);
        }
        // dQ of the non-query rows is zero
        {
            const int d4 = d >> 2;
            for (int f = lane; f < (ne - nq) * d4; f += 32) {
                const int r = nq + f / d4, c4 = f % d4;
                *reinterpret_cast<float4*>(a.dqkv + ((size_t)n * ne + r) * 3 * d + 4 * c4) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        bool waited = false;
        for (int ib = 0; ib < nq; ib += ipp) {
            const int i = ib + il;                                       // < nqp
            const bool active = i < nq;
            uint32_t mb[ATT_MAX_COPIES];
            float q[HD];
            if (ib == 0) {
                att_meta_resolve(a, n, lane, 0, ipp, il, mc, mb);
#pragma unroll
                for (int kc = 0; kc < NCH; kc++) {
                    q[4 * kc] = qc[kc].x; q[4 * kc + 1] = qc[kc].y; q[4 * kc + 2] = qc[kc].z; q[4 * kc + 3] = qc[kc].w;
                }
            } else {                                               // further passes (nq > 32 / heads): loaded in place
                AttMeta mt;
                att_meta_load(a, n, lane, ib, mt);
                float4 qt[NCH];
                load_q(n, active ? i : 0, qt);
                att_meta_resolve(a, n, lane, ib, ipp, il, mt, mb);
#pragma unroll
                for (int kc = 0; kc < NCH; kc++) {
                    q[4 * kc] = qt[kc].x; q[4 * kc + 1] = qt[kc].y; q[4 * kc + 2] = qt[kc].z; q[4 * kc + 3] = qt[kc].w;
                }
            }
            if (!active) {
#pragma unroll
                for (int k = 0; k < HD; k++) q[k] = 0.f;
            }
            if (!waited) { att_mbar_wait(bar, parity); parity ^= 1; waited = true; }
            const float* kbase = kv;
            float mx[ATT_MAX_COPIES] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll 4
            for (int j = 0; j < NEB; j++) {
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                for (int kc = 0; kc < NCH; kc++) {
                    const float4 k4 = *reinterpret_cast<const float4*>(kbase + j * ldk + rot[kc]);
                    s0 = fmaf(q[4 * kc], k4.x, s0);
                    s1 = fmaf(q[4 * kc + 1], k4.y, s1);
                    s2 = fmaf(q[4 * kc + 2], k4.z, s2);
                    s3 = fmaf(q[4 * kc + 3], k4.w, s3);
                }
                const float lg = ((s0 + s1) + (s2 + s3)) * inv_scale;
                lgs[j * 32] = lg;
#pragma unroll
                for (int c = 0; c < ATT_MAX_COPIES; c++)
                    if (!((mb[c] >> j) & 1u)) mx[c] = fmaxf(mx[c], lg);
            }
            float dq[HD];
#pragma unroll
            for (int k = 0; k < HD; k++) dq[k] = 0.f;
#pragma unroll
            for (int c = 0; c < ATT_MAX_COPIES; c++) {
                if (c < a.C) {
                    const uint32_t bits = mb[c];
                    const float m = mx[c];
                    float go[HD];                                      // my dO row of this copy, rotated like q
                    const float* gsrc = a.dout + (((size_t)c * a.N + n) * nq + (active ? i : 0)) * d;
#pragma unroll
                    for (int kc = 0; kc < NCH; kc++) {
                        const float4 v = active ? __ldg(reinterpret_cast<const float4*>(gsrc + rot[kc]))
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
                        go[4 * kc] = v.x; go[4 * kc + 1] = v.y; go[4 * kc + 2] = v.z; go[4 * kc + 3] = v.w;
                    }
                    float* swr = sw + ((size_t)c * NEB * H + h) * nqp + i;   // [j * sstr]
                    float* sdr = sdl + ((size_t)c * NEB * H + h) * nqp + i;
                    float ssum = 0.f;
#pragma unroll 4
                    for (int j = 0; j < NEB; j++) {
                        const float e = ((bits >> j) & 1u) ? 0.f : __expf(lgs[j * 32] - m);
                        ssum += e;
                        swr[j * sstr] = e;
                    }
                    const float r = ssum > 0.f ? 1.f / ssum : 0.f;
                    float t = 0.f;
#pragma unroll 4
                    for (int j = 0; j < NEB; j++) {
                        const float w = swr[j * sstr] * r;
                        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                        for (int kc = 0; kc < NCH; kc++) {
                            const float4 v4 = *reinterpret_cast<const float4*>(kbase + j * ldk + d + rot[kc]);
                            s0 = fmaf(go[4 * kc], v4.x, s0);
                            s1 = fmaf(go[4 * kc + 1], v4.y, s1);
                            s2 = fmaf(go[4 * kc + 2], v4.z, s2);
                            s3 = fmaf(go[4 * kc + 3], v4.w, s3);
                        }
                        const float dw = (s0 + s1) + (s2 + s3);
                        swr[j * sstr] = w;
                        sdr[j * sstr] = dw;
                        t = fmaf(w, dw, t);
                    }
#pragma unroll 4
                    for (int j = 0; j < NEB; j++) {
                        const float dl = swr[j * sstr] * (sdr[j * sstr] - t) * inv_scale;
                        sdr[j * sstr] = dl;
#pragma unroll
                        for (int kc = 0; kc < NCH; kc++) {
                            const float4 k4 = *reinterpret_cast<const float4*>(kbase + j * ldk + rot[kc]);
                            dq[4 * kc] = fmaf(dl, k4.x, dq[4 * kc]);
                            dq[4 * kc + 1] = fmaf(dl, k4.y, dq[4 * kc + 1]);
                            dq[4 * kc + 2] = fmaf(dl, k4.z, dq[4 * kc + 2]);
                            dq[4 * kc + 3] = fmaf(dl, k4.w, dq[4 * kc + 3]);
                        }
                    }
                }
            }
            if (active) {
                float* dst = a.dqkv + ((size_t)n * ne + i) * 3 * d;
#pragma unroll
                for (int kc = 0; kc < NCH; kc++)
                    *reinterpret_cast<float4*>(dst + rot[kc]) =
                        make_float4(dq[4 * kc], dq[4 * kc + 1], dq[4 * kc + 2], dq[4 * kc + 3]);
            }
        }
        __syncwarp();          // scratch complete; nobody reads K / V any more
        // ---- phase 2: lane = 4 features (head h2, chunk kc) of every entity row -----------------------------
        // dK_j = sum_{c,i} dlogit_ij Q_i, dV_j = sum_{c,i} w_ij dO_i: the lane keeps its slice of the Q_i / dO_i rows of
        // 8 agents in registers (coalesced global loads), sweeps the entities j with the scratch values broadcast from
        // shared memory, and accumulates into the K|V tile (dead as K|V by now), which becomes dK|dV.
        if (lane < (d >> 2)) {
            const int h2 = lane / NCH, kc2 = lane - h2 * NCH;
            bool first = true;
            for (int c = 0; c < a.C; c++) {
                for (int i0 = 0; i0 < nq; i0 += 8) {
                    float4 qs[8], gs[8];
#pragma unroll
                    for (int ii = 0; ii < 8; ii++) {
                        const bool ok = i0 + ii < nq;
                        qs[ii] = ok ? __ldg(reinterpret_cast<const float4*>(a.qkv + ((size_t)n * ne + i0 + ii) * 3 * d) + lane)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
                        gs[ii] = ok ? __ldg(reinterpret_cast<const float4*>(a.dout + (((size_t)c * a.N + n) * nq + i0 + ii) * d) + lane)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    const float* swb = sw + ((size_t)c * NEB * H + h2) * nqp + i0;
                    const float* sdb = sdl + ((size_t)c * NEB * H + h2) * nqp + i0;
#pragma unroll 2
                    for (int j = 0; j < NEB; j++) {
                        float4 dk = make_float4(0.f, 0.f, 0.f, 0.f), dv = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int ii = 0; ii < 8; ii++) {
                            const float dl = sdb[j * sstr + ii], wv = swb[j * sstr + ii];
                            dk.x = fmaf(dl, qs[ii].x, dk.x); dk.y = fmaf(dl, qs[ii].y, dk.y);
                            dk.z = fmaf(dl, qs[ii].z, dk.z); dk.w = fmaf(dl, qs[ii].w, dk.w);
                            dv.x = fmaf(wv, gs[ii].x, dv.x); dv.y = fmaf(wv, gs[ii].y, dv.y);
                            dv.z = fmaf(wv, gs[ii].z, dv.z); dv.w = fmaf(wv, gs[ii].w, dv.w);
                        }
                        float4* kd = reinterpret_cast<float4*>(kv + j * ldk + h2 * HD + 4 * kc2);
                        float4* vd = reinterpret_cast<float4*>(kv + j * ldk + d + h2 * HD + 4 * kc2);
                        if (!first) {
                            const float4 k0 = *kd, v0 = *vd;
                            dk.x += k0.x; dk.y += k0.y; dk.z += k0.z; dk.w += k0.w;
                            dv.x += v0.x; dv.y += v0.y; dv.z += v0.z; dv.w += v0.w;
                        }
                        *kd = dk;
                        *vd = dv;
                    }
                    first = false;
                }
            }
        }
        __syncwarp();
        // stream the dK | dV tile out as columns d..3d of dQKV (coalesced 128-bit stores)
        {
            const int d2 = (2 * d) >> 2;
            float* dst = a.dqkv + ((size_t)n * ne) * 3 * d + d;
            for (int f = lane; f < ne * d2; f += 32) {
                const int rr = f / d2, c4 = f - rr * d2;
                const float4 v = *reinterpret_cast<const float4*>(kv + rr * ldk + 4 * c4);
                *reinterpret_cast<float4*>(dst + (size_t)rr * 3 * d + 4 * c4) = v;
            }
        }
        __syncwarp();
    }
}

// =====================================================================================================================
// Backward, specialised for the benchmark's geometry (4 heads of 32, <= 8 query rows): TWO agents per lane.
// ncu on attn_bwd_kernel<32, 4> (profiles/r2n_ncu_full.md): DRAM 35 %, l1/smem 61 %, 8.9 % warps -- it is bound by the shared-memory
// pipe.  With lane = (agent, head) the eight agent-lanes of a head read the same 16 bytes of a K / V row: a 128-bit load delivers
// 64 distinct bytes per wavefront and every row chunk is fetched once per agent pass.  Here lane = (agent pair ip, head h, half dh
// of the head dim): the eight lanes of a quarter-warp read eight DISTINCT 16-byte chunks (all 32 banks, chunk order rotated by the
// head) and every chunk feeds two agents -- half the 128-bit loads of phase 1, each at full width.  The halves of a dot product meet
// with one shuffle; per-row scalars (max, sum, t) are computed by the lane that owns the row (dh = 0: agent ip, dh = 1: agent ip + 4).
// The mask copies share work that is linear in the dlogits: dQ_i = sum_j (sum_c dl^c_ij) K_j and dK_j = sum_i (sum_c dl^c_ij) Q_i, so the
// dlogits are summed over the copies in the scratch and dQ / dK take ONE pass for all copies (the generic kernel: one per copy); dV
// and dw = <dO, V> take the dO rows of all copies in registers and sweep V / write the tile once.  sum_c dl^c_j = (sum_c w^c_j dw^c_j -
// sum_c w^c_j t_c) / scale needs no per-copy dw scratch, and the logits live in the column that later holds the summed dlogits: C = 3
// costs 9 instead of 13 contraction passes per unit and 12 KB instead of 21 KB of scratch (6 warps per SM instead of 4).
// =====================================================================================================================
// dV_j = sum_c sum_i w^c_ij dO^c_i for the NC mask copies at once: the lane keeps its 4-feature slice of the dO rows of all copies
// in registers, so the V half of the tile is written exactly once (no read-modify-write per copy)
template <int NC>
__device__ __forceinline__ void att_h4_dv_pass(const AttnArgs& a, long long n, int lane, int nq, int NEB, float* kv, const float* sw) {
    constexpr int HD = 32, H = 4, d = HD * H, ldk = 2 * d, nqp = 8, sstr = H * nqp;
    const int h2 = lane >> 3;
    float4 gs[NC][8];
#pragma unroll
    for (int c = 0; c < NC; c++)
#pragma unroll
        for (int ii = 0; ii < 8; ii++)
            gs[c][ii] = ii < nq ? __ldg(reinterpret_cast<const float4*>(a.dout + (((size_t)c * a.N + n) * nq + ii) * d) + lane)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
    const float* swb = sw + h2 * nqp;
#pragma unroll 2
    for (int j = 0; j < NEB; j++) {
        float4 dv = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const float4 wa = *reinterpret_cast<const float4*>(swb + ((size_t)c * NEB + j) * sstr);
            const float4 wb = *reinterpret_cast<const float4*>(swb + ((size_t)c * NEB + j) * sstr + 4);
            const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int ii = 0; ii < 8; ii++) {
                dv.x = fmaf(wv[ii], gs[c][ii].x, dv.x); dv.y = fmaf(wv[ii], gs[c][ii].y, dv.y);
                dv.z = fmaf(wv[ii], gs[c][ii].z, dv.z); dv.w = fmaf(wv[ii], gs[c][ii].w, dv.w);
            }
        }
        *reinterpret_cast<float4*>(kv + j * ldk + d + 4 * lane) = dv;
    }
}

template <bool GROUPED>
__global__ void __launch_bounds__(32 * ATT_MAX_WARPS, 1) attn_bwd_h4_kernel(const __grid_constant__ AttnGroup grp, int NEB, int tile_floats, int warp_floats, const __grid_constant__ AttnMaps maps, int use_tmap) {
    int prob = 0, cta = (int)blockIdx.x, ncta = (int)gridDim.x;
    if (GROUPED) {
        while (prob + 1 < ATT_MAX_GROUP && (int)blockIdx.x >= grp.cta_begin[prob + 1]) prob++;
        cta = (int)blockIdx.x - grp.cta_begin[prob];
        ncta = grp.cta_begin[prob + 1] - grp.cta_begin[prob];
    }
    const AttnArgs& a = grp.a[prob];
    const CUtensorMap& tmap = maps.m[prob];
    extern __shared__ __align__(128) float smem_raw_[];
    float* smem = smem_raw_ + (((128u - (att_smem_u32(smem_raw_) & 127u)) & 127u) >> 2);
    constexpr int HD = 32, H = 4, d = HD * H, ldk = 2 * d, nqp = 8;
    const int ne = a.ne, nq = a.nq;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    const int dh = lane & 1, h = (lane >> 1) & 3, ip = lane >> 3;          // half of the head dim, head, agent pair (ip, ip + 4)
    const int i0 = ip, i1 = ip + 4, io = dh ? i1 : i0;                      // io: the row whose scalars this lane owns
    int rot[4];                                                             // float offset of my kc-th chunk inside a row
#pragma unroll
    for (int kc = 0; kc < 4; kc++) rot[kc] = h * HD + dh * 16 + 4 * ((kc + h) & 3);
    float* kv = smem + (size_t)warp * warp_floats;                       // [NEB][ldk]
    constexpr int sstr = H * nqp;                                        // scratch stride between entities j (= 32)
    float* sw = kv + tile_floats;                                        // [C][NEB][H][nqp]  softmax weights of every copy
    float* sdl = sw + a.C * NEB * sstr;                                  // [NEB][H][nqp]     logits, then sum_c w dw, then the dlogits
                                                                         //                   SUMMED over the copies
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)wpc * warp_floats) + warp;
    if (lane == 0) {
        att_mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    pdl_launch_dependents();
    pdl_wait();
    const float inv_scale = 1.f / sqrtf((float)HD);
    const long long gw = (long long)cta * wpc + warp, GW = (long long)ncta * wpc;
    uint32_t parity = 0;
    auto load_row = [&](const float* base, bool ok, float (&dst)[16]) {    // my 16 floats of a [d]-wide row, rotated chunk order
#pragma unroll
        for (int kc = 0; kc < 4; kc++) {
            const float4 v = ok ? __ldg(reinterpret_cast<const float4*>(base + rot[kc])) : make_float4(0.f, 0.f, 0.f, 0.f);
            dst[4 * kc] = v.x; dst[4 * kc + 1] = v.y; dst[4 * kc + 2] = v.z; dst[4 * kc + 3] = v.w;
        }
    };
    // two agents' dot products of their 16-float register rows with my 16 floats of the smem row at `row`; two accumulators
    // per agent keep the dependent FMA chains at 8
    auto dot2 = [&](const float* row, const float (&x0)[16], const float (&x1)[16], float& r0, float& r1) {
        float p0a = 0.f, p0b = 0.f, p1a = 0.f, p1b = 0.f;
#pragma unroll
        for (int kc = 0; kc < 4; kc++) {
            const float4 k4 = *reinterpret_cast<const float4*>(row + rot[kc]);
            p0a = fmaf(x0[4 * kc], k4.x, p0a); p0b = fmaf(x0[4 * kc + 1], k4.y, p0b);
            p0a = fmaf(x0[4 * kc + 2], k4.z, p0a); p0b = fmaf(x0[4 * kc + 3], k4.w, p0b);
            p1a = fmaf(x1[4 * kc], k4.x, p1a); p1b = fmaf(x1[4 * kc + 1], k4.y, p1b);
            p1a = fmaf(x1[4 * kc + 2], k4.z, p1a); p1b = fmaf(x1[4 * kc + 3], k4.w, p1b);
        }
        r0 = p0a + p0b;
        r1 = p1a + p1b;
    };
    for (long long n = gw; n < a.N; n += GW) {
        for (int f = ne * ldk + lane; f < NEB * ldk; f += 32) kv[f] = 0.f;   // padding rows (the tile is reused for dK|dV)
        if (use_tmap) att_tma_load_tile(kv, &tmap, d, ne, n, bar, lane, 2 * d);
        else att_tma_load_rows(kv, a.qkv + ((size_t)n * ne) * 3 * d + d, 3 * d, ne, 2 * d, ldk, bar, lane);
        AttMeta mc;
        att_meta_load(a, n, lane, 0, mc);
        const bool act0 = i0 < nq, act1 = i1 < nq;
        float q0[16], q1[16];
        load_row(a.qkv + ((size_t)n * ne + (act0 ? i0 : 0)) * 3 * d, act0, q0);
        load_row(a.qkv + ((size_t)n * ne + (act1 ? i1 : 0)) * 3 * d, act1, q1);
        // dQ of the non-query rows is zero
        {
            const int d4 = d >> 2;
            for (int f = lane; f < (ne - nq) * d4; f += 32) {
                const int r = nq + f / d4, c4 = f % d4;
                *reinterpret_cast<float4*>(a.dqkv + ((size_t)n * ne + r) * 3 * d + 4 * c4) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        uint32_t mb[ATT_MAX_COPIES];                                       // mask words of MY OWN row io
        att_meta_resolve(a, n, lane, 0, 8, io, mc, mb);
        att_mbar_wait(bar, parity);
        parity ^= 1;
        // ---- logits of both agents (shared by the copies); the owner keeps its row in the scratch column it owns ----------
        float* slr = sdl + h * nqp + io;                                 // [j * sstr]: my own row's column of the scratch
        float mx[ATT_MAX_COPIES] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll 4
        for (int j = 0; j < NEB; j++) {
            float p0, p1;
            dot2(kv + j * ldk, q0, q1, p0, p1);
            p0 += __shfl_xor_sync(0xffffffffu, p0, 1);
            p1 += __shfl_xor_sync(0xffffffffu, p1, 1);
            const float lg = (dh ? p1 : p0) * inv_scale;
            slr[j * sstr] = lg;
#pragma unroll
            for (int c = 0; c < ATT_MAX_COPIES; c++)
                if (!((mb[c] >> j) & 1u)) mx[c] = fmaxf(mx[c], lg);
        }
        // ---- softmax numerators of every copy (the logits die here: their column becomes the sum_c w dw accumulator) ------------
        float rs[ATT_MAX_COPIES] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < ATT_MAX_COPIES; c++) {
            if (c < a.C) {
                const uint32_t bits = mb[c];
                const float m = mx[c];
                float* swr = sw + ((size_t)c * NEB * H + h) * nqp + io;
                float ssum = 0.f;
#pragma unroll 4
                for (int j = 0; j < NEB; j++) {
                    const float e = ((bits >> j) & 1u) ? 0.f : __expf(slr[j * sstr] - m);
                    ssum += e;
                    swr[j * sstr] = e;
                }
                rs[c] = ssum > 0.f ? 1.f / ssum : 0.f;
            }
        }
        // ---- ONE pass over V for all copies: dw^c_ij = <dO^c_i, V_j>, t_c = sum_j w^c_j dw^c_j, A_j = sum_c w^c_j dw^c_j -------------
        // (dlogit^c_j = w^c_j (dw^c_j - t_c) / scale, so sum_c dlogit^c_j = (A_j - sum_c w^c_j t_c) / scale: no per-copy dw scratch)
        float ts[ATT_MAX_COPIES] = {0.f, 0.f, 0.f};
        auto v_pass = [&](auto nc_tag) {
            constexpr int NC = decltype(nc_tag)::value;
            float g[NC][2][16];                                        // dO rows of both agents for every copy (my half dims)
#pragma unroll
            for (int c = 0; c < NC; c++) {
                load_row(a.dout + (((size_t)c * a.N + n) * nq + (act0 ? i0 : 0)) * d, act0, g[c][0]);
                load_row(a.dout + (((size_t)c * a.N + n) * nq + (act1 ? i1 : 0)) * d, act1, g[c][1]);
            }
            float* swo = sw + h * nqp + io;
#pragma unroll 2
            for (int j = 0; j < NEB; j++) {
                float p[NC][2];
#pragma unroll
                for (int c = 0; c < NC; c++) { p[c][0] = 0.f; p[c][1] = 0.f; }
                float pb[NC][2];
#pragma unroll
                for (int c = 0; c < NC; c++) { pb[c][0] = 0.f; pb[c][1] = 0.f; }
#pragma unroll
                for (int kc = 0; kc < 4; kc++) {
                    const float4 v4 = *reinterpret_cast<const float4*>(kv + j * ldk + d + rot[kc]);
#pragma unroll
                    for (int c = 0; c < NC; c++)
#pragma unroll
                        for (int ag = 0; ag < 2; ag++) {
                            p[c][ag] = fmaf(g[c][ag][4 * kc], v4.x, p[c][ag]); pb[c][ag] = fmaf(g[c][ag][4 * kc + 1], v4.y, pb[c][ag]);
                            p[c][ag] = fmaf(g[c][ag][4 * kc + 2], v4.z, p[c][ag]); pb[c][ag] = fmaf(g[c][ag][4 * kc + 3], v4.w, pb[c][ag]);
                        }
                }
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    float d0 = p[c][0] + pb[c][0], d1 = p[c][1] + pb[c][1];
                    d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
                    d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
                    const float dw = dh ? d1 : d0;                     // <dO^c_io, V_j> of my own row
                    const float w = swo[((size_t)c * NEB + j) * sstr] * rs[c];
                    swo[((size_t)c * NEB + j) * sstr] = w;
                    const float wd = w * dw;
                    ts[c] += wd;
                    acc += wd;
                }
                slr[j * sstr] = acc;
            }
        };
        if (a.C == 1) v_pass(std::integral_constant<int, 1>{});
        else if (a.C == 2) v_pass(std::integral_constant<int, 2>{});
        else v_pass(std::integral_constant<int, 3>{});
        // ---- dQ: one pass over K with the copy-summed dlogits (kept in the scratch for the dK pass) -----------------------------
        float dq0[16], dq1[16];
#pragma unroll
        for (int k = 0; k < 16; k++) { dq0[k] = 0.f; dq1[k] = 0.f; }
        const float* swo = sw + h * nqp + io;
#pragma unroll 4
        for (int j = 0; j < NEB; j++) {
            float wt = 0.f;
#pragma unroll
            for (int c = 0; c < ATT_MAX_COPIES; c++)
                if (c < a.C) wt = fmaf(swo[((size_t)c * NEB + j) * sstr], ts[c], wt);
            const float dl = (slr[j * sstr] - wt) * inv_scale;                     // summed dlogit of my own row
            slr[j * sstr] = dl;
            const float dlo = __shfl_xor_sync(0xffffffffu, dl, 1);                 // ... and of my partner's row
            const float dl0 = dh ? dlo : dl, dl1 = dh ? dl : dlo;
#pragma unroll
            for (int kc = 0; kc < 4; kc++) {
                const float4 k4 = *reinterpret_cast<const float4*>(kv + j * ldk + rot[kc]);
                dq0[4 * kc] = fmaf(dl0, k4.x, dq0[4 * kc]); dq0[4 * kc + 1] = fmaf(dl0, k4.y, dq0[4 * kc + 1]);
                dq0[4 * kc + 2] = fmaf(dl0, k4.z, dq0[4 * kc + 2]); dq0[4 * kc + 3] = fmaf(dl0, k4.w, dq0[4 * kc + 3]);
                dq1[4 * kc] = fmaf(dl1, k4.x, dq1[4 * kc]); dq1[4 * kc + 1] = fmaf(dl1, k4.y, dq1[4 * kc + 1]);
                dq1[4 * kc + 2] = fmaf(dl1, k4.z, dq1[4 * kc + 2]); dq1[4 * kc + 3] = fmaf(dl1, k4.w, dq1[4 * kc + 3]);
            }
        }
        if (act0) {
            float* dst = a.dqkv + ((size_t)n * ne + i0) * 3 * d;
#pragma unroll
            for (int kc = 0; kc < 4; kc++)
                *reinterpret_cast<float4*>(dst + rot[kc]) = make_float4(dq0[4 * kc], dq0[4 * kc + 1], dq0[4 * kc + 2], dq0[4 * kc + 3]);
        }
        if (act1) {
            float* dst = a.dqkv + ((size_t)n * ne + i1) * 3 * d;
#pragma unroll
            for (int kc = 0; kc < 4; kc++)
                *reinterpret_cast<float4*>(dst + rot[kc]) = make_float4(dq1[4 * kc], dq1[4 * kc + 1], dq1[4 * kc + 2], dq1[4 * kc + 3]);
        }
        __syncwarp();          // scratch complete; nobody reads K / V any more
        // ---- phase 2: lane = features 4*lane .. 4*lane+3 (head lane / 8) of every entity row; the tile becomes dK | dV ----------
        {
            const int h2 = lane >> 3;
            float4 qs[8];
#pragma unroll
            for (int ii = 0; ii < 8; ii++)
                qs[ii] = ii < nq ? __ldg(reinterpret_cast<const float4*>(a.qkv + ((size_t)n * ne + ii) * 3 * d) + lane)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            const float* sdb = sdl + h2 * nqp;
#pragma unroll 2
            for (int j = 0; j < NEB; j++) {                              // dK_j = sum_i (sum_c dlogit^c_ij) Q_i
                float4 dk = make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 dla = *reinterpret_cast<const float4*>(sdb + j * sstr), dlb = *reinterpret_cast<const float4*>(sdb + j * sstr + 4);
                const float dlv[8] = {dla.x, dla.y, dla.z, dla.w, dlb.x, dlb.y, dlb.z, dlb.w};
#pragma unroll
                for (int ii = 0; ii < 8; ii++) {
                    dk.x = fmaf(dlv[ii], qs[ii].x, dk.x); dk.y = fmaf(dlv[ii], qs[ii].y, dk.y);
                    dk.z = fmaf(dlv[ii], qs[ii].z, dk.z); dk.w = fmaf(dlv[ii], qs[ii].w, dk.w);
                }
                *reinterpret_cast<float4*>(kv + j * ldk + 4 * lane) = dk;
            }
            if (a.C == 1) att_h4_dv_pass<1>(a, n, lane, nq, NEB, kv, sw);
            else if (a.C == 2) att_h4_dv_pass<2>(a, n, lane, nq, NEB, kv, sw);
            else att_h4_dv_pass<3>(a, n, lane, nq, NEB, kv, sw);
        }
        __syncwarp();
        // stream the dK | dV tile out as columns d..3d of dQKV (coalesced 128-bit stores)
        {
            const int d2 = (2 * d) >> 2;
            float* dst = a.dqkv + ((size_t)n * ne) * 3 * d + d;
            for (int f = lane; f < ne * d2; f += 32) {
                const int rr = f / d2, c4 = f - rr * d2;
                const float4 v = *reinterpret_cast<const float4*>(kv + rr * ldk + 4 * c4);
                *reinterpret_cast<float4*>(dst + (size_t)rr * 3 * d + 4 * c4) = v;
            }
        }
        __syncwarp();
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
    REFIL_CHECK_ARG(H <= 32 && (32 % H) == 0, "masked_attn: n_heads %d must divide 32", H);
    REFIL_CHECK_ARG(((uintptr_t)qkv % 16) == 0, "masked_attn: QKV must be 16-byte aligned");
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

// =====================================================================================================================
// EntityPoolingLayer (/root/reference/src/modules/layers/attention.py:82-132): the `pooling_type` ablation of the attention
// layer, with the SAME mask interface (explicit tensors + closed-form partition modes, up to 3 copies).
//   E [N, ne, d] = in_trans(x1) (dense kernels);   OUT[c, n, i, :] = pool_j ( masked_c(n, i, j) ? 0 : E[n, j, :] )
//   mean divides by ne, masked entities contribute zeros (:118-124); max takes the elementwise maximum INCLUDING those zeros
//   and its gradient goes to the first maximal entity (torch.max(dim) index rule) unless that entity is a masked zero.
// One warp per (b, t) unit; lane = 4 consecutive columns of a 128-column chunk (coalesced 512-byte row reads); the mask words
// of the unit's query rows are resolved once (lane = query row) and broadcast with shuffles.
// =====================================================================================================================
#define POOL_MEAN 0
#define POOL_MAX 1

__global__ void __launch_bounds__(128) entity_pool_fwd_kernel(AttnArgs a, const float* __restrict__ E, int pool_type) {
    const int lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    const long long gw = (long long)blockIdx.x * wpc + (threadIdx.x >> 5), GW = (long long)gridDim.x * wpc;
    const int ne = a.ne, nq = a.nq, d = a.d;
    const float inv_ne = 1.f / (float)ne;
    for (long long n = gw; n < a.N; n += GW) {
        AttMeta m;
        uint32_t mb[ATT_MAX_COPIES];
        att_meta_load(a, n, lane, 0, m);
        att_meta_resolve(a, n, lane, 0, 32, lane, m, mb);          // lane = query row: its mask word per copy
        const float* e0 = E + (size_t)n * ne * d;
        for (int col = lane * 4; col < d; col += 128) {
            for (int c = 0; c < a.C; c++) {
                for (int i = 0; i < nq; i++) {
                    const uint32_t word = __shfl_sync(0xffffffffu, mb[c], i);
                    float4 acc = pool_type == POOL_MAX ? make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY)
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int j = 0; j < ne; j++) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (!((word >> j) & 1u)) v = __ldg(reinterpret_cast<const float4*>(e0 + (size_t)j * d + col));
                        if (pool_type == POOL_MAX) {
                            acc.x = fmaxf(acc.x, v.x); acc.y = fmaxf(acc.y, v.y); acc.z = fmaxf(acc.z, v.z); acc.w = fmaxf(acc.w, v.w);
                        } else {
                            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                        }
                    }
                    if (pool_type == POOL_MEAN) { acc.x *= inv_ne; acc.y *= inv_ne; acc.z *= inv_ne; acc.w *= inv_ne; }
                    *reinterpret_cast<float4*>(a.out + (((size_t)c * a.N + n) * nq + i) * d + col) = acc;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(128) entity_pool_bwd_kernel(AttnArgs a, const float* __restrict__ E, float* __restrict__ dE,
                                                             int pool_type) {
    const int lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    const long long gw = (long long)blockIdx.x * wpc + (threadIdx.x >> 5), GW = (long long)gridDim.x * wpc;
    const int ne = a.ne, nq = a.nq, d = a.d;
    const float inv_ne = 1.f / (float)ne;
    for (long long n = gw; n < a.N; n += GW) {
        AttMeta m;
        uint32_t mb[ATT_MAX_COPIES];
        att_meta_load(a, n, lane, 0, m);
        att_meta_resolve(a, n, lane, 0, 32, lane, m, mb);
        const float* e0 = E + (size_t)n * ne * d;
        float* de0 = dE + (size_t)n * ne * d;
        for (int col = lane * 4; col < d; col += 128) {
            float4 acc[ATT_MAX_NE];                                 // gradient of my 4 columns of every entity row
#pragma unroll
            for (int j = 0; j < ATT_MAX_NE; j++) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int c = 0; c < a.C; c++) {
                for (int i = 0; i < nq; i++) {
                    const uint32_t word = __shfl_sync(0xffffffffu, mb[c], i);
                    const float4 g = __ldg(reinterpret_cast<const float4*>(a.dout + (((size_t)c * a.N + n) * nq + i) * d + col));
                    if (pool_type == POOL_MEAN) {
#pragma unroll
                        for (int j = 0; j < ATT_MAX_NE; j++)
                            if (j < ne && !((word >> j) & 1u)) {
                                acc[j].x += g.x * inv_ne; acc[j].y += g.y * inv_ne; acc[j].z += g.z * inv_ne; acc[j].w += g.w * inv_ne;
                            }
                    } else {
                        // recompute the first arg-max per column; -1 = the maximum is a masked zero (no gradient)
                        float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                        int jx = -1, jy = -1, jz = -1, jw = -1;
                        for (int j = 0; j < ne; j++) {
                            const bool mk = (word >> j) & 1u;
                            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (!mk) v = __ldg(reinterpret_cast<const float4*>(e0 + (size_t)j * d + col));
                            const int id = mk ? -1 : j;
                            if (v.x > best.x) { best.x = v.x; jx = id; }
                            if (v.y > best.y) { best.y = v.y; jy = id; }
                            if (v.z > best.z) { best.z = v.z; jz = id; }
                            if (v.w > best.w) { best.w = v.w; jw = id; }
                        }
#pragma unroll
                        for (int j = 0; j < ATT_MAX_NE; j++) {
                            if (jx == j) acc[j].x += g.x;
                            if (jy == j) acc[j].y += g.y;
                            if (jz == j) acc[j].z += g.z;
                            if (jw == j) acc[j].w += g.w;
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < ATT_MAX_NE; j++)
                if (j < ne) *reinterpret_cast<float4*>(de0 + (size_t)j * d + col) = acc[j];
        }
    }
}

static int pool_fill(AttnArgs& a, const char* name, const float* E, const uint8_t* m0, const uint8_t* m1, const uint8_t* m2,
                     long long s0, long long s1, long long s2, int mode0, int mode1, int mode2, const uint8_t* group_bits,
                     const uint8_t* entity_mask, int N, int T, int ne, int nq, int d, int C, int pool_type) {
    REFIL_CHECK_ARG(pool_type == POOL_MEAN || pool_type == POOL_MAX, "%s: pooling type %d (0 = mean, 1 = max)", name, pool_type);
    REFIL_CHECK_ARG(d >= 4 && d % 4 == 0, "%s: embed dim %d must be a multiple of 4", name, d);
    // same mask contract as the attention kernels (checked there with a formal head count of d / 8)
    return attn_fill(a, E, m0, m1, m2, s0, s1, s2, mode0, mode1, mode2, group_bits, entity_mask, N, T, ne, nq, 8, 1, C) ||
           ((a.d = d), REFIL_OK);
}

extern "C" int refil_entity_pool_fwd(const float* E, float* out, const uint8_t* mask0, const uint8_t* mask1,
                                     const uint8_t* mask2, long long mask_stride0, long long mask_stride1,
                                     long long mask_stride2, int mode0, int mode1, int mode2, const uint8_t* group_bits,
                                     const uint8_t* entity_mask, int N, int T, int n_entities, int n_queries, int embed_dim,
                                     int n_copies, int pool_type, cudaStream_t stream) {
    AttnArgs a{};
    int rc = pool_fill(a, "entity_pool_fwd", E, mask0, mask1, mask2, mask_stride0, mask_stride1, mask_stride2, mode0, mode1,
                       mode2, group_bits, entity_mask, N, T, n_entities, n_queries, embed_dim, n_copies, pool_type);
    if (rc) return rc;
    REFIL_CHECK_ARG(out && ((uintptr_t)out % 16) == 0, "entity_pool_fwd: out is null / unaligned");
    a.out = out;
    int grid = refil_cdiv(N, 4);
    if (grid > 8 * refil_num_sms()) grid = 8 * refil_num_sms();
    entity_pool_fwd_kernel<<<grid, 128, 0, stream>>>(a, E, pool_type);
    REFIL_CHECK_LAUNCH("entity_pool_fwd");
    return REFIL_OK;
}

extern "C" int refil_entity_pool_bwd(const float* E, const float* dout, float* dE, const uint8_t* mask0, const uint8_t* mask1,
                                     const uint8_t* mask2, long long mask_stride0, long long mask_stride1,
                                     long long mask_stride2, int mode0, int mode1, int mode2, const uint8_t* group_bits,
                                     const uint8_t* entity_mask, int N, int T, int n_entities, int n_queries, int embed_dim,
                                     int n_copies, int pool_type, cudaStream_t stream) {
    AttnArgs a{};
    int rc = pool_fill(a, "entity_pool_bwd", E, mask0, mask1, mask2, mask_stride0, mask_stride1, mask_stride2, mode0, mode1,
                       mode2, group_bits, entity_mask, N, T, n_entities, n_queries, embed_dim, n_copies, pool_type);
    if (rc) return rc;
    REFIL_CHECK_ARG(dout && dE && ((uintptr_t)dout % 16) == 0 && ((uintptr_t)dE % 16) == 0, "entity_pool_bwd: dout / dE");
    a.dout = dout;
    int grid = refil_cdiv(N, 4);
    if (grid > 8 * refil_num_sms()) grid = 8 * refil_num_sms();
    entity_pool_bwd_kernel<<<grid, 128, 0, stream>>>(a, E, dE, pool_type);
    REFIL_CHECK_LAUNCH("entity_pool_bwd");
    return REFIL_OK;
}

template <class K, class... Extra>
static int attn_launch(K kernel, const AttnGroup& a, int n_problems, size_t smem, int grid, int warps, cudaStream_t stream,
                       const char* name, Extra... extra) {
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
    (void)n_problems;
    cudaError_t le = refil_launch(kernel, dim3(grid), dim3(32 * warps), smem, stream, true, a, extra...);
    if (le != cudaSuccess) {
        refil_set_error("%s: launch failed: %s", name, cudaGetErrorString(le));
        return REFIL_ERR_CUDA;
    }
    REFIL_CHECK_LAUNCH(name);
    return REFIL_OK;
}

// dispatch on (head dim, heads): both are template parameters so that every tile offset is an immediate
#define ATT_DISPATCH(KERNEL, G, name, hd, heads, ...)                                                           \
    switch ((hd) * 100 + (heads)) {                                                                           \
        case 801: return attn_launch(KERNEL<8, 1, G>, __VA_ARGS__);                                              \
        case 802: return attn_launch(KERNEL<8, 2, G>, __VA_ARGS__);                                              \
        case 804: return attn_launch(KERNEL<8, 4, G>, __VA_ARGS__);                                              \
        case 808: return attn_launch(KERNEL<8, 8, G>, __VA_ARGS__);                                              \
        case 1601: return attn_launch(KERNEL<16, 1, G>, __VA_ARGS__);                                            \
        case 1602: return attn_launch(KERNEL<16, 2, G>, __VA_ARGS__);                                            \
        case 1604: return attn_launch(KERNEL<16, 4, G>, __VA_ARGS__);                                            \
        case 1608: return attn_launch(KERNEL<16, 8, G>, __VA_ARGS__);                                            \
        case 3201: return attn_launch(KERNEL<32, 1, G>, __VA_ARGS__);                                            \
        case 3202: return attn_launch(KERNEL<32, 2, G>, __VA_ARGS__);                                            \
        case 3204: return attn_launch(KERNEL<32, 4, G>, __VA_ARGS__);                                            \
        case 3208: return attn_launch(KERNEL<32, 8, G>, __VA_ARGS__);                                            \
        default:                                                                                              \
            refil_set_error("%s: (head dim %d, heads %d) not instantiated (heads in {1,2,4,8})", name, hd, heads); \
            return REFIL_ERR_UNSUPPORTED;                                                                     \
    }

// CUtensorMap of QKV viewed as [N][ne][3d] fp32 with a (2d, ne, 1) box: one TMA instruction stages a unit's K|V tile
typedef CUresult (*att_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static att_encode_fn attn_encoder() {
    static att_encode_fn enc = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            enc = (att_encode_fn)p;
    }
    return enc;
}

// QKV viewed as [N][ne][3d] fp32 with a (box_cols, box_rows, 1) box
static int attn_make_tmap_box(CUtensorMap* tm, const float* qkv, int N, int ne, int d, int box_cols, int box_rows) {
    memset(tm, 0, sizeof(*tm));
    if (box_cols > 256 || box_rows > 256) return 0;
    att_encode_fn enc = attn_encoder();
    if (!enc) return 0;
    const cuuint64_t dims[3] = {(cuuint64_t)3 * d, (cuuint64_t)ne, (cuuint64_t)N};
    const cuuint64_t strides[2] = {(cuuint64_t)3 * d * 4, (cuuint64_t)ne * 3 * d * 4};
    const cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)qkv, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 1 : 0;
}

// the whole K|V tile of a unit in one box (backward kernel); 0 = fall back to per-row bulk copies
static int attn_make_tmap(CUtensorMap* tm, const float* qkv, int N, int ne, int d) {
    return attn_make_tmap_box(tm, qkv, N, ne, d, 2 * d, ne);
}

static int attn_geometry(const char* name, int N, int warp_floats, int* warps, int* grid, size_t* smem, int ctas_per_sm = 1) {
    const size_t per_warp = (size_t)warp_floats * sizeof(float);
    int w = (int)((220 * 1024) / (per_warp + 8));
    if (w > ATT_MAX_WARPS) w = ATT_MAX_WARPS;
    if (w < 1) {
        refil_set_error("%s: one unit needs %zu bytes of shared memory (> 220 KB)", name, per_warp);
        return REFIL_ERR_UNSUPPORTED;
    }
    *warps = w;
    *smem = (size_t)w * per_warp + (size_t)w * 8 + 16 + 128;
    if (ctas_per_sm > 1 && (*smem + 1024) * ctas_per_sm > 227 * 1024) ctas_per_sm = 1;   // the tiles of two CTAs must fit
    int g = refil_num_sms() * ctas_per_sm;
    const int need = refil_cdiv(N, w);
    if (g > need) g = need;
    *grid = g;
    return REFIL_OK;
}

// fwd = true: forward (out), else backward (dout, dqkv)
static int attn_group_launch(bool fwd, const RefilAttnDesc* descs, int n_problems, int N, int T, int n_entities, int n_queries,
                             int embed_dim, int n_heads, cudaStream_t stream) {
    const char* name = fwd ? "masked_attn_fwd" : "masked_attn_bwd";
    REFIL_CHECK_ARG(descs && n_problems >= 1 && n_problems <= ATT_MAX_GROUP, "%s: 1..%d problems (got %d)", name, ATT_MAX_GROUP,
                    n_problems);
    AttnGroup grp{};
    AttnMaps maps{};
    int max_c = 1, use_tmap = 1;
    for (int g = 0; g < n_problems; g++) {
        const RefilAttnDesc& d = descs[g];
        AttnArgs& a = grp.a[g];
        int rc = attn_fill(a, d.qkv, d.mask0, d.mask1, d.mask2, d.mask_stride0, d.mask_stride1, d.mask_stride2, d.mode0, d.mode1,
                           d.mode2, d.group_bits, d.entity_mask, N, T, n_entities, n_queries, embed_dim, n_heads, d.n_copies);
        if (rc) return rc;
        if (fwd) {
            REFIL_CHECK_ARG(d.out != nullptr, "masked_attn_fwd: out is null");
            a.out = d.out;
            use_tmap &= attn_make_tmap_box(&maps.m[g], d.qkv, N, n_entities, embed_dim, embed_dim, n_entities);
        } else {
            REFIL_CHECK_ARG(d.dout && d.dqkv, "masked_attn_bwd: dout / dqkv is null");
            REFIL_CHECK_ARG(((uintptr_t)d.dout % 16) == 0 && ((uintptr_t)d.dqkv % 16) == 0, "masked_attn_bwd: alignment");
            a.dout = d.dout;
            a.dqkv = d.dqkv;
            use_tmap &= attn_make_tmap(&maps.m[g], d.qkv, N, n_entities, embed_dim);
        }
        if (d.n_copies > max_c) max_c = d.n_copies;
    }
    const int hd = embed_dim / n_heads, neb = (n_entities + 7) / 8 * 8;
    const int tile_floats = neb * 2 * embed_dim;              // K tile | V tile, [neb][d] each
    // the benchmark's geometry (4 heads of 32, <= 8 query rows) takes the two-agents-per-lane kernels; REFIL_ATTN=generic selects
    // the generic ones (A/B runs)
    static int attn_generic = -1;
    if (attn_generic < 0) {
        const char* e = getenv("REFIL_ATTN");
        attn_generic = (e && e[0] == 'g') ? 1 : 0;
    }
    const bool h4 = !attn_generic && hd == 32 && n_heads == 4 && n_queries <= 8;
    const bool bwd_h4 = !fwd && h4;
    int warp_floats;
    if (fwd) {
        warp_floats = tile_floats + neb * 32;                 // multiples of 32 floats: every warp tile is 128-byte aligned
    } else {
        const int ipp = 32 / n_heads, nqp = (n_queries + ipp - 1) / ipp * ipp;
        if (bwd_h4) warp_floats = tile_floats + (max_c + 1) * neb * 32;      // w per copy + ONE logit / dlogit column set
        else warp_floats = (tile_floats + 2 * max_c * neb * n_heads * nqp + neb * 32 + 31) / 32 * 32;
    }
    int warps, grid;
    size_t smem;
    int rc = attn_geometry(name, N, warp_floats, &warps, &grid, &smem, hd <= 16 ? 2 : 1);
    if (rc) return rc;
    if (fwd) smem += (size_t)warps * 8;                       // two mbarriers per warp (K tile, V tile)
    if (n_problems > 1) {
        // the group shares one wave of CTAs, dealt out in proportion to the mask copies (= work) of each problem
        const int total = refil_num_sms() * ((hd <= 16 && (smem + 1024) * 2 <= 227 * 1024) ? 2 : 1);
        int sum_c = 0;
        for (int g = 0; g < n_problems; g++) sum_c += descs[g].n_copies;
        int begin = 0;
        for (int g = 0; g < n_problems; g++) {
            grp.cta_begin[g] = begin;
            int share = (int)((long long)total * descs[g].n_copies / sum_c);
            const int need = refil_cdiv(N, warps);
            if (share < 1) share = 1;
            if (share > need) share = need;
            begin += share;
        }
        for (int g = n_problems; g <= ATT_MAX_GROUP; g++) grp.cta_begin[g] = begin;
        grid = begin;
        if (fwd) {
            if (h4) return attn_launch(attn_fwd_h4_kernel<true>, grp, n_problems, smem, grid, warps, stream, "masked_attn_fwd", neb, tile_floats, warp_floats, maps, use_tmap);
            ATT_DISPATCH(attn_fwd_kernel, true, "masked_attn_fwd", hd, n_heads, grp, n_problems, smem, grid, warps, stream, "masked_attn_fwd", neb, tile_floats, warp_floats, maps, use_tmap)
        }
        if (bwd_h4) return attn_launch(attn_bwd_h4_kernel<true>, grp, n_problems, smem, grid, warps, stream, "masked_attn_bwd", neb, tile_floats, warp_floats, maps, use_tmap);
        ATT_DISPATCH(attn_bwd_kernel, true, "masked_attn_bwd", hd, n_heads, grp, n_problems, smem, grid, warps, stream, "masked_attn_bwd", neb, tile_floats, warp_floats, maps, use_tmap)
    }
    if (fwd) {
        if (h4) return attn_launch(attn_fwd_h4_kernel<false>, grp, n_problems, smem, grid, warps, stream, "masked_attn_fwd", neb, tile_floats, warp_floats, maps, use_tmap);
        ATT_DISPATCH(attn_fwd_kernel, false, "masked_attn_fwd", hd, n_heads, grp, n_problems, smem, grid, warps, stream, "masked_attn_fwd", neb, tile_floats, warp_floats, maps, use_tmap)
    }
    if (bwd_h4) return attn_launch(attn_bwd_h4_kernel<false>, grp, n_problems, smem, grid, warps, stream, "masked_attn_bwd", neb, tile_floats, warp_floats, maps, use_tmap);
    ATT_DISPATCH(attn_bwd_kernel, false, "masked_attn_bwd", hd, n_heads, grp, n_problems, smem, grid, warps, stream, "masked_attn_bwd", neb, tile_floats, warp_floats, maps, use_tmap)
}

extern "C" int refil_masked_attn_fwd_group(const RefilAttnDesc* descs, int n_problems, int N, int T, int n_entities,
                                           int n_queries, int embed_dim, int n_heads, cudaStream_t stream) {
    return attn_group_launch(true, descs, n_problems, N, T, n_entities, n_queries, embed_dim, n_heads, stream);
}

extern "C" int refil_masked_attn_bwd_group(const RefilAttnDesc* descs, int n_problems, int N, int T, int n_entities,
                                           int n_queries, int embed_dim, int n_heads, cudaStream_t stream) {
    return attn_group_launch(false, descs, n_problems, N, T, n_entities, n_queries, embed_dim, n_heads, stream);
}

extern "C" int refil_masked_attn_fwd(const float* qkv, float* out, const uint8_t* mask0, const uint8_t* mask1,
                                     const uint8_t* mask2, long long mask_stride0, long long mask_stride1,
                                     long long mask_stride2, int mode0, int mode1, int mode2,
                                     const uint8_t* group_bits, const uint8_t* entity_mask, int N, int T,
                                     int n_entities, int n_queries, int embed_dim, int n_heads, int n_copies,
                                     cudaStream_t stream) {
    RefilAttnDesc d{qkv, out, nullptr, nullptr, mask0, mask1, mask2, mask_stride0, mask_stride1, mask_stride2, mode0, mode1, mode2,
                    group_bits, entity_mask, n_copies};
    return attn_group_launch(true, &d, 1, N, T, n_entities, n_queries, embed_dim, n_heads, stream);
}

extern "C" int refil_masked_attn_bwd(const float* qkv, const float* dout, float* dqkv, const uint8_t* mask0,
                                     const uint8_t* mask1, const uint8_t* mask2, long long mask_stride0,
                                     long long mask_stride1, long long mask_stride2, int mode0, int mode1, int mode2,
                                     const uint8_t* group_bits, const uint8_t* entity_mask, int N, int T,
                                     int n_entities, int n_queries, int embed_dim, int n_heads, int n_copies,
                                     cudaStream_t stream) {
    RefilAttnDesc d{qkv, nullptr, dout, dqkv, mask0, mask1, mask2, mask_stride0, mask_stride1, mask_stride2, mode0, mode1, mode2,
                    group_bits, entity_mask, n_copies};
    return attn_group_launch(false, &d, 1, N, T, n_entities, n_queries, embed_dim, n_heads, stream);
}
