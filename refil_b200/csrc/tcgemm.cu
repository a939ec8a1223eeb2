// Dense layers on the 5th-generation tensor cores with fp32-grade accuracy: 3xTF32 split GEMM on tcgen05.mma.
//
//   C[M, N] = epi( g(A)[M, K] * B[N, K]^T )        ("TN": both operands K-major)
//     forward  : A = activations, B = nn.Linear weight [out, in], epi = +bias, relu, inactive-agent row mask
//     bwd-data : A = upstream gradient with fused relu' / row mask (g), B = W^T given by strides (elements of B are
//                read as B[j*sbj + i*sbi], so no transposed copy of the weight is ever made)
//   (nn.Linear of /root/reference/src/modules/layers/attention.py:21-22, agents/entity_rnn_agent.py:12,23-25,
//    mixers/flex_qmix.py:29,39 and their autograd.)
//
// Why 3xTF32: the parity bar is 1e-4 relative on fp32 utilities / TD loss and bit-exact greedy indices; one TF32 (or
// bf16) pass has ~2^-11 relative input error and does not hold it.  Each fp32 operand x is
// split into hi = x & 0xffffe000 (exactly representable in TF32) and lo = x - hi (exact in fp32; <= 2^-11 |x|), and
//   A*B ~= A_lo*B_hi + A_hi*B_lo + A_hi*B_hi   accumulated in fp32 in tensor memory,
// which leaves ~2^-21 relative error per product -- fp32 grade.
//
// Two operand paths share the epilogue, the resident weight tile and the host entry point (REFIL_TC_MODE=ss|ts, default ts):
//   TS  tc_gemm_ts_kernel / tc_wgrad_ts_kernel (further down): the A operand lives in TENSOR MEMORY -- raw fp32 chunks arrive by
//       TMA, converter warps split them into TMEM columns, the MMA reads only the resident weight tile from shared memory.
//   SS  tc_gemm_tn_kernel / tc_gemm_wgrad_kernel: both operands in shared memory (the predecessor, kept for A/B runs).
//
// Structure of the SS kernel (one persistent CTA per SM, 17 warps, warp-specialised; the role index is made warp-uniform with
// a shuffle so that the MMA warp's descriptors live in uniform registers -- no per-instruction R2UR waterfall):
//   warps 0-7   epilogue: warp w owns TMEM lane quadrant w & 3 (rows) and every other 32-column slab (w >> 2):
//               tcgen05.ld 32 x 32 fp32 (lane = row) -> bias / relu / row mask -> 128B-swizzled 4 KB staging tile ->
//               ONE TMA tensor store (cp.async.bulk.tensor, or cp.reduce...add for split-K partials) per slab
//   warp  8     TMEM allocation + tcgen05.mma issue by one elected lane (kind::tf32, M=128, N=BN, K=8 per instruction)
//   warps 9-12, 13-16  two A-producer groups taking alternate K chunks, each with the NEXT chunk's global loads already
//               in flight in registers (4 chunks = 64 KB of loads in flight per SM): coalesced 128-bit loads ->
//               hi/lo split -> 128B-swizzled K-major tiles
//   The weight tile (BN x KS, hi and lo) is split ONCE per CTA and stays resident in shared memory: every CTA keeps one
//   (n-tile, k-slice) and walks the m-tiles.  A reduction too long for a resident tile (bwd-data of in_trans, K = 3d)
//   is cut into k-slices whose partial tiles are summed in L2 by the TMA reduce-add store into a zeroed C.
//   smem = B_hi/B_lo [KS/32][BN][32] + ring of S stages {A_hi, A_lo [128][32]} + 8 staging tiles; mbarrier full/empty
//   pairs; TMEM holds two accumulators so the epilogue of tile i overlaps the main loop of tile i+1.
#include <cuda.h>   // CUtensorMap (encode function fetched through cudaGetDriverEntryPoint; libcuda is not linked)

#include <stdlib.h>
#include <string.h>

#include "common.cuh"

#define TC_BM 128
#define TC_BK 32
#define TC_EPI_WARPS 8
#define TC_MMA_WARP 8
#define TC_PROD_WARP0 9
#define TC_THREADS (32 * 17)
#define TC_STG_BYTES 4096
#define TC_MAX_STAGES 4
#define TCW_THREADS 416

struct TcArgs {
    const float* A; long long lda;
    const float* relu_y; long long ldy;      // optional: A element is zeroed where relu_y <= 0
    const uint8_t* a_rowmask; int a_na, a_ne, a_mper;   // optional: A row zeroed (stack of [C, N, na] rows)
    const float* B; long long sbj, sbi;      // B[j, i] = B[j*sbj + i*sbi], j < N (output col), i < K (reduction)
    int kb_valid, b_vec;                     // B is zero for i >= kb_valid (zero-padded reduction); b_vec: 16-byte row loads allowed
    const float* bias; int relu;
    const uint8_t* c_rowmask; int c_na, c_ne, c_mper;
    int M, N, K;                             // K = whole reduction length
    int KS, k_slices;                        // reduction length of one k-slice (multiple of 32), K / KS
    int BN, n_tiles, m_tiles, stages;
    int row_group;                           // > 0: row m of A and C is physical row (m / row_group) * stride + m % row_group (rank-3 tensor maps)
    int accumulate;                          // C += A B^T (reduce-add stores, no clearing of C)
    uint32_t idesc;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar), done;
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
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row atoms 1024 bytes apart (cute::UMMA::SmemDescriptor);
// the start-address field (16-byte units) is the low 14 bits, so tiles are addressed by adding (bytes >> 4)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);        // start address, 16-byte units
    d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset: next 8-row group
    d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

__device__ __forceinline__ bool tc_row_masked(const uint8_t* em, int na, int ne, int mper, long long r) {
    if (!em) return false;
    int idx = (int)(r % mper);
    int n = idx / na, a = idx - n * na;
    return em[(size_t)n * ne + a] != 0;
}

// split 4 fp32 values and store them at the swizzled position of (row, 16-byte chunk) in the hi / lo tiles
__device__ __forceinline__ void tc_store_split(float* hi_tile, float* lo_tile, int row, int chunk, float4 v) {
    const int off = row * 32 + ((chunk ^ (row & 7)) << 2);
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
    *reinterpret_cast<float4*>(hi_tile + off) = h;
    *reinterpret_cast<float4*>(lo_tile + off) = l;
}

// TMA tensor stores of one 32 x 32 fp32 staging tile (128B-swizzled) at (column, row) of C; bulk-group completion
__device__ __forceinline__ void tc_tma_store(const CUtensorMap* tmap, int col, int row, uint32_t saddr) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(tmap), "r"(col), "r"(row), "r"(saddr) : "memory");
}
__device__ __forceinline__ void tc_tma_reduce_add(const CUtensorMap* tmap, int col, int row, uint32_t saddr) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(tmap), "r"(col), "r"(row), "r"(saddr) : "memory");
}
// the same for a row-grouped C (rank-3 map: column, row inside the group, group)
__device__ __forceinline__ void tc_tma_store3(const CUtensorMap* tmap, int col, int grp, uint32_t saddr) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(tmap), "r"(col), "r"(0), "r"(grp), "r"(saddr) : "memory");
}
__device__ __forceinline__ void tc_tma_reduce_add3(const CUtensorMap* tmap, int col, int grp, uint32_t saddr) {
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(tmap), "r"(col), "r"(0), "r"(grp), "r"(saddr) : "memory");
}
__device__ __forceinline__ void tc_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tc_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tc_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// the 8 float4 of one A chunk a producer thread owns: 16-byte chunk c of rows r0 + 16 i of m-tile mt, k-chunk kc
__device__ __forceinline__ void tc_load_a(const TcArgs& a, int mt, int kcol, int r0, int c, float4 (&va)[8]) {
    const long long rowb = (long long)mt * TC_BM + r0;
    const float* pa = a.A + rowb * a.lda + kcol + c * 4;
    const long long rstep = 16 * a.lda;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const long long row = rowb + 16 * i;
        va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < a.M && !tc_row_masked(a.a_rowmask, a.a_na, a.a_ne, a.a_mper, row))
            va[i] = __ldg(reinterpret_cast<const float4*>(pa + i * rstep));
    }
}
// relu' mask of the same elements: zero where the forward output was <= 0
__device__ __forceinline__ void tc_mask_relu(const TcArgs& a, int mt, int kcol, int r0, int c, float4 (&va)[8]) {
    const long long rowb = (long long)mt * TC_BM + r0;
    const float* py = a.relu_y + rowb * a.ldy + kcol + c * 4;
    const long long ystep = 16 * a.ldy;
    float4 vy[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        vy[i] = make_float4(1.f, 1.f, 1.f, 1.f);
        if (rowb + 16 * i < a.M) vy[i] = __ldg(reinterpret_cast<const float4*>(py + i * ystep));
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (!(vy[i].x > 0.f)) va[i].x = 0.f;
        if (!(vy[i].y > 0.f)) va[i].y = 0.f;
        if (!(vy[i].z > 0.f)) va[i].z = 0.f;
        if (!(vy[i].w > 0.f)) va[i].w = 0.f;
    }
}

// Resident weight tile [KS/32][hi | lo][BN][32]: every global load of the tile is issued before the first split / store
// (BN * KS <= 16384 elements -> at most 8 float4 per thread), so the prologue of a launch costs one memory latency, not eight
template <int NTHREADS>
__device__ __forceinline__ void tc_load_resident_b(const TcArgs& a, uint8_t* smem_b, uint32_t b_bytes, int nt, int k_off) {
    const int BN = a.BN, k4 = a.KS >> 2, tot = BN * k4;
    // element f of the tile -> (row r = output column j, 16-byte chunk kq along K).  Forward (sbi == 1, rows of W contiguous along
    // K): kq fastest, one float4 per thread.  Backward-data (B = W^T by strides, sbj == 1: contiguous along j): r fastest, so that
    // the four scalar loads of a warp are four coalesced 128-byte rows instead of 128 scattered sectors.
    const bool k_contig = a.sbi == 1;
    constexpr int MAXIT = 8;
    const int kv = a.kb_valid;
    const bool vec = k_contig && a.b_vec;
    float4 v[MAXIT];
#pragma unroll
    for (int it = 0; it < MAXIT; it++) {
        const int f = (int)threadIdx.x + it * NTHREADS;
        v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f < tot) {
            const int r = k_contig ? f / k4 : f % BN, kq = k_contig ? f - r * k4 : f / BN;
            const long long j = (long long)nt * BN + r;
            if (j < a.N) {
                const int k0 = k_off + kq * 4;
                const float* p = a.B + j * a.sbj + (long long)k0 * a.sbi;
                if (vec && k0 + 4 <= kv) v[it] = __ldg(reinterpret_cast<const float4*>(p));
                else {                       // strided (W^T), unaligned rows (lda = 53) or the zero-padded tail of the reduction
                    if (k0 < kv) v[it].x = __ldg(p);
                    if (k0 + 1 < kv) v[it].y = __ldg(p + a.sbi);
                    if (k0 + 2 < kv) v[it].z = __ldg(p + 2 * a.sbi);
                    if (k0 + 3 < kv) v[it].w = __ldg(p + 3 * a.sbi);
                }
            }
        }
    }
#pragma unroll
    for (int it = 0; it < MAXIT; it++) {
        const int f = (int)threadIdx.x + it * NTHREADS;
        if (f < tot) {
            const int r = k_contig ? f / k4 : f % BN, kq = k_contig ? f - r * k4 : f / BN;
            const int kc = kq >> 3, c = kq & 7;
            float* hi = reinterpret_cast<float*>(smem_b + (size_t)kc * 2 * b_bytes);
            tc_store_split(hi, hi + BN * 32, r, c, v[it]);
        }
    }
    for (int f = (int)threadIdx.x + MAXIT * NTHREADS; f < tot; f += NTHREADS) {      // never taken within the smem budget
        const int r = f / k4, kq = f - r * k4;
        const int kc = kq >> 3, c = kq & 7;
        const long long j = (long long)nt * BN + r;
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < a.N) {
            const int k0 = k_off + kq * 4;
            const float* p = a.B + j * a.sbj + (long long)k0 * a.sbi;
            if (k0 < kv) w.x = __ldg(p);
            if (k0 + 1 < kv) w.y = __ldg(p + a.sbi);
            if (k0 + 2 < kv) w.z = __ldg(p + 2 * a.sbi);
            if (k0 + 3 < kv) w.w = __ldg(p + 3 * a.sbi);
        }
        float* hi = reinterpret_cast<float*>(smem_b + (size_t)kc * 2 * b_bytes);
        tc_store_split(hi, hi + BN * 32, r, c, w);
    }
}

// Epilogue role (warps 0-7; warp w owns TMEM lane quadrant w & 3 and every other 32-column slab, w >> 2):
// TMEM (lane = row) -> registers -> bias / relu / row mask -> swizzled staging tile -> one TMA store per slab
__device__ __forceinline__ void tc_epilogue(const TcArgs& a, const CUtensorMap* tmap_c, int warp, int lane, int nt, int mt0,
                                            int mt_step, int my_tiles, uint32_t tmem_base, uint8_t* smem_stg,
                                            const float* epi_bias, uint64_t* tmem_full, uint64_t* tmem_empty) {
    const int BN = a.BN;
    {
        const int quad = warp & 3, half = warp >> 2;
        float* stg = reinterpret_cast<float*>(smem_stg + (size_t)warp * TC_STG_BYTES);
        const uint32_t stg_addr = smem_u32(stg);
        float* my_stg = stg + lane * 32;
        const int sw = lane & 7;
        for (int it = 0; it < my_tiles; it++) {
            const int mt = mt0 + it * mt_step;
            const int buf = it & 1;
            mbar_wait(&tmem_full[buf], (it >> 1) & 1);
            tc_fence_after();
            const int row0 = mt * TC_BM + quad * 32;
            const long long myrow = (long long)row0 + lane;
            const bool masked = myrow < a.M && tc_row_masked(a.c_rowmask, a.c_na, a.c_ne, a.c_mper, myrow);
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BN);
            for (int c0 = half * 32; c0 < BN; c0 += 64) {
                uint32_t r[32];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                    "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr + (uint32_t)c0));
                if (lane == 0) tc_bulk_wait_read();        // the previous store of this warp has read the staging tile
                __syncwarp();
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const float4 b4 = *reinterpret_cast<const float4*>(epi_bias + c0 + 4 * q);
                    float4 v;
                    v.x = __uint_as_float(r[4 * q + 0]) + b4.x;
                    v.y = __uint_as_float(r[4 * q + 1]) + b4.y;
                    v.z = __uint_as_float(r[4 * q + 2]) + b4.z;
                    v.w = __uint_as_float(r[4 * q + 3]) + b4.w;
                    if (a.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    if (masked) v = make_float4(0.f, 0.f, 0.f, 0.f);
                    *reinterpret_cast<float4*>(my_stg + ((q ^ sw) << 2)) = v;
                }
                fence_async_smem();                        // staging writes -> visible to the TMA (async proxy)
                __syncwarp();
                if (lane == 0 && row0 < a.M) {
                    const bool add = a.k_slices > 1 || a.accumulate;
                    if (a.row_group > 0) {
                        if (add) tc_tma_reduce_add3(tmap_c, nt * BN + c0, row0 / a.row_group, stg_addr);
                        else tc_tma_store3(tmap_c, nt * BN + c0, row0 / a.row_group, stg_addr);
                    } else if (add) tc_tma_reduce_add(tmap_c, nt * BN + c0, row0, stg_addr);
                    else tc_tma_store(tmap_c, nt * BN + c0, row0, stg_addr);
                    tc_bulk_commit();
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[buf]);
        }
        if (lane == 0) tc_bulk_wait_all();
    }
}

// split the 8 float4 of chunk cc into the hi / lo tiles of its pipeline stage and hand the stage to the MMA warp
__device__ __forceinline__ void tc_store_chunk(const float4 (&va)[8], uint8_t* smem_a, uint32_t stage_bytes, int soff, int cc,
                                               int S, uint64_t* full_bar, uint64_t* empty_bar) {
    const int stage = cc % S;
    const uint32_t phase = (uint32_t)(cc / S) & 1u;
    mbar_wait(&empty_bar[stage], phase ^ 1);
    float* ahi = reinterpret_cast<float*>(smem_a + (size_t)stage * stage_bytes) + soff;
    float* alo = ahi + TC_BM * 32;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(va[i].x) & 0xffffe000u);
        h.y = __uint_as_float(__float_as_uint(va[i].y) & 0xffffe000u);
        h.z = __uint_as_float(__float_as_uint(va[i].z) & 0xffffe000u);
        h.w = __uint_as_float(__float_as_uint(va[i].w) & 0xffffe000u);
        l.x = va[i].x - h.x; l.y = va[i].y - h.y; l.z = va[i].z - h.z; l.w = va[i].w - h.w;
        *reinterpret_cast<float4*>(ahi + i * 512) = h;
        *reinterpret_cast<float4*>(alo + i * 512) = l;
    }
    fence_async_smem();           // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    mbar_arrive(&full_bar[stage]);
}

__global__ void __launch_bounds__(TC_THREADS, 1) tc_gemm_tn_kernel(TcArgs a, const __grid_constant__ CUtensorMap tmap_c) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte aligned base (dynamic smem is only guaranteed 16-byte aligned)
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int BN = a.BN;
    const int S = a.stages, KC = a.KS / TC_BK;
    const uint32_t a_bytes = TC_BM * 128, b_bytes = (uint32_t)BN * 128;
    const uint32_t stage_bytes = 2 * a_bytes;
    uint8_t* smem_b = smem;                                   // [KC][hi | lo][BN][128 B]
    uint8_t* smem_a = smem + (size_t)KC * 2 * b_bytes;        // [S][hi | lo][128][128 B]
    uint8_t* smem_stg = smem_a + (size_t)S * stage_bytes;     // [8 epilogue warps][32 rows][128 B], 128B-swizzled
    __shared__ uint64_t full_bar[TC_MAX_STAGES], empty_bar[TC_MAX_STAGES], tmem_full[2], tmem_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float epi_bias[256];

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform role index
    const int lane = threadIdx.x & 31;
    // this CTA's (n-tile, k-slice) is fixed; its m-tiles are strided over the CTAs of the same group
    const int groups = a.n_tiles * a.k_slices;
    const int gid = blockIdx.x % groups;
    const int nt = gid % a.n_tiles, k_off = (gid / a.n_tiles) * a.KS;
    const int mt0 = blockIdx.x / groups, mt_step = gridDim.x / groups;
    const int my_tiles = mt0 < a.m_tiles ? (a.m_tiles - mt0 + mt_step - 1) / mt_step : 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) { mbar_init(&full_bar[s], 128); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 32 * TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    // resident weight tile: split once per CTA
    {
        tc_load_resident_b<TC_THREADS>(a, smem_b, b_bytes, nt, k_off);
        for (int c = threadIdx.x; c < BN; c += TC_THREADS) epi_bias[c] = a.bias ? __ldg(a.bias + (long long)nt * BN + c) : 0.f;
        fence_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp < TC_EPI_WARPS) {
        // ===================== epilogue =====================
        tc_epilogue(a, &tmap_c, warp, lane, nt, mt0, mt_step, my_tiles, tmem_base, smem_stg, epi_bias, tmem_full, tmem_empty);
    } else if (warp == TC_MMA_WARP) {
        // ===================== MMA issuer =====================
        // every operand below is warp-uniform (uniform registers); one elected lane issues
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint64_t desc_a0 = tc_smem_desc(smem_u32(smem_a)), desc_b0 = tc_smem_desc(smem_u32(smem_b));
        const uint64_t a_lo_off = (uint64_t)(a_bytes >> 4), b_lo_off = (uint64_t)(b_bytes >> 4);
        const uint32_t idesc = a.idesc;
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0; it < my_tiles; it++) {
            const int buf = it & 1;
            mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_u + (uint32_t)(buf * BN);
            for (int kc = 0; kc < KC; kc++) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint64_t a_hi = desc_a0 + (uint64_t)((uint32_t)stage * (stage_bytes >> 4));
                const uint64_t b_hi = desc_b0 + (uint64_t)((uint32_t)kc * ((2 * b_bytes) >> 4));
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < TC_BK / 8; ks++) {
                        const uint64_t adv = (uint64_t)(ks * 2);       // +32 bytes inside the 128B swizzle row
                        tc_mma(tmem_d, a_hi + a_lo_off + adv, b_hi + adv, idesc, (kc | ks) != 0);
                        tc_mma(tmem_d, a_hi + adv, b_hi + b_lo_off + adv, idesc, 1);
                        tc_mma(tmem_d, a_hi + adv, b_hi + adv, idesc, 1);
                    }
                    tc_commit(&empty_bar[stage]);                  // frees the smem stage when these MMAs retire
                    if (kc == KC - 1) tc_commit(&tmem_full[buf]);  // accumulator complete -> epilogue
                }
                __syncwarp();
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================== A producers: group 0 = warps 9-12 (even chunks), group 1 = warps 13-16 (odd chunks) ====
        // thread -> fixed 16-byte chunk c of rows r0 + 16 i (i < 8); chunk cc of the CTA = (m-tile cc / KC, k-chunk cc % KC)
        const int grp = (warp - TC_PROD_WARP0) >> 2;
        const int pt = (int)threadIdx.x - (TC_PROD_WARP0 * 32 + grp * 128);     // 0..127 within the group
        const int r0 = pt >> 3, c = pt & 7;
        const int soff = r0 * 32 + ((c ^ (r0 & 7)) << 2);   // float offset of (r0, c); row r0 + 16 i adds 512 i
        const int total = my_tiles * KC;
        const bool has_y = a.relu_y != nullptr;
        // ping-pong register buffers: while chunk cc is split and stored, the loads of this group's next chunk (cc + 2)
        // are already in flight -> 4 chunks (64 KB) of global loads in flight per SM
        float4 va[8], vb[8];
        int cc = grp;
        if (has_y) {
            for (; cc < total; cc += 2) {
                const int mt = mt0 + (cc / KC) * mt_step, kcol = k_off + (cc % KC) * TC_BK;
                tc_load_a(a, mt, kcol, r0, c, va);
                tc_mask_relu(a, mt, kcol, r0, c, va);
                tc_store_chunk(va, smem_a, stage_bytes, soff, cc, S, full_bar, empty_bar);
            }
        } else if (cc < total) {
            tc_load_a(a, mt0 + (cc / KC) * mt_step, k_off + (cc % KC) * TC_BK, r0, c, va);
            for (;;) {
                int cn = cc + 2;
                if (cn < total) tc_load_a(a, mt0 + (cn / KC) * mt_step, k_off + (cn % KC) * TC_BK, r0, c, vb);
                tc_store_chunk(va, smem_a, stage_bytes, soff, cc, S, full_bar, empty_bar);
                cc = cn;
                if (cc >= total) break;
                cn = cc + 2;
                if (cn < total) tc_load_a(a, mt0 + (cn / KC) * mt_step, k_off + (cn % KC) * TC_BK, r0, c, va);
                tc_store_chunk(vb, smem_a, stage_bytes, soff, cc, S, full_bar, empty_bar);
                cc = cn;
                if (cc >= total) break;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    }
}

// =====================================================================================================================
// TS variant: the A operand of tcgen05.mma lives in TENSOR MEMORY (lane = row, column = k), only B is read from smem.
// In the SS kernel above the tensor pipe and the shared-memory port are both saturated by the operand reads (A 4 KB + B 4 KB
// per 64-cycle MMA = 128 B/clk) while the producers also have to write the split tiles there.  Here
//   warp 9      one lane streams RAW fp32 A chunks [128 x 32] into a 4-deep smem ring with TMA tensor loads (128B swizzle,
//               out-of-range rows zero-filled by the hardware) -- no registers are held while a load is in flight;
//   warps 10-17 converters, thread = (row of the tile, 16-column half): 4 conflict-free LDS.128 of its swizzled row -> row
//               mask / relu' -> hi/lo split -> two tcgen05.st (16 columns each) into a 4-deep ring of TMEM columns;
//   warp 8      issues tcgen05.mma kind::tf32 with [tmem] A and the resident smem weight tile as B;
//   warps 0-7   the same TMA-store epilogue.
// Shared-memory traffic per 128 x 32 chunk drops from 144 KB to 80 KB, the A stages leave shared memory entirely
// (TMEM: 2 x BN accumulator columns + 4 x 64 operand columns), and 64 KB of loads are in flight per SM at no register cost.
// =====================================================================================================================
#define TS_MMA_WARP 8
#define TS_TMA_WARP 9
#define TS_CONV_WARP0 10
#define TS_CONV_WARPS 8
#define TS_THREADS (32 * (TS_CONV_WARP0 + TS_CONV_WARPS))
#define TS_RAW_STAGES 4
#define TS_A_STAGES 4
#define TS_A_COL0 256                     // operand ring: TMEM columns [256, 512), 64 per stage (hi | lo)

__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

__device__ __forceinline__ void tc_tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

// A launch carries up to TC_MAX_GROUP independent problems of one (N, K) geometry (blockIdx.y = problem): the same layer of
// several networks -- the eight hypernetworks of the online and target mixers -- runs as ONE launch whose CTAs each walk eight
// times as many m-tiles, instead of eight launches that each pay the prologue (weight-tile split, pipeline fill) and the tail.
#define TC_MAX_GROUP 8
struct TcGroup { TcArgs a[TC_MAX_GROUP]; int cta_begin[TC_MAX_GROUP + 1]; };   // grouped: CTAs [cta_begin[p], cta_begin[p+1]) -> problem p
struct TcMaps { CUtensorMap m[TC_MAX_GROUP]; };

// GROUPED = false: the single-problem instantiation reads its arguments at fixed parameter offsets (uniform registers, as before
// grouping existed); measured: indexing the parameter block with blockIdx.y costs the single launches 3-17 %.
template <bool GROUPED>
__global__ void __launch_bounds__(TS_THREADS, 1) tc_gemm_ts_kernel(const __grid_constant__ TcGroup grp,
                                                                  const __grid_constant__ TcMaps maps_c,
                                                                  const __grid_constant__ TcMaps maps_a) {
    // grouped launch: the CTAs are dealt out to the problems in proportion to their m-tiles
    int prob = 0, bid = (int)blockIdx.x, nbid = (int)gridDim.x;
    if (GROUPED) {
        while (prob + 1 < TC_MAX_GROUP && (int)blockIdx.x >= grp.cta_begin[prob + 1]) prob++;
        bid = (int)blockIdx.x - grp.cta_begin[prob];
        nbid = grp.cta_begin[prob + 1] - grp.cta_begin[prob];
    }
    const TcArgs& a = grp.a[prob];
    const CUtensorMap& tmap_c = maps_c.m[prob];
    const CUtensorMap& tmap_a = maps_a.m[prob];
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int BN = a.BN;
    const int KC = a.KS / TC_BK;
    const uint32_t raw_bytes = TC_BM * 128, b_bytes = (uint32_t)BN * 128;
    uint8_t* smem_b = smem;                                             // [KC][hi | lo][BN][128 B]
    uint8_t* smem_r = smem + (size_t)KC * 2 * b_bytes;                  // [4][128][128 B] raw fp32 A chunks (TMA, swizzled)
    uint8_t* smem_stg = smem_r + (size_t)TS_RAW_STAGES * raw_bytes;     // [8 epilogue warps][32 rows][128 B]
    __shared__ uint64_t raw_full[TS_RAW_STAGES], raw_empty[TS_RAW_STAGES], a_full[TS_A_STAGES], a_empty[TS_A_STAGES];
    __shared__ uint64_t tmem_full[2], tmem_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float epi_bias[256];

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform role index
    const int lane = threadIdx.x & 31;
    const int groups = a.n_tiles * a.k_slices;
    const int gid = bid % groups;
    const int nt = gid % a.n_tiles, k_off = (gid / a.n_tiles) * a.KS;
    const int mt0 = bid / groups, mt_step = nbid / groups;
    const int my_tiles = mt0 < a.m_tiles ? (a.m_tiles - mt0 + mt_step - 1) / mt_step : 0;
    const int total = my_tiles * KC;
    pdl_launch_dependents();       // the next kernel of the stream may be scheduled: its prologue overlaps this kernel's tail
    // The TMA lane initialises the barriers and puts the first raw A chunks in flight at once: they do not depend on the weight
    // tile, so their latency overlaps the prologue below instead of following it
    auto tma_chunk = [&](int st, int mt, int kc) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&raw_full[st])), "r"(raw_bytes)
                     : "memory");
        if (a.row_group > 0)          // row-grouped A: (column, row inside the group, group) -- the box lands as the same 128 x 32 chunk
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                ::"r"(smem_u32(smem_r + (size_t)st * raw_bytes)), "l"(&tmap_a), "r"(k_off + kc * TC_BK), "r"(0),
                  "r"(mt * TC_BM / a.row_group), "r"(smem_u32(&raw_full[st]))
                : "memory");
        else
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                ::"r"(smem_u32(smem_r + (size_t)st * raw_bytes)), "l"(&tmap_a), "r"(k_off + kc * TC_BK), "r"(mt * TC_BM),
                  "r"(smem_u32(&raw_full[st]))
                : "memory");
    };
    const int pre = min(total, TS_RAW_STAGES);                 // chunks requested before the prologue
    if (threadIdx.x == TS_TMA_WARP * 32) {
        for (int s = 0; s < TS_RAW_STAGES; s++) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 32 * TS_CONV_WARPS); }
        for (int s = 0; s < TS_A_STAGES; s++) { mbar_init(&a_full[s], 32 * TS_CONV_WARPS); mbar_init(&a_empty[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 32 * TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        pdl_wait();                                            // the A operand is the previous kernel's output
        int kc = 0, mt = mt0;
        for (int cc = 0; cc < pre; cc++) {
            tma_chunk(cc, mt, kc);
            if (++kc == KC) { kc = 0; mt += mt_step; }
        }
    }
    if (warp == TS_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    // resident weight tile: split once per CTA
    {
        tc_load_resident_b<TS_THREADS>(a, smem_b, b_bytes, nt, k_off);
        for (int c = threadIdx.x; c < BN; c += TS_THREADS) epi_bias[c] = a.bias ? __ldg(a.bias + (long long)nt * BN + c) : 0.f;
        fence_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();                    // weights / bias above are never written inside a step; everything below may depend on it
    const uint32_t tmem_base = tmem_base_s;

    if (warp < TC_EPI_WARPS) {
        tc_epilogue(a, &tmap_c, warp, lane, nt, mt0, mt_step, my_tiles, tmem_base, smem_stg, epi_bias, tmem_full, tmem_empty);
    } else if (warp == TS_MMA_WARP) {
        // ===================== MMA issuer: A from tensor memory, B = resident smem weight tile =====================
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint64_t desc_b0 = tc_smem_desc(smem_u32(smem_b));
        const uint64_t b_lo_off = (uint64_t)(b_bytes >> 4);
        const uint32_t idesc = a.idesc;
        int st = 0;
        uint32_t ph = 0;
        for (int it = 0; it < my_tiles; it++) {
            const int buf = it & 1;
            mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_u + (uint32_t)(buf * BN);
            for (int kc = 0; kc < KC; kc++) {
                mbar_wait(&a_full[st], ph);
                tc_fence_after();
                const uint32_t a_hi = tmem_u + (uint32_t)(TS_A_COL0 + st * 64), a_lo = a_hi + 32u;
                const uint64_t b_hi = desc_b0 + (uint64_t)((uint32_t)kc * ((2 * b_bytes) >> 4));
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < TC_BK / 8; ks++) {
                        const uint64_t adv = (uint64_t)(ks * 2);       // +32 bytes inside the 128B swizzle row of B
                        const uint32_t ac = (uint32_t)(ks * 8);        // +8 columns of the TMEM operand
                        tc_mma_ts(tmem_d, a_lo + ac, b_hi + adv, idesc, (kc | ks) != 0);
                        tc_mma_ts(tmem_d, a_hi + ac, b_hi + b_lo_off + adv, idesc, 1);
                        tc_mma_ts(tmem_d, a_hi + ac, b_hi + adv, idesc, 1);
                    }
                    tc_commit(&a_empty[st]);                       // frees the TMEM operand stage when these MMAs retire
                    if (kc == KC - 1) tc_commit(&tmem_full[buf]);
                }
                __syncwarp();
                if (++st == TS_A_STAGES) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == TS_TMA_WARP) {
        // ===================== TMA producer: raw fp32 chunks, 4 in flight =====================
        if (lane == 0) {
            int st = pre % TS_RAW_STAGES, kc = pre % KC, mt = mt0 + (pre / KC) * mt_step;
            uint32_t ph = (uint32_t)(pre / TS_RAW_STAGES) & 1u;
            for (int cc = pre; cc < total; cc++) {
                mbar_wait(&raw_empty[st], ph ^ 1);
                tma_chunk(st, mt, kc);
                if (++kc == KC) { kc = 0; mt += mt_step; }
                if (++st == TS_RAW_STAGES) { st = 0; ph ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // ===================== converters: thread = (row, 16-column half); raw smem -> mask -> hi / lo -> tensor memory ====
        const int quad = warp & 3;                               // TMEM lane quadrant this warp may access
        const int half = (warp - TS_CONV_WARP0) >> 2;            // columns [16 half, 16 half + 16) of the 32-wide chunk
        const int r = quad * 32 + lane;                          // my row of the 128-row tile
        const uint32_t t_lane = (uint32_t)(quad * 32) << 16;
        const bool has_y = a.relu_y != nullptr;
        const int roff = r * 32, rsw = r & 7;
        int sr = 0, sa = 0, kc = 0, mt = mt0;
        uint32_t phr = 0, pha = 0;
        for (int cc = 0; cc < total; cc++) {
            const long long row = (long long)mt * TC_BM + r;
            const bool keep = row < a.M && !tc_row_masked(a.a_rowmask, a.a_na, a.a_ne, a.a_mper, row);
            float4 y[4];
#pragma unroll
            for (int c = 0; c < 4; c++) y[c] = make_float4(1.f, 1.f, 1.f, 1.f);
            if (has_y && keep) {                                 // relu' mask: my 64-byte segment of the forward output row
                const float* py = a.relu_y + row * a.ldy + k_off + kc * TC_BK + half * 16;
#pragma unroll
                for (int c = 0; c < 4; c++) y[c] = __ldg(reinterpret_cast<const float4*>(py + 4 * c));
            }
            mbar_wait(&raw_full[sr], phr);
            const float* rawrow = reinterpret_cast<const float*>(smem_r + (size_t)sr * raw_bytes) + roff;
            float4 v[4];
#pragma unroll
            for (int c = 0; c < 4; c++) v[c] = *reinterpret_cast<const float4*>(rawrow + (((half * 4 + c) ^ rsw) << 2));
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                float4 x = v[c];
                x.x = (keep && y[c].x > 0.f) ? x.x : 0.f;
                x.y = (keep && y[c].y > 0.f) ? x.y : 0.f;
                x.z = (keep && y[c].z > 0.f) ? x.z : 0.f;
                x.w = (keep && y[c].w > 0.f) ? x.w : 0.f;
                const uint32_t hx = __float_as_uint(x.x) & 0xffffe000u, hy = __float_as_uint(x.y) & 0xffffe000u,
                               hz = __float_as_uint(x.z) & 0xffffe000u, hw = __float_as_uint(x.w) & 0xffffe000u;
                hi[4 * c] = hx; hi[4 * c + 1] = hy; hi[4 * c + 2] = hz; hi[4 * c + 3] = hw;
                lo[4 * c] = __float_as_uint(x.x - __uint_as_float(hx));
                lo[4 * c + 1] = __float_as_uint(x.y - __uint_as_float(hy));
                lo[4 * c + 2] = __float_as_uint(x.z - __uint_as_float(hz));
                lo[4 * c + 3] = __float_as_uint(x.w - __uint_as_float(hw));
            }
            mbar_arrive(&raw_empty[sr]);                         // my part of the raw chunk is in registers
            if (++sr == TS_RAW_STAGES) { sr = 0; phr ^= 1; }
            mbar_wait(&a_empty[sa], pha ^ 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + t_lane + (uint32_t)(TS_A_COL0 + sa * 64 + half * 16);
            tc_tmem_st16(taddr, hi);
            tc_tmem_st16(taddr + 32u, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            mbar_arrive(&a_full[sa]);
            if (++sa == TS_A_STAGES) { sa = 0; pha ^= 1; }
            if (++kc == KC) { kc = 0; mt += mt_step; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TS_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    }
}

// CUtensorMap of C viewed as [M][N] fp32 (row stride ldc) with a (32 columns, 32 rows) box, 128B swizzle
typedef CUresult (*tc_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static tc_encode_fn tc_encoder() {
    static tc_encode_fn enc = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            enc = (tc_encode_fn)p;
    }
    return enc;
}

static int tc_make_tmap_c(CUtensorMap* tm, const float* C, long long ldc, int M, int N) {
    memset(tm, 0, sizeof(*tm));
    tc_encode_fn enc = tc_encoder();
    if (!enc) return 0;
    const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)ldc * 4};
    const cuuint32_t box[2] = {32, 32};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)C, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 1 : 0;
}

// CUtensorMap of A viewed as [M][K] fp32 (row stride lda) with a (32 columns, 128 rows) box, 128B swizzle: one box = one raw
// 128 x 32 chunk in exactly the K-major swizzled layout the converters (and tcgen05) expect; rows >= M are zero-filled
static int tc_make_tmap_a(CUtensorMap* tm, const float* A, long long lda, int M, int K) {
    memset(tm, 0, sizeof(*tm));
    tc_encode_fn enc = tc_encoder();
    if (!enc) return 0;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)lda * 4};
    const cuuint32_t box[2] = {TC_BK, TC_BM};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)A, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 1 : 0;
}

// Row-grouped operands: logical row m = g * R + i (i < R) lives at physical row g * S + i (the first R of every S rows: the agent
// rows of an [N, ne, d] entity tensor).  Rank-3 maps (column, i, g) whose boxes (32, R, 128 / R) resp. (32, R, 32 / R) land in
// shared memory exactly like the rank-2 boxes (32 columns x 128 resp. 32 rows, 128B swizzle), so the kernel only changes coordinates.
static int tc_make_tmap_grouped(CUtensorMap* tm, const float* P, long long ld, int G, int cols, int R, int S, int box_rows, bool load) {
    memset(tm, 0, sizeof(*tm));
    tc_encode_fn enc = tc_encoder();
    if (!enc) return 0;
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)R, (cuuint64_t)G};
    const cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)S * ld * 4};
    const cuuint32_t box[3] = {32, (cuuint32_t)R, (cuuint32_t)(box_rows / R)};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)P, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, load ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 1 : 0;
}

// CUtensorMap of a row-major [M][P] fp32 matrix with a (128 columns, 32 rows) box, no swizzle: one raw reduction chunk of
// the weight-gradient kernel; columns >= P and rows >= M are zero-filled
static int tc_make_tmap_rows(CUtensorMap* tm, const float* X, long long ldx, int M, int P, int box_cols = 128) {
    memset(tm, 0, sizeof(*tm));
    tc_encode_fn enc = tc_encoder();
    if (!enc) return 0;
    const cuuint64_t dims[2] = {(cuuint64_t)P, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)ldx * 4};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, 32};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)X, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 1 : 0;
}

// the same for row groups (first R of every S rows): rank-3 map (column, row inside the group, group), box (box_cols, R, 32 / R)
static int tc_make_tmap_rows_grouped(CUtensorMap* tm, const float* X, long long ldx, int G, int P, int R, int S, int box_cols) {
    memset(tm, 0, sizeof(*tm));
    tc_encode_fn enc = tc_encoder();
    if (!enc) return 0;
    const cuuint64_t dims[3] = {(cuuint64_t)P, (cuuint64_t)R, (cuuint64_t)G};
    const cuuint64_t strides[2] = {(cuuint64_t)ldx * 4, (cuuint64_t)S * ldx * 4};
    const cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)R, (cuuint32_t)(32 / R)};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)X, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 1 : 0;
}

// resident split weight tile (2 * BN * KS fp32) must leave room for >= 2 A stages and the staging tiles
#define TC_B_BUDGET (128 * 1024)

// k-slices: 1 if a 32-wide n-tile of the whole reduction fits the budget comfortably, else the smallest divisor of K
// whose slice (a multiple of 32) lets a min(N, 128)-wide tile stay resident
static int tc_pick_slices(int N, int K) {
    const int bn = N < 128 ? N : 128;
    for (int s = 1; s <= 8; s++) {
        if (K % s) continue;
        const int ks = K / s;
        if (ks % TC_BK) continue;
        if ((long long)2 * bn * ks * 4 <= TC_B_BUDGET) return s;
    }
    return 0;
}

// n-tile width: multiple of 32 dividing N, <= 256, with the resident split weight within budget
static int tc_pick_bn(int N, int KS, int cap = 256) {
    int best = 0;
    for (int bn = 32; bn <= cap && bn <= N; bn += 32)
        if (N % bn == 0 && (long long)2 * bn * KS * 4 <= TC_B_BUDGET) best = bn;
    return best;
}

// Can this problem run on the tensor-core path?  (else the caller uses the fp32 FFMA kernel)
extern "C" int refil_tc_gemm_supported(int M, int N, int K) {
    if (M < 1 || K < TC_BK || K % TC_BK != 0 || N < 32 || N % 32 != 0) return 0;
    const int s = tc_pick_slices(N, K);
    if (s < 1) return 0;
    return tc_pick_bn(N, K / s) > 0;
}

extern "C" int refil_tc_gemm_k_slices(int N, int K) {
    if (K < TC_BK || K % TC_BK != 0 || N < 32 || N % 32 != 0) return 0;
    return tc_pick_slices(N, K);
}

// m-tiles a CTA of the dense kernel should at least get (REFIL_TC_MIN_TILES, default 6; refil_tc_set_min_tiles overrides it for the
// launches that follow: the learner asks for short kernels on its critical chain and long-lived CTAs on the side chains)
static int g_min_tiles_override = 0;
static int tc_min_tiles() {
    static int env_tiles = -1;
    if (env_tiles < 0) {
        const char* e = getenv("REFIL_TC_MIN_TILES");
        env_tiles = e ? atoi(e) : 6;
        if (env_tiles < 1) env_tiles = 1;
    }
    return g_min_tiles_override > 0 ? g_min_tiles_override : env_tiles;
}
extern "C" int refil_tc_set_min_tiles(int min_tiles) {
    const int prev = g_min_tiles_override;
    g_min_tiles_override = min_tiles > 0 ? min_tiles : 0;
    return prev;
}

static int tc_mode_ts() {
    // operand path: "ts" (default) keeps the A operand in tensor memory (TMA-fed raw ring + converter warps), "ss" is the
    // all-shared-memory kernel; REFIL_TC_MODE=ss selects the latter for A/B comparisons
    static int mode_ts = -1;
    if (mode_ts < 0) {
        const char* e = getenv("REFIL_TC_MODE");
        mode_ts = (e && e[0] == 's' && e[1] == 's') ? 0 : 1;
    }
    return mode_ts;
}

// one problem of a group: validate, fill TcArgs and the two tensor maps
static int tc_fill(const RefilGemmDesc& d, int N_group, int K_group, int mode_ts, TcArgs& a, CUtensorMap* tmap_c, CUtensorMap* tmap_a) {
    const int M = d.M;
    const int K = d.k_len > 0 ? d.k_len : K_group;        // ... and reduce over fewer / more k-slices of the same length
    const int N = d.n_cols > 0 ? d.n_cols : N_group;      // a problem may be narrower / wider than the group's N (same n-tile width)
    REFIL_CHECK_ARG(d.A && d.B && d.C, "tc_gemm_tn: null pointer");
    REFIL_CHECK_ARG(refil_tc_gemm_supported(M, N, K), "tc_gemm_tn: unsupported shape M=%d N=%d K=%d", M, N, K);
    REFIL_CHECK_ARG((d.lda % 4) == 0 && (d.ldc % 4) == 0 && ((uintptr_t)d.A % 16) == 0 && ((uintptr_t)d.C % 16) == 0,
                    "tc_gemm_tn: A / C must be 16-byte aligned with leading dimensions divisible by 4");
    REFIL_CHECK_ARG(!d.relu_y || ((d.ldy % 4) == 0 && ((uintptr_t)d.relu_y % 16) == 0), "tc_gemm_tn: relu_y alignment");
    REFIL_CHECK_ARG(d.b_k_valid >= 0 && d.b_k_valid <= K, "tc_gemm_tn: b_k_valid=%d outside [0, K=%d]", d.b_k_valid, K);
    a = TcArgs{};
    a.kb_valid = d.b_k_valid > 0 ? d.b_k_valid : K;
    a.b_vec = (d.b_stride_k == 1 && (d.b_stride_n % 4) == 0 && ((uintptr_t)d.B % 16) == 0) ? 1 : 0;
    a.A = d.A; a.lda = d.lda; a.relu_y = d.relu_y; a.ldy = d.ldy;
    a.a_rowmask = d.a_row_entity_mask; a.a_na = d.a_na > 0 ? d.a_na : 1; a.a_ne = d.a_ne;
    a.a_mper = d.a_rows_per_copy > 0 ? d.a_rows_per_copy : 1;
    a.B = d.B; a.sbj = d.b_stride_n; a.sbi = d.b_stride_k;
    a.bias = d.bias; a.relu = d.relu;
    a.c_rowmask = d.c_row_entity_mask; a.c_na = d.c_na > 0 ? d.c_na : 1; a.c_ne = d.c_ne;
    a.c_mper = d.c_rows_per_copy > 0 ? d.c_rows_per_copy : 1;
    a.M = M; a.N = N; a.K = K;
    a.k_slices = tc_pick_slices(N, K);
    a.KS = K / a.k_slices;
    REFIL_CHECK_ARG(a.k_slices == 1 || (!d.bias && !d.relu && !d.c_row_entity_mask),
                    "tc_gemm_tn: a sliced reduction (K=%d) cannot carry a non-linear epilogue", K);
    const int BN = tc_pick_bn(N, a.KS, mode_ts ? 128 : 256);      // ts: two accumulators + the operand ring share 512 columns
    a.BN = BN;
    a.n_tiles = N / BN;
    a.m_tiles = refil_cdiv(M, TC_BM);
    // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, both K-major, N>>3, M>>4
    a.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    a.accumulate = d.accumulate ? 1 : 0;
    REFIL_CHECK_ARG(!a.accumulate || (!d.bias && !d.relu && !d.c_row_entity_mask), "tc_gemm_tn: accumulate cannot carry a non-linear epilogue");
    if (d.row_group > 0) {
        const int R = d.row_group;
        REFIL_CHECK_ARG(mode_ts && (32 % R) == 0 && d.row_group_stride >= R && (M % R) == 0,
                        "tc_gemm_tn: row groups of %d rows (stride %d, M=%d) need the tensor-memory path, R | 32 and R | M", R,
                        d.row_group_stride, M);
        REFIL_CHECK_ARG(!d.relu_y && !d.a_row_entity_mask && !d.c_row_entity_mask, "tc_gemm_tn: row groups cannot carry row functors");
        a.row_group = R;
        if (!tc_make_tmap_grouped(tmap_c, d.C, d.ldc, M / R, N, R, d.row_group_stride, 32, false) ||
            !tc_make_tmap_grouped(tmap_a, d.A, d.lda, M / R, K, R, d.row_group_stride, TC_BM, true)) {
            refil_set_error("tc_gemm_tn: cuTensorMapEncodeTiled failed for row groups (R=%d S=%d M=%d N=%d K=%d)", R, d.row_group_stride, M, N, K);
            return REFIL_ERR_CUDA;
        }
        return REFIL_OK;
    }
    if (!tc_make_tmap_c(tmap_c, d.C, d.ldc, M, N) || (mode_ts && !tc_make_tmap_a(tmap_a, d.A, d.lda, M, K))) {
        refil_set_error("tc_gemm_tn: cuTensorMapEncodeTiled failed (A=%p lda=%lld C=%p ldc=%lld M=%d N=%d K=%d)", (const void*)d.A,
                        d.lda, (const void*)d.C, d.ldc, M, N, K);
        return REFIL_ERR_CUDA;
    }
    return REFIL_OK;
}

extern "C" int refil_tc_gemm_tn_group(const RefilGemmDesc* descs, int n_problems, int N, int K, cudaStream_t stream) {
    REFIL_CHECK_ARG(descs && n_problems >= 1 && n_problems <= TC_MAX_GROUP, "tc_gemm_tn_group: 1..%d problems (got %d)",
                    TC_MAX_GROUP, n_problems);
    const int mode_ts = tc_mode_ts();
    REFIL_CHECK_ARG(mode_ts || n_problems == 1, "tc_gemm_tn_group: grouped launches need the tensor-memory operand path");
    TcGroup grp{};
    TcMaps mc{}, ma{};
    int max_tiles = 1;
    for (int g = 0; g < n_problems; g++) {
        int rc = tc_fill(descs[g], N, K, mode_ts, grp.a[g], &mc.m[g], &ma.m[g]);
        if (rc) return rc;
        if (grp.a[g].m_tiles > max_tiles) max_tiles = grp.a[g].m_tiles;
    }
    TcArgs& a0 = grp.a[0];
    const int BN = a0.BN;
    for (int g = 1; g < n_problems; g++)
        REFIL_CHECK_ARG(grp.a[g].BN == BN && grp.a[g].KS == a0.KS && grp.a[g].k_slices <= a0.k_slices,
                        "tc_gemm_tn_group: the problems of a group must share the n-tile width and the k-slice length (most slices first)");
    const size_t b_res = (size_t)2 * BN * a0.KS * 4, stage_bytes = 2 * (size_t)TC_BM * 128;
    const size_t stg_bytes = (size_t)TC_EPI_WARPS * TC_STG_BYTES;
    const size_t budget = 227 * 1024 - 1024 /* alignment slack */ - 2048 /* static: barriers, bias */;
    int stages = (int)((budget - b_res - stg_bytes) / stage_bytes);
    if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
    REFIL_CHECK_ARG(stages >= 2, "tc_gemm_tn: shared memory budget (N=%d K=%d)", N, K);
    for (int g = 0; g < n_problems; g++) grp.a[g].stages = stages;
    const size_t smem = mode_ts ? b_res + (size_t)TS_RAW_STAGES * TC_BM * 128 + stg_bytes + 1024
                                : b_res + stages * stage_bytes + stg_bytes + 1024;
    static size_t attr_smem[2] = {0, 0};
    if (smem > attr_smem[mode_ts]) {
        cudaError_t e = mode_ts ? cudaFuncSetAttribute(tc_gemm_ts_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                : cudaFuncSetAttribute(tc_gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess && mode_ts)
            e = cudaFuncSetAttribute(tc_gemm_ts_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            refil_set_error("tc_gemm_tn: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
            return REFIL_ERR_CUDA;
        }
        attr_smem[mode_ts] = smem;
    }
    if (a0.k_slices > 1) {            // partial tiles are reduce-added into C
        for (int g = 0; g < n_problems; g++) {
            if (grp.a[g].accumulate) continue;
            REFIL_CHECK_ARG(grp.a[g].k_slices > 1 && grp.a[g].row_group == 0,
                            "tc_gemm_tn_group: next to a sliced reduction every other problem must accumulate");
            cudaError_t e = cudaMemset2DAsync(descs[g].C, (size_t)descs[g].ldc * 4, 0, (size_t)grp.a[g].N * 4, (size_t)descs[g].M, stream);
            if (e != cudaSuccess) {
                refil_set_error("tc_gemm_tn: cudaMemset2DAsync: %s", cudaGetErrorString(e));
                return REFIL_ERR_CUDA;
            }
        }
    }
    const int sms = refil_num_sms();
    // every CTA pays a fixed prologue (split of its resident weight tile, pipeline fill): give it at least `min_tiles` m-tiles, so
    // that a small problem (a 16-episode shard) leaves SMs to the independent networks running on the other streams
    const int min_tiles = tc_min_tiles();
    long long sum_units = 0;                                  // (m-tile, n-tile, k-slice) units of the whole launch
    for (int g = 0; g < n_problems; g++) sum_units += (long long)grp.a[g].m_tiles * grp.a[g].n_tiles * grp.a[g].k_slices;
    int begin = 0;
    for (int g = 0; g < n_problems; g++) {
        // CTAs per (n-tile, k-slice) of this problem: its share of one wave, in proportion to its m-tiles
        const int mt = grp.a[g].m_tiles, groups = grp.a[g].n_tiles * grp.a[g].k_slices;
        int per_g = (int)((long long)sms * mt / sum_units);
        if (per_g < 1) per_g = 1;
        const int want = refil_cdiv(mt, min_tiles);
        if (per_g > want) per_g = want;
        if (per_g > mt) per_g = mt;
        grp.cta_begin[g] = begin;
        begin += per_g * groups;
    }
    for (int g = n_problems; g <= TC_MAX_GROUP; g++) grp.cta_begin[g] = begin;
    const dim3 grid(begin, 1);
    if (mode_ts) {
        // a k-sliced problem is preceded by the memset of C (not a kernel): plain stream order there
        const bool pdl = a0.k_slices == 1;
        cudaError_t le = n_problems == 1 ? refil_launch(tc_gemm_ts_kernel<false>, grid, dim3(TS_THREADS), smem, stream, pdl, grp, mc, ma)
                                         : refil_launch(tc_gemm_ts_kernel<true>, grid, dim3(TS_THREADS), smem, stream, pdl, grp, mc, ma);
        if (le != cudaSuccess) {
            refil_set_error("tc_gemm_tn (ts): launch failed: %s", cudaGetErrorString(le));
            return REFIL_ERR_CUDA;
        }
        REFIL_CHECK_LAUNCH("tc_gemm_tn (ts)");
        return REFIL_OK;
    }
    tc_gemm_tn_kernel<<<grid.x, TC_THREADS, smem, stream>>>(a0, mc.m[0]);
    REFIL_CHECK_LAUNCH("tc_gemm_tn");
    return REFIL_OK;
}

extern "C" int refil_tc_gemm_tn(const float* A, long long lda, const float* relu_y, long long ldy,
                                const uint8_t* a_row_entity_mask, int a_na, int a_ne, int a_rows_per_copy,
                                const float* B, long long b_stride_n, long long b_stride_k, int b_k_valid,
                                const float* bias, int relu, const uint8_t* c_row_entity_mask, int c_na, int c_ne,
                                int c_rows_per_copy, float* C, long long ldc, int M, int N, int K, cudaStream_t stream) {
    RefilGemmDesc d{A, lda, relu_y, ldy, a_row_entity_mask, a_na, a_ne, a_rows_per_copy, B, b_stride_n, b_stride_k, b_k_valid,
                    bias, relu, c_row_entity_mask, c_na, c_ne, c_rows_per_copy, C, ldc, M, 0, 0, 0, 0, 0};
    return refil_tc_gemm_tn_group(&d, 1, N, K, stream);
}

extern "C" int refil_tc_gemm_tn_rows(const float* A, long long lda, const float* B, long long b_stride_n, long long b_stride_k,
                                     float* C, long long ldc, int n_groups, int group_rows, int group_stride, int accumulate,
                                     int N, int K, cudaStream_t stream) {
    RefilGemmDesc d{};
    d.A = A; d.lda = lda; d.B = B; d.b_stride_n = b_stride_n; d.b_stride_k = b_stride_k; d.C = C; d.ldc = ldc;
    d.M = n_groups * group_rows; d.row_group = group_rows; d.row_group_stride = group_stride; d.accumulate = accumulate;
    return refil_tc_gemm_tn_group(&d, 1, N, K, stream);
}

// =====================================================================================================================
// Weight gradient on the tensor cores:  dW[P, Q] += g(X)[M, P]^T Y[M, Q],  db[P] += colsum g(X)
//   (autograd of the same nn.Linear layers; X = upstream gradient with fused relu' / row mask, Y = layer input)
// The reduction runs over the ROWS of both row-major operands, i.e. both are "MN-major" for the tensor core: a chunk of
// 32 rows is staged as-is in the only MN-major layout tf32 supports (cute::UMMA::Layout_MN_SW128_32B_Atom, descriptor layout
// type SWIZZLE_128B_BASE32B): atoms of 4 rows x 128 bytes, 32-byte groups XOR-swizzled with the row index, and tcgen05.mma
// transposes on the fly (instruction descriptor a_major = b_major = MN).  Each CTA owns one 128-wide
// p-tile and a contiguous range of row chunks, accumulates its partial [128 x BQ] tile in TMEM and adds it to dW with
// vector atomics.  The bias gradient is the extra column Q: the Y tile carries one more 32-wide atom whose first
// column is the constant 1 (written once), so db costs no extra pass over X.
// =====================================================================================================================
struct TcWArgs {
    const float* X; long long ldx;
    const float* relu_y; long long ldy;
    const uint8_t* x_rowmask; int na, ne, mper;
    const float* Y; long long ldyy;
    int y_shift, y_period;               // Y row m reads row m - y_shift, zero where (m / y_shift) % y_period == 0 (GRU h_{t-1})
    float* dW; long long lddw;
    float* db;
    int M, P, Q, BQ, p_tiles, splits, stages, chunks_per_split;
    int q_valid, dw_vec;                 // columns >= q_valid of dW do not exist (zero-padded Y); dw_vec: 16-byte atomics allowed
    int y_tma;                           // TS kernel: raw Y chunks arrive by TMA (y_shift == 0), else the producers load them
    int row_group;                       // > 0: row m of X and Y is physical row (m / row_group) * stride + m % row_group (rank-3 maps)
    uint32_t idesc;
};

// MN-major tf32: atoms of 4 k-rows x 128 B (512 B); LBO = stride between atoms along MN, SBO = between atoms along K
__device__ __forceinline__ uint64_t tc_smem_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                        // SWIZZLE_128B_BASE32B
    return d;
}

// float offset of (row r < 32, 16-byte chunk c16) inside a [32 rows][n_atoms * 32 floats] MN-major tile
__device__ __forceinline__ int tc_mn_off(int r, int c16, int n_atoms) {
    const int ka = r >> 2, kr = r & 3, at = c16 >> 3, c8 = c16 & 7;
    return (ka * n_atoms + at) * 128 + kr * 32 + ((((c8 >> 1) ^ kr) << 3) | ((c8 & 1) << 2));
}

__global__ void __launch_bounds__(TCW_THREADS, 1) tc_gemm_wgrad_kernel(TcWArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int BQ = a.BQ, S = a.stages;
    const int xa = 4, ya = BQ / 32;                         // MN atoms of the X tile (128 wide) and the Y tile
    const uint32_t x_bytes = 32 * 128 * 4, y_bytes = 32 * (uint32_t)BQ * 4;
    const uint32_t stage_bytes = 2 * x_bytes + 2 * y_bytes; // X_hi | X_lo | Y_hi | Y_lo
    __shared__ uint64_t full_bar[TC_MAX_STAGES], empty_bar[TC_MAX_STAGES], acc_bar;
    __shared__ uint32_t tmem_base_s;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform role index
    const int lane = threadIdx.x & 31;
    const int ptile = blockIdx.x % a.p_tiles, split = blockIdx.x / a.p_tiles;
    const int chunks_total = (a.M + 31) / 32;
    const int c_begin = split * a.chunks_per_split, c_end = min(chunks_total, c_begin + a.chunks_per_split);
    const int n_chunks = max(0, c_end - c_begin);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) { mbar_init(&full_bar[s], 128); mbar_init(&empty_bar[s], 1); }
        mbar_init(&acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    // constant "ones" atom of every stage's Y tile (hi: column Q = 1, rest 0; lo: 0), written once (bias gradient only)
    for (int s = 0; s < S && BQ > a.Q; s++) {
        float* yhi = reinterpret_cast<float*>(smem + (size_t)s * stage_bytes + 2 * x_bytes);
        float* ylo = reinterpret_cast<float*>(smem + (size_t)s * stage_bytes + 2 * x_bytes + y_bytes);
        for (int f = threadIdx.x; f < 32 * 32; f += TCW_THREADS) {         // 32 rows x 32 floats of the last MN atom
            const int r = f >> 5, e = f & 31;
            const int off = tc_mn_off(r, (ya - 1) * 8 + (e >> 2), ya) + (e & 3);
            yhi[off] = (e == 0) ? 1.f : 0.f;
            ylo[off] = 0.f;
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp < 4) {
        // ===================== epilogue: TMEM partial tile -> vector atomics into dW / db =====================
        if (n_chunks > 0) {
            mbar_wait(&acc_bar, 0);
            tc_fence_after();
            const int p = ptile * 128 + warp * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
            for (int c0 = 0; c0 < a.Q + (BQ > a.Q ? 16 : 0); c0 += 16) {
                uint32_t r[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr + (uint32_t)c0));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (p < a.P) {
                    if (c0 < a.Q) {
                        float* dst = a.dW + (long long)p * a.lddw + c0;
                        if (a.dw_vec) {
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                float4 v = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                                       __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                                atomicAdd(reinterpret_cast<float4*>(dst + 4 * q), v);
                            }
                        } else {             // dW rows are not 16-byte aligned (fc1: 53 columns) or end inside this slab
#pragma unroll
                            for (int q = 0; q < 16; q++)
                                if (c0 + q < a.q_valid) atomicAdd(dst + q, __uint_as_float(r[q]));
                        }
                    } else if (a.db) {
                        atomicAdd(a.db + p, __uint_as_float(r[0]));       // column Q: sum over rows of g(X)[:, p]
                    }
                }
            }
        }
    } else if (warp == 4) {
        // ===================== MMA issuer =====================
        // warp-uniform descriptors (uniform registers); one elected lane issues
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t smem0 = smem_u32(smem);
        const uint32_t idesc = a.idesc;
        const uint64_t x_step = (uint64_t)((xa * 1024) >> 4), y_step = (uint64_t)((ya * 1024) >> 4);
        const uint64_t x_lo_off = (uint64_t)(x_bytes >> 4), y_lo_off = (uint64_t)(y_bytes >> 4);
        int stage = 0;
        uint32_t phase = 0;
        for (int ch = 0; ch < n_chunks; ch++) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sx = smem0 + (uint32_t)stage * stage_bytes;
            const uint64_t x_hi = tc_smem_desc_mn(sx, 512, xa * 512);
            const uint64_t y_hi = tc_smem_desc_mn(sx + 2 * x_bytes, 512, ya * 512);
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ks++) {                          // 4 MMAs of K = 8 rows = 2 k-atoms each
                    const uint64_t xo = (uint64_t)ks * x_step, yo = (uint64_t)ks * y_step;
                    tc_mma(tmem_u, x_hi + x_lo_off + xo, y_hi + yo, idesc, (ch | ks) != 0);
                    tc_mma(tmem_u, x_hi + xo, y_hi + y_lo_off + yo, idesc, 1);
                    tc_mma(tmem_u, x_hi + xo, y_hi + yo, idesc, 1);
                }
                tc_commit(&empty_bar[stage]);
                if (ch == n_chunks - 1) tc_commit(&acc_bar);
            }
            __syncwarp();
            if (++stage == S) { stage = 0; phase ^= 1; }
        }
    } else {
        // ===================== producers (two groups, alternate chunks): straight 128-byte row-segment copies ====
        const int grp = warp < 9 ? 0 : 1;
        const int pt = threadIdx.x - (grp ? 288 : 160);
        // X tile: 32 rows x 32 chunks(16 B); thread -> fixed chunk cx of rows rx0 + 4 i  (rx0 < 4, i < 8)
        const int cx = pt & 31, rx0 = pt >> 5;
        const int pcol = ptile * 128 + cx * 4;
        const bool x_in = pcol < a.P;                     // P is a multiple of 4 (checked on the host)
        const int offx0 = tc_mn_off(rx0, cx, xa), offx_step = xa * 128;            // row + 4 -> next k-atom
        // Y tile: 32 rows x cpr chunks; thread -> fixed chunk cy of rows ry0 + ry_step i  (ry_step = 128 / cpr in {4,8,16})
        const int cpr = a.Q >> 2, ry0 = pt / cpr, cy = pt - ry0 * cpr, ry_step = 128 / cpr, ny = (32 * cpr) >> 7;
        const int offy0 = tc_mn_off(ry0, cy, ya), offy_step = (ry_step >> 2) * ya * 128;
        const float* px = a.X + pcol;
        const float* pr = a.relu_y ? a.relu_y + pcol : nullptr;
        const float* py = a.Y + cy * 4;
        int stage = 0;
        uint32_t phase = 0;
        for (int ch = 0; ch < n_chunks; ch++) {
            if ((ch & 1) == grp) {
                const long long m0 = (long long)(c_begin + ch) * 32;
                uint32_t valid = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const long long row = m0 + rx0 + 4 * i;
                    if (x_in && row < a.M && !tc_row_masked(a.x_rowmask, a.na, a.ne, a.mper, row)) valid |= 1u << i;
                }
                // every global load of the chunk is issued before the first use
                float4 vx[8], vr[8], vy[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    vx[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (valid & (1u << i)) vx[i] = __ldg(reinterpret_cast<const float4*>(px + (m0 + rx0 + 4 * i) * a.ldx));
                }
                if (pr) {
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        vr[i] = make_float4(1.f, 1.f, 1.f, 1.f);
                        if (valid & (1u << i)) vr[i] = __ldg(reinterpret_cast<const float4*>(pr + (m0 + rx0 + 4 * i) * a.ldy));
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    vy[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    const long long row = m0 + ry0 + ry_step * i;
                    if (i < ny && row < a.M) {
                        if (a.y_shift == 0) vy[i] = __ldg(reinterpret_cast<const float4*>(py + row * a.ldyy));
                        else if ((((unsigned)row) / (unsigned)a.y_shift) % (unsigned)a.y_period != 0u)
                            vy[i] = __ldg(reinterpret_cast<const float4*>(py + (row - a.y_shift) * a.ldyy));
                    }
                }
                if (pr) {
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        if (!(vr[i].x > 0.f)) vx[i].x = 0.f;
                        if (!(vr[i].y > 0.f)) vx[i].y = 0.f;
                        if (!(vr[i].z > 0.f)) vx[i].z = 0.f;
                        if (!(vr[i].w > 0.f)) vx[i].w = 0.f;
                    }
                }
                mbar_wait(&empty_bar[stage], phase ^ 1);
                float* xhi = reinterpret_cast<float*>(smem + (size_t)stage * stage_bytes) + offx0;
                float* xlo = xhi + 32 * 128;
                float* yhi = reinterpret_cast<float*>(smem + (size_t)stage * stage_bytes) + 2 * 32 * 128 + offy0;
                float* ylo = yhi + 32 * BQ;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(vx[i].x) & 0xffffe000u);
                    h.y = __uint_as_float(__float_as_uint(vx[i].y) & 0xffffe000u);
                    h.z = __uint_as_float(__float_as_uint(vx[i].z) & 0xffffe000u);
                    h.w = __uint_as_float(__float_as_uint(vx[i].w) & 0xffffe000u);
                    l.x = vx[i].x - h.x; l.y = vx[i].y - h.y; l.z = vx[i].z - h.z; l.w = vx[i].w - h.w;
                    *reinterpret_cast<float4*>(xhi + i * offx_step) = h;
                    *reinterpret_cast<float4*>(xlo + i * offx_step) = l;
                }
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (i < ny) {
                        float4 h, l;
                        h.x = __uint_as_float(__float_as_uint(vy[i].x) & 0xffffe000u);
                        h.y = __uint_as_float(__float_as_uint(vy[i].y) & 0xffffe000u);
                        h.z = __uint_as_float(__float_as_uint(vy[i].z) & 0xffffe000u);
                        h.w = __uint_as_float(__float_as_uint(vy[i].w) & 0xffffe000u);
                        l.x = vy[i].x - h.x; l.y = vy[i].y - h.y; l.z = vy[i].z - h.z; l.w = vy[i].w - h.w;
                        *reinterpret_cast<float4*>(yhi + i * offy_step) = h;
                        *reinterpret_cast<float4*>(ylo + i * offy_step) = l;
                    }
                }
                fence_async_smem();
                mbar_arrive(&full_bar[stage]);
            }
            if (++stage == S) { stage = 0; phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u));
    }
}

// =====================================================================================================================
// TS variant of the weight gradient: X^T is the tcgen05 "A" operand in TENSOR MEMORY.
// The tensor core wants A as lane = output row (p, a column of X) x column = reduction index (m, a row of X): exactly the
// transpose of the row-major gradient chunk.  A TMA tensor load drops the raw [32 rows][128 columns] chunk into a smem ring
// (un-swizzled; columns beyond P and rows beyond M are zero-filled by the hardware), and the four converter warps read it
// COLUMN-wise -- lane = p, one coalesced 128-byte LDS per row -- apply relu' / the row mask, split hi / lo and tcgen05.st
// their 32 values as 32 TMEM columns: the transposition costs nothing.  Only Y (the layer input, MN-major hi/lo tiles with
// the ones-atom for the bias gradient) is staged through registers by the two producer groups, now with the next chunk's
// loads in flight.  Shared-memory traffic per 32-row chunk: 132 KB instead of 180 KB, and the X path holds no registers
// while its loads are in flight.
//   warps 0-3  converters during the main loop, then the epilogue (TMEM partial tile -> vector atomics into dW / db)
//   warp  4    TMEM allocation + MMA issue ([tmem] A, MN-major smem B)
//   warp  5    TMA producer (X chunks, and the matching relu_y chunks when the relu' mask is fused)
//   warps 6-13 two Y producer groups taking alternate chunks
// =====================================================================================================================
#define TW_MMA_WARP 4
#define TW_TMA_WARP 5
#define TW_Y_WARP0 6
#define TW_THREADS (32 * 14)
#define TW_X_STAGES 4                     // TMEM operand ring (64 columns per stage: hi | lo) and max raw smem ring
#define TW_X_COL0 256

struct TcWGeom {
    int raw_stages, y_stages, y_raw_stages;
};

struct TcWGroup { TcWArgs a[TC_MAX_GROUP]; int cta_begin[TC_MAX_GROUP + 1]; };   // grouped: CTAs [cta_begin[p], cta_begin[p+1]) -> problem p

// grouped launch: one grid, every problem owns exactly its p_tiles x splits CTAs (own operands, row count and P)
template <bool GROUPED>
__global__ void __launch_bounds__(TW_THREADS, 1) tc_wgrad_ts_kernel(const __grid_constant__ TcWGroup grp, TcWGeom g,
                                                                   const __grid_constant__ TcMaps maps_x,
                                                                   const __grid_constant__ TcMaps maps_r,
                                                                   const __grid_constant__ TcMaps maps_y) {
    int prob = 0, bid = (int)blockIdx.x;
    if (GROUPED) {
        while (prob + 1 < TC_MAX_GROUP && (int)blockIdx.x >= grp.cta_begin[prob + 1]) prob++;
        bid = (int)blockIdx.x - grp.cta_begin[prob];
    }
    const TcWArgs& a = grp.a[prob];
    const CUtensorMap& tmap_x = maps_x.m[prob];
    const CUtensorMap& tmap_r = maps_r.m[prob];
    const CUtensorMap& tmap_y = maps_y.m[prob];
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int BQ = a.BQ, SY = g.y_stages, RS = g.raw_stages, RY = g.y_raw_stages;
    const int ya = BQ / 32;                                   // MN atoms of the Y tile
    const bool has_r = a.relu_y != nullptr;
    const uint32_t raw_bytes = 32 * 128 * 4, y_bytes = 32 * (uint32_t)BQ * 4, y_stage_bytes = 2 * y_bytes;
    uint8_t* smem_x = smem;                                               // [RS][32][128] raw X chunks
    uint8_t* smem_rl = smem_x + (size_t)RS * raw_bytes;                   // [RS][32][128] raw relu_y chunks (if fused)
    uint8_t* smem_y = smem_rl + (has_r ? (size_t)RS * raw_bytes : 0);     // [SY][Y_hi | Y_lo]
    const uint32_t yraw_bytes = 32 * (uint32_t)a.Q * 4;
    uint8_t* smem_yr = smem_y + (size_t)SY * y_stage_bytes;               // [RY][32][Q] raw Y chunks (TMA path)
    __shared__ uint64_t xr_full[TW_X_STAGES], xr_empty[TW_X_STAGES], xa_full[TW_X_STAGES], xa_empty[TW_X_STAGES];
    __shared__ uint64_t yr_full[TW_X_STAGES], yr_empty[TW_X_STAGES];
    __shared__ uint64_t y_full[TC_MAX_STAGES], y_empty[TC_MAX_STAGES], acc_bar;
    __shared__ uint32_t tmem_base_s;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform role index
    const int lane = threadIdx.x & 31;
    const int ptile = bid % a.p_tiles, split = bid / a.p_tiles;
    const int chunks_total = (a.M + 31) / 32;
    const int c_begin = split * a.chunks_per_split, c_end = min(chunks_total, c_begin + a.chunks_per_split);
    const int n_chunks = max(0, c_end - c_begin);
    pdl_launch_dependents();

    // the TMA lane initialises the barriers and requests the first raw chunks before the rest of the prologue
    auto tma_chunk = [&](int st, int ch) {
        const int m0 = (c_begin + ch) * 32;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&xr_full[st])),
                     "r"(raw_bytes * (has_r ? 2u : 1u)) : "memory");
        if (a.row_group > 0) {            // row-grouped X (no relu_y): (column, row inside the group, group)
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                ::"r"(smem_u32(smem_x + (size_t)st * raw_bytes)), "l"(&tmap_x), "r"(ptile * 128), "r"(0), "r"(m0 / a.row_group),
                  "r"(smem_u32(&xr_full[st]))
                : "memory");
            return;
        }
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
            ::"r"(smem_u32(smem_x + (size_t)st * raw_bytes)), "l"(&tmap_x), "r"(ptile * 128), "r"(m0), "r"(smem_u32(&xr_full[st]))
            : "memory");
        if (has_r)
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                ::"r"(smem_u32(smem_rl + (size_t)st * raw_bytes)), "l"(&tmap_r), "r"(ptile * 128), "r"(m0),
                  "r"(smem_u32(&xr_full[st]))
                : "memory");
    };
    auto tma_ychunk = [&](int st, int ch) {
        const int m0 = (c_begin + ch) * 32;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&yr_full[st])), "r"(yraw_bytes) : "memory");
        if (a.row_group > 0)
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                ::"r"(smem_u32(smem_yr + (size_t)st * yraw_bytes)), "l"(&tmap_y), "r"(0), "r"(0), "r"(m0 / a.row_group),
                  "r"(smem_u32(&yr_full[st]))
                : "memory");
        else
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                ::"r"(smem_u32(smem_yr + (size_t)st * yraw_bytes)), "l"(&tmap_y), "r"(0), "r"(m0), "r"(smem_u32(&yr_full[st]))
                : "memory");
    };
    const int pre = min(n_chunks, RS);
    const int pre_y = a.y_tma ? min(n_chunks, RY) : 0;
    if (threadIdx.x == TW_TMA_WARP * 32) {
        for (int s = 0; s < TW_X_STAGES; s++) {
            mbar_init(&xr_full[s], 1); mbar_init(&xr_empty[s], 128);
            mbar_init(&xa_full[s], 128); mbar_init(&xa_empty[s], 1);
            mbar_init(&yr_full[s], 1); mbar_init(&yr_empty[s], 128);
        }
        for (int s = 0; s < TC_MAX_STAGES; s++) { mbar_init(&y_full[s], 128); mbar_init(&y_empty[s], 1); }
        mbar_init(&acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        pdl_wait();                                            // X (and relu_y) come from the previous kernels
        for (int ch = 0; ch < max(pre, pre_y); ch++) {
            if (ch < pre) tma_chunk(ch, ch);
            if (ch < pre_y) tma_ychunk(ch, ch);
        }
    }
    if (warp == TW_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    // constant "ones" atom of every stage's Y tile (hi: column Q = 1, rest 0; lo: 0), written once (bias gradient only)
    for (int s = 0; s < SY && BQ > a.Q; s++) {
        float* yhi = reinterpret_cast<float*>(smem_y + (size_t)s * y_stage_bytes);
        float* ylo = reinterpret_cast<float*>(smem_y + (size_t)s * y_stage_bytes + y_bytes);
        for (int f = threadIdx.x; f < 32 * 32; f += TW_THREADS) {         // 32 rows x 32 floats of the last MN atom
            const int r = f >> 5, e = f & 31;
            const int off = tc_mn_off(r, (ya - 1) * 8 + (e >> 2), ya) + (e & 3);
            yhi[off] = (e == 0) ? 1.f : 0.f;
            ylo[off] = 0.f;
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();
    const uint32_t tmem_base = tmem_base_s;

    if (warp < 4) {
        // ===================== converters: lane = p (column of X); raw chunk read column-wise -> TMEM rows =====================
        const int pl = warp * 32 + lane;                         // my output row p of the tile = my TMEM lane
        const uint32_t t_lane = (uint32_t)(warp * 32) << 16;
        int sr = 0, sa = 0;
        uint32_t phr = 0, pha = 0;
        for (int ch = 0; ch < n_chunks; ch++) {
            const long long m0 = (long long)(c_begin + ch) * 32;
            // rows of the chunk that survive the row mask (lane k evaluates row m0 + k)
            const long long myrow = m0 + lane;
            const uint32_t keep = __ballot_sync(0xffffffffu, myrow < a.M && !tc_row_masked(a.x_rowmask, a.na, a.ne, a.mper, myrow));
            mbar_wait(&xr_full[sr], phr);
            const float* xs = reinterpret_cast<const float*>(smem_x + (size_t)sr * raw_bytes) + pl;
            const float* rs = reinterpret_cast<const float*>(smem_rl + (size_t)sr * raw_bytes) + pl;
            mbar_wait(&xa_empty[sa], pha ^ 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + t_lane + (uint32_t)(TW_X_COL0 + sa * 64);
#pragma unroll
            for (int hf = 0; hf < 2; hf++) {
                float x[16];
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = xs[(hf * 16 + k) * 128];
                if (has_r) {
#pragma unroll
                    for (int k = 0; k < 16; k++)
                        if (!(rs[(hf * 16 + k) * 128] > 0.f)) x[k] = 0.f;
                }
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const float xv = ((keep >> (hf * 16 + k)) & 1u) ? x[k] : 0.f;
                    const uint32_t h = __float_as_uint(xv) & 0xffffe000u;
                    hi[k] = h;
                    lo[k] = __float_as_uint(xv - __uint_as_float(h));
                }
                tc_tmem_st16(taddr + (uint32_t)(hf * 16), hi);
                tc_tmem_st16(taddr + 32u + (uint32_t)(hf * 16), lo);
            }
            mbar_arrive(&xr_empty[sr]);                          // my column of the raw chunk has been consumed
            if (++sr == RS) { sr = 0; phr ^= 1; }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            mbar_arrive(&xa_full[sa]);
            if (++sa == TW_X_STAGES) { sa = 0; pha ^= 1; }
        }
        // ===================== epilogue: TMEM partial tile -> vector atomics into dW / db =====================
        if (n_chunks > 0) {
            mbar_wait(&acc_bar, 0);
            tc_fence_after();
            const int p = ptile * 128 + pl;
            const uint32_t tacc = tmem_base + t_lane;
            for (int c0 = 0; c0 < a.Q + (BQ > a.Q ? 16 : 0); c0 += 16) {
                uint32_t r[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(tacc + (uint32_t)c0));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (p < a.P) {
                    if (c0 < a.Q) {
                        float* dst = a.dW + (long long)p * a.lddw + c0;
                        if (a.dw_vec) {
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                float4 v = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                                       __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                                atomicAdd(reinterpret_cast<float4*>(dst + 4 * q), v);
                            }
                        } else {             // dW rows are not 16-byte aligned (fc1: 53 columns) or end inside this slab
#pragma unroll
                            for (int q = 0; q < 16; q++)
                                if (c0 + q < a.q_valid) atomicAdd(dst + q, __uint_as_float(r[q]));
                        }
                    } else if (a.db) {
                        atomicAdd(a.db + p, __uint_as_float(r[0]));       // column Q: sum over rows of g(X)[:, p]
                    }
                }
            }
        }
    } else if (warp == TW_MMA_WARP) {
        // ===================== MMA issuer: A = X^T hi / lo in tensor memory, B = MN-major Y tiles =====================
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t sy0 = smem_u32(smem_y);
        const uint32_t idesc = a.idesc;
        const uint64_t y_step = (uint64_t)((ya * 1024) >> 4), y_lo_off = (uint64_t)(y_bytes >> 4);
        int sx = 0, sy = 0;
        uint32_t phx = 0, phy = 0;
        for (int ch = 0; ch < n_chunks; ch++) {
            mbar_wait(&xa_full[sx], phx);
            mbar_wait(&y_full[sy], phy);
            tc_fence_after();
            const uint32_t x_hi = tmem_u + (uint32_t)(TW_X_COL0 + sx * 64), x_lo = x_hi + 32u;
            const uint64_t y_hi = tc_smem_desc_mn(sy0 + (uint32_t)sy * y_stage_bytes, 512, ya * 512);
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ks++) {                          // 4 MMAs of K = 8 rows each
                    const uint64_t yo = (uint64_t)ks * y_step;
                    const uint32_t xc = (uint32_t)(ks * 8);
                    tc_mma_ts(tmem_u, x_lo + xc, y_hi + yo, idesc, (ch | ks) != 0);
                    tc_mma_ts(tmem_u, x_hi + xc, y_hi + y_lo_off + yo, idesc, 1);
                    tc_mma_ts(tmem_u, x_hi + xc, y_hi + yo, idesc, 1);
                }
                tc_commit(&xa_empty[sx]);
                tc_commit(&y_empty[sy]);
                if (ch == n_chunks - 1) tc_commit(&acc_bar);
            }
            __syncwarp();
            if (++sx == TW_X_STAGES) { sx = 0; phx ^= 1; }
            if (++sy == SY) { sy = 0; phy ^= 1; }
        }
    } else if (warp == TW_TMA_WARP) {
        // ===================== TMA producer: raw X (and relu_y) chunks =====================
        // one lane streams the X (and relu_y) chunks and the raw Y chunks: two rings, issued in chunk order (a slot of either ring
        // frees once the MMAs of an EARLIER chunk have retired, and every earlier chunk of both rings has been requested by then)
        if (lane == 0) {
            int sx = pre % RS, sy = a.y_tma ? pre_y % RY : 0;
            uint32_t phx = (uint32_t)(pre / RS) & 1u, phy = a.y_tma ? (uint32_t)(pre_y / RY) & 1u : 0u;
            for (int ch = min(pre, a.y_tma ? pre_y : pre); ch < n_chunks; ch++) {
                if (ch >= pre) {
                    mbar_wait(&xr_empty[sx], phx ^ 1);
                    tma_chunk(sx, ch);
                    if (++sx == RS) { sx = 0; phx ^= 1; }
                }
                if (a.y_tma && ch >= pre_y) {
                    mbar_wait(&yr_empty[sy], phy ^ 1);
                    tma_ychunk(sy, ch);
                    if (++sy == RY) { sy = 0; phy ^= 1; }
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== Y producers (two groups, alternate chunks, next chunk's loads in flight) =====================
        const int grp = (warp - TW_Y_WARP0) >> 2;
        const int pt = (int)threadIdx.x - (TW_Y_WARP0 * 32 + grp * 128);
        // Y tile: 32 rows x cpr chunks; thread -> fixed chunk cy of rows ry0 + ry_step i  (ry_step = 128 / cpr in {4,8,16})
        const int cpr = a.Q >> 2, ry0 = pt / cpr, cy = pt - ry0 * cpr, ry_step = 128 / cpr, ny = (32 * cpr) >> 7;
        const int offy0 = tc_mn_off(ry0, cy, ya), offy_step = (ry_step >> 2) * ya * 128;
        const float* py = a.Y + cy * 4;
        auto load_y = [&](int ch, float4 (&vy)[8]) {
            const long long m0 = (long long)(c_begin + ch) * 32;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                vy[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                const long long row = m0 + ry0 + ry_step * i;
                if (i < ny && row < a.M) {
                    if (a.y_shift == 0) vy[i] = __ldg(reinterpret_cast<const float4*>(py + row * a.ldyy));
                    else if ((((unsigned)row) / (unsigned)a.y_shift) % (unsigned)a.y_period != 0u)
                        vy[i] = __ldg(reinterpret_cast<const float4*>(py + (row - a.y_shift) * a.ldyy));
                }
            }
        };
        auto store_y = [&](int ch, const float4 (&vy)[8]) {
            const int stage = ch % SY;
            const uint32_t phase = (uint32_t)(ch / SY) & 1u;
            mbar_wait(&y_empty[stage], phase ^ 1);
            float* yhi = reinterpret_cast<float*>(smem_y + (size_t)stage * y_stage_bytes) + offy0;
            float* ylo = yhi + 32 * BQ;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (i < ny) {
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(vy[i].x) & 0xffffe000u);
                    h.y = __uint_as_float(__float_as_uint(vy[i].y) & 0xffffe000u);
                    h.z = __uint_as_float(__float_as_uint(vy[i].z) & 0xffffe000u);
                    h.w = __uint_as_float(__float_as_uint(vy[i].w) & 0xffffe000u);
                    l.x = vy[i].x - h.x; l.y = vy[i].y - h.y; l.z = vy[i].z - h.z; l.w = vy[i].w - h.w;
                    *reinterpret_cast<float4*>(yhi + i * offy_step) = h;
                    *reinterpret_cast<float4*>(ylo + i * offy_step) = l;
                }
            }
            fence_async_smem();
            mbar_arrive(&y_full[stage]);
        };
        if (a.y_tma) {
            // raw chunk from the TMA ring (rows beyond M arrive as zeros): my 16-byte pieces, conflict-free LDS.128
            const float* yr0 = reinterpret_cast<const float*>(smem_yr) + (size_t)ry0 * a.Q + cy * 4;
            for (int ch = grp; ch < n_chunks; ch += 2) {
                const int st = ch % RY;
                mbar_wait(&yr_full[st], (uint32_t)(ch / RY) & 1u);
                const float* src = yr0 + (size_t)st * (yraw_bytes >> 2);
                float4 vy[8];
#pragma unroll
                for (int i = 0; i < 8; i++)
                    vy[i] = i < ny ? *reinterpret_cast<const float4*>(src + (size_t)ry_step * i * a.Q) : make_float4(0.f, 0.f, 0.f, 0.f);
                mbar_arrive(&yr_empty[st]);
                store_y(ch, vy);
            }
        }
        float4 va[8], vb[8];
        int ch = a.y_tma ? n_chunks : grp;
        if (ch < n_chunks) {
            load_y(ch, va);
            for (;;) {
                int cn = ch + 2;
                if (cn < n_chunks) load_y(cn, vb);
                store_y(ch, va);
                ch = cn;
                if (ch >= n_chunks) break;
                cn = ch + 2;
                if (cn < n_chunks) load_y(cn, va);
                store_y(ch, vb);
                ch = cn;
                if (ch >= n_chunks) break;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TW_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    }
}

extern "C" int refil_tc_wgrad_supported(int M, int P, int Q) {
    if (M < 32 || P < 1) return 0;       // P: valid columns of X (its row stride only has to be a multiple of 4 floats)
    return (Q == 32 || Q == 64 || Q == 128) ? 1 : 0;
}

static int tcw_min_chunks() {
    // every split ends with a [128 x Q] tile of atomics: at least `min_chunks` 32-row chunks of reduction per split
    static int min_chunks = -1;
    if (min_chunks < 0) {
        const char* e = getenv("REFIL_TC_MIN_CHUNKS");
        min_chunks = e ? atoi(e) : 24;
        if (min_chunks < 1) min_chunks = 1;
    }
    return min_chunks;
}

static int tcw_fill(const RefilWgradDesc& d, int P_group, int Q, int sm_share, TcWArgs& a) {
    const int M = d.M;
    const int P = d.p_cols > 0 ? d.p_cols : P_group;     // a problem may own fewer / more 128-wide p-tiles than the group's P
    REFIL_CHECK_ARG(d.X && d.Y && d.dW, "tc_gemm_wgrad: null pointer");
    REFIL_CHECK_ARG(refil_tc_wgrad_supported(M, P, Q), "tc_gemm_wgrad: unsupported shape M=%d P=%d Q=%d", M, P, Q);
    REFIL_CHECK_ARG((d.ldx % 4) == 0 && (d.ldyy % 4) == 0 && ((uintptr_t)d.X % 16) == 0 && ((uintptr_t)d.Y % 16) == 0 &&
                    ((uintptr_t)d.dW % 4) == 0, "tc_gemm_wgrad: alignment");
    REFIL_CHECK_ARG(d.q_valid >= 0 && d.q_valid <= Q && d.ldx >= (P + 3) / 4 * 4, "tc_gemm_wgrad: q_valid=%d / ldx=%lld", d.q_valid,
                    d.ldx);
    REFIL_CHECK_ARG(!d.relu_y || ((d.ldy % 4) == 0 && ((uintptr_t)d.relu_y % 16) == 0), "tc_gemm_wgrad: relu_y alignment");
    a = TcWArgs{};
    a.X = d.X; a.ldx = d.ldx; a.relu_y = d.relu_y; a.ldy = d.ldy;
    a.x_rowmask = d.x_row_entity_mask; a.na = d.na > 0 ? d.na : 1; a.ne = d.ne; a.mper = d.rows_per_copy > 0 ? d.rows_per_copy : 1;
    a.Y = d.Y; a.ldyy = d.ldyy; a.dW = d.dW; a.lddw = d.lddw; a.db = d.db;
    a.y_shift = d.y_shift_rows > 0 ? d.y_shift_rows : 0; a.y_period = d.y_period > 0 ? d.y_period : 1;
    a.M = M; a.P = P; a.Q = Q;
    a.q_valid = d.q_valid > 0 ? d.q_valid : Q;
    a.dw_vec = (a.q_valid == Q && (d.lddw % 4) == 0 && ((uintptr_t)d.dW % 16) == 0) ? 1 : 0;
    a.p_tiles = refil_cdiv(P, 128);
    if (d.row_group > 0) {
        REFIL_CHECK_ARG((32 % d.row_group) == 0 && d.row_group_stride >= d.row_group && (M % d.row_group) == 0 && !d.relu_y &&
                        !d.x_row_entity_mask && d.y_shift_rows <= 0,
                        "tc_gemm_wgrad: row groups of %d rows need R | 32, R | M and no relu_y / row mask / shifted Y", d.row_group);
        a.row_group = d.row_group;
    }
    const int chunks_total = refil_cdiv(M, 32);
    int splits = sm_share / a.p_tiles;
    if (splits < 1) splits = 1;
    if (splits > refil_cdiv(chunks_total, tcw_min_chunks())) splits = refil_cdiv(chunks_total, tcw_min_chunks());
    if (splits > chunks_total) splits = chunks_total;
    a.chunks_per_split = refil_cdiv(chunks_total, splits);
    a.splits = refil_cdiv(chunks_total, a.chunks_per_split);
    return REFIL_OK;
}

extern "C" int refil_tc_gemm_wgrad_group(const RefilWgradDesc* descs, int n_problems, int P, int Q, cudaStream_t stream) {
    REFIL_CHECK_ARG(descs && n_problems >= 1 && n_problems <= TC_MAX_GROUP, "tc_gemm_wgrad_group: 1..%d problems (got %d)",
                    TC_MAX_GROUP, n_problems);
    const int mode_ts = tc_mode_ts();
    REFIL_CHECK_ARG(mode_ts || n_problems == 1, "tc_gemm_wgrad_group: grouped launches need the tensor-memory operand path");
    TcWGroup grp{};
    const bool has_r = descs[0].relu_y != nullptr, has_b = descs[0].db != nullptr;
    int max_grid = 1;
    // every problem gets its share of one wave of CTAs in proportion to its work (rows x p-tiles); tcw_fill divides the share by the
    // p-tiles again to get the row splits
    long long sum_work = 0;
    auto work = [&](int g) {
        const int Pg = descs[g].p_cols > 0 ? descs[g].p_cols : P;
        return (long long)(descs[g].M > 0 ? descs[g].M : 1) * refil_cdiv(Pg, 128);
    };
    for (int g = 0; g < n_problems; g++) sum_work += work(g);
    for (int g = 0; g < n_problems; g++) {
        REFIL_CHECK_ARG((descs[g].relu_y != nullptr) == has_r && (descs[g].db != nullptr) == has_b,
                        "tc_gemm_wgrad_group: the problems of a group must agree on relu_y / db being present");
        int rc = tcw_fill(descs[g], P, Q, (int)((long long)refil_num_sms() * work(g) / sum_work), grp.a[g]);
        if (rc) return rc;
        grp.a[g].BQ = has_b ? Q + 32 : Q;   // the bias gradient rides as one extra 32-wide atom whose first column is 1
        grp.cta_begin[g] = g == 0 ? 0 : grp.cta_begin[g - 1] + grp.a[g - 1].p_tiles * grp.a[g - 1].splits;
        max_grid = grp.cta_begin[g] + grp.a[g].p_tiles * grp.a[g].splits;
    }
    for (int g = n_problems; g <= TC_MAX_GROUP; g++) grp.cta_begin[g] = max_grid;
    TcWArgs& a = grp.a[0];
    if (mode_ts) {
        // X^T operand in tensor memory (K-major by construction), Y tiles MN-major in shared memory
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(a.BQ >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        TcWGeom geo{};
        const size_t raw_bytes = 32 * 128 * 4, y_stage = 2 * (size_t)32 * a.BQ * 4;
        geo.raw_stages = has_r ? 3 : 4;
        const size_t raw_total = (size_t)geo.raw_stages * raw_bytes * (has_r ? 2 : 1);
        // Y: raw chunks by TMA into their own ring (the producers only split hi / lo out of shared memory) unless a problem reads
        // shifted rows (GRU h_{t-1}); REFIL_TCW_Y=ldg keeps the register-staged global loads (A/B runs)
        static int y_ldg = -1;
        if (y_ldg < 0) {
            const char* e = getenv("REFIL_TCW_Y");
            y_ldg = (e && e[0] == 'l') ? 1 : 0;
        }
        int y_tma = y_ldg ? 0 : 1;
        for (int g = 0; g < n_problems; g++)
            if (grp.a[g].y_shift != 0) y_tma = 0;
        for (int g = 0; g < n_problems; g++)
            REFIL_CHECK_ARG(grp.a[g].row_group == 0 || y_tma, "tc_gemm_wgrad: row groups need the TMA path of Y");
        const size_t yraw_bytes = (size_t)32 * Q * 4;
        // The raw ring depth must be EVEN: the two producer groups take alternate chunks, so with an even ring every slot always
        // belongs to the same group and that group's previous visit orders the slot's phases.  With an odd ring a group could ask
        // for chunk c + 3 of a slot while chunk c (the other group's) is still in flight -- TMA completions are not ordered -- and the
        // one-bit parity wait would pass a phase early (seen as a launch failure at 39 chunks per CTA with all SMs loaded).
        geo.y_raw_stages = y_tma ? ((225 * 1024 - raw_total - 4 * yraw_bytes) >= 2 * y_stage ? 4 : 2) : 0;
        int ys = (int)((225 * 1024 - raw_total - geo.y_raw_stages * yraw_bytes) / y_stage);
        if (ys > TC_MAX_STAGES) ys = TC_MAX_STAGES;
        if (y_tma && ys > 3) ys = 3;                  // the raw ring holds the latency; the split tiles only double-buffer the MMA
        REFIL_CHECK_ARG(ys >= 2, "tc_gemm_wgrad: shared memory budget (Q=%d)", Q);
        geo.y_stages = ys;
        TcMaps mx{}, mr{}, my{};
        for (int g = 0; g < n_problems; g++) {
            grp.a[g].idesc = idesc;
            grp.a[g].y_tma = y_tma;
            const RefilWgradDesc& d = descs[g];
            const int Pg = grp.a[g].P, R = grp.a[g].row_group;
            const bool ok = R > 0 ? (tc_make_tmap_rows_grouped(&mx.m[g], d.X, d.ldx, d.M / R, Pg, R, d.row_group_stride, 128) &&
                                     tc_make_tmap_rows_grouped(&my.m[g], d.Y, d.ldyy, d.M / R, Q, R, d.row_group_stride, Q))
                                  : (tc_make_tmap_rows(&mx.m[g], d.X, d.ldx, d.M, Pg) &&
                                     (!has_r || tc_make_tmap_rows(&mr.m[g], d.relu_y, d.ldy, d.M, Pg)) &&
                                     (!y_tma || tc_make_tmap_rows(&my.m[g], d.Y, d.ldyy, d.M, Q, Q)));
            if (!ok) {
                refil_set_error("tc_gemm_wgrad: cuTensorMapEncodeTiled failed (X=%p ldx=%lld M=%d P=%d)", (const void*)d.X, d.ldx, d.M, P);
                return REFIL_ERR_CUDA;
            }
        }
        const size_t smem = raw_total + geo.y_raw_stages * yraw_bytes + (size_t)ys * y_stage + 1024;
        static size_t attr_smem_ts = 0;
        if (smem > attr_smem_ts) {
            cudaError_t e = cudaFuncSetAttribute(tc_wgrad_ts_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_wgrad_ts_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) {
                refil_set_error("tc_gemm_wgrad: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
                return REFIL_ERR_CUDA;
            }
            attr_smem_ts = smem;
        }
        cudaError_t le = n_problems == 1
            ? refil_launch(tc_wgrad_ts_kernel<false>, dim3(max_grid, 1), dim3(TW_THREADS), smem, stream, true, grp, geo, mx, mr, my)
            : refil_launch(tc_wgrad_ts_kernel<true>, dim3(max_grid, 1), dim3(TW_THREADS), smem, stream, true, grp, geo, mx, mr, my);
        if (le != cudaSuccess) {
            refil_set_error("tc_gemm_wgrad (ts): launch failed: %s", cudaGetErrorString(le));
            return REFIL_ERR_CUDA;
        }
        REFIL_CHECK_LAUNCH("tc_gemm_wgrad (ts)");
        return REFIL_OK;
    }
    const size_t stage_bytes = 2 * (size_t)32 * 128 * 4 + 2 * (size_t)32 * a.BQ * 4;
    int stages = (int)((200 * 1024) / stage_bytes);
    if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
    a.stages = stages;
    // D=f32, A=B=tf32, A and B MN-major (bits 15, 16), N = BQ, M = 128
    a.idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(a.BQ >> 3) << 17) |
              ((uint32_t)(128 >> 4) << 24);
    const size_t smem = stages * stage_bytes + 1024;
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            refil_set_error("tc_gemm_wgrad: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
            return REFIL_ERR_CUDA;
        }
        attr_smem = smem;
    }
    tc_gemm_wgrad_kernel<<<a.p_tiles * a.splits, TCW_THREADS, smem, stream>>>(a);
    REFIL_CHECK_LAUNCH("tc_gemm_wgrad");
    return REFIL_OK;
}

extern "C" int refil_tc_gemm_wgrad(const float* X, long long ldx, const float* relu_y, long long ldy,
                                   const uint8_t* x_row_entity_mask, int na, int ne, int rows_per_copy,
                                   const float* Y, long long ldyy, int y_shift_rows, int y_period, float* dW,
                                   long long lddw, int q_valid, float* db, int M, int P, int Q, cudaStream_t stream) {
    RefilWgradDesc d{X, ldx, relu_y, ldy, x_row_entity_mask, na, ne, rows_per_copy, Y, ldyy, y_shift_rows, y_period, dW, lddw,
                     q_valid, db, M};
    return refil_tc_gemm_wgrad_group(&d, 1, P, Q, stream);
}
