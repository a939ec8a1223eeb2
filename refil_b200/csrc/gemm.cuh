// FP32 SIMT GEMM building block (fixed summation order per tile, FFMA) used by every dense layer of the learner:
//   fwd   C = act(A W^T + b)            (nn.Linear of attention.py:21-22, entity_rnn_agent.py:12,23-25, flex_qmix.py:29,39)
//   bwd-d dA = mask(dC) W
//   bwd-w dW += mask(dC)^T A            (split over the long M axis, fp32 atomics into the flat grad buffer)
// Operands are "matrix functors" (row-major views with fused masks / virtual concatenation) so that the
// reference's cat / masked_fill / relu-backward passes never materialise in HBM.
//
// Tile: BM x BN x 16, 256 threads, TM x TN register micro-tile, double-buffered shared memory.
#pragma once
#include "common.cuh"

namespace rg {

constexpr int BK = 16;
constexpr int NTHREADS = 256;

// ---------------------------------------------------------------------------------------------
// matrix functors: at(r, c) with r < R, c < C guaranteed by the caller; at4 only if kVec and c%4==0, c+3<C
// ---------------------------------------------------------------------------------------------
struct RowMask {  // row r of a [copies, N, na] stack is masked iff entity_mask[n, a] != 0
    const uint8_t* em;  // [N, ne] or null
    int na, ne, mper;   // mper = N*na
    __device__ __forceinline__ bool masked(int r) const {
        if (!em) return false;
        int idx = r % mper;
        int n = idx / na, a = idx - n * na;
        return em[(size_t)n * ne + a] != 0;
    }
};

struct MatPlain {
    const float* p; int ld; int R, C;
    static constexpr bool kVec = true;
    __device__ __forceinline__ bool vec_ok() const { return (ld & 3) == 0 && ((uintptr_t)p & 15) == 0; }
    __device__ __forceinline__ float at(int r, int c) const { return __ldg(p + (size_t)r * ld + c); }
    __device__ __forceinline__ float4 at4(int r, int c) const {
        return __ldg(reinterpret_cast<const float4*>(p + (size_t)r * ld + c));
    }
};

// [ents | onehot(last_action) | 1?]  -- EntityMAC._build_inputs / QLearner._get_mixer_ins concat, never materialised
struct MatConcat {
    const float* ents; int ed; const int32_t* la; int A; int ones; int R, C;  // C = ed + A + ones
    static constexpr bool kVec = false;
    __device__ __forceinline__ bool vec_ok() const { return false; }
    __device__ __forceinline__ float at(int r, int c) const {
        if (c < ed) return __ldg(ents + (size_t)r * ed + c);
        if (c < ed + A) return (la[r] == c - ed) ? 1.f : 0.f;
        return 1.f;
    }
    __device__ __forceinline__ float4 at4(int r, int c) const { return make_float4(at(r, c), at(r, c + 1), at(r, c + 2), at(r, c + 3)); }
};

// upstream gradient with fused relu-backward (y > 0) and row mask
struct MatGrad {
    const float* p; int ld; const float* y; int ldy; RowMask rm; int R, C;
    static constexpr bool kVec = true;
    __device__ __forceinline__ bool vec_ok() const {
        return (ld & 3) == 0 && ((uintptr_t)p & 15) == 0 && (!y || ((ldy & 3) == 0 && ((uintptr_t)y & 15) == 0));
    }
    __device__ __forceinline__ float at(int r, int c) const {
        if (rm.masked(r)) return 0.f;
        float v = __ldg(p + (size_t)r * ld + c);
        if (y && !(__ldg(y + (size_t)r * ldy + c) > 0.f)) v = 0.f;
        return v;
    }
    __device__ __forceinline__ float4 at4(int r, int c) const {
        if (rm.masked(r)) return make_float4(0.f, 0.f, 0.f, 0.f);
        float4 v = __ldg(reinterpret_cast<const float4*>(p + (size_t)r * ld + c));
        if (y) {
            float4 m = __ldg(reinterpret_cast<const float4*>(y + (size_t)r * ldy + c));
            if (!(m.x > 0.f)) v.x = 0.f;
            if (!(m.y > 0.f)) v.y = 0.f;
            if (!(m.z > 0.f)) v.z = 0.f;
            if (!(m.w > 0.f)) v.w = 0.f;
        }
        return v;
    }
};

// h_{t-1} view of the GRU state stack HS[rows=(seq, t, agent), r]: row r -> HS[r - na] if t > 0 else 0
struct MatPrevT {
    const float* p; int ld; int na, T; int R, C;
    static constexpr bool kVec = true;
    __device__ __forceinline__ bool vec_ok() const { return (ld & 3) == 0 && ((uintptr_t)p & 15) == 0; }
    __device__ __forceinline__ bool first(int r) const { return ((r / na) % T) == 0; }
    __device__ __forceinline__ float at(int r, int c) const { return first(r) ? 0.f : __ldg(p + (size_t)(r - na) * ld + c); }
    __device__ __forceinline__ float4 at4(int r, int c) const {
        if (first(r)) return make_float4(0.f, 0.f, 0.f, 0.f);
        return __ldg(reinterpret_cast<const float4*>(p + (size_t)(r - na) * ld + c));
    }
};

// ---------------------------------------------------------------------------------------------
// tile loaders: fill Xs[BK][BX + PAD] for tile rows x0.. and reduction k0..
//   KC role: tile index x = matrix row, reduction k = matrix col (reduction-contiguous)
//   NC role: tile index x = matrix col, reduction k = matrix row (output-contiguous)
// ---------------------------------------------------------------------------------------------
constexpr int PAD = 4;

template <int BX, class Mat, bool KC>
struct TileLoader {
    static constexpr int NF4 = BX * BK / 4;                       // float4 slots in the tile
    static constexpr int PER = (NF4 + NTHREADS - 1) / NTHREADS;   // float4 slots per thread
    Mat mat;
    float4 r[PER];

    __device__ __forceinline__ void fetch(int x0, int k0, int Xdim, int Kend, int tid, bool vec) {
#pragma unroll
        for (int i = 0; i < PER; i++) {
            int f = tid + i * NTHREADS;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (NF4 % NTHREADS == 0 || f < NF4) {
                if (KC) {
                    int x = x0 + f / (BK / 4), k = k0 + (f % (BK / 4)) * 4;
                    if (x < Xdim) {
                        if (vec && k + 3 < Kend) v = mat.at4(x, k);
                        else {
                            if (k < Kend) v.x = mat.at(x, k);
                            if (k + 1 < Kend) v.y = mat.at(x, k + 1);
                            if (k + 2 < Kend) v.z = mat.at(x, k + 2);
                            if (k + 3 < Kend) v.w = mat.at(x, k + 3);
                        }
                    }
                } else {
                    int k = k0 + f / (BX / 4), x = x0 + (f % (BX / 4)) * 4;
                    if (k < Kend) {
                        if (vec && x + 3 < Xdim) v = mat.at4(k, x);
                        else {
                            if (x < Xdim) v.x = mat.at(k, x);
                            if (x + 1 < Xdim) v.y = mat.at(k, x + 1);
                            if (x + 2 < Xdim) v.z = mat.at(k, x + 2);
                            if (x + 3 < Xdim) v.w = mat.at(k, x + 3);
                        }
                    }
                }
            }
            r[i] = v;
        }
    }
    __device__ __forceinline__ void store(float* Xs, int tid) const {
#pragma unroll
        for (int i = 0; i < PER; i++) {
            int f = tid + i * NTHREADS;
            if (NF4 % NTHREADS == 0 || f < NF4) {
                if (KC) {
                    int x = f / (BK / 4), k = (f % (BK / 4)) * 4;
                    Xs[(k + 0) * (BX + PAD) + x] = r[i].x;
                    Xs[(k + 1) * (BX + PAD) + x] = r[i].y;
                    Xs[(k + 2) * (BX + PAD) + x] = r[i].z;
                    Xs[(k + 3) * (BX + PAD) + x] = r[i].w;
                } else {
                    int k = f / (BX / 4), x = (f % (BX / 4)) * 4;
                    *reinterpret_cast<float4*>(&Xs[k * (BX + PAD) + x]) = r[i];
                }
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------
// epilogues
// ---------------------------------------------------------------------------------------------
struct EpiStore {  // C = [relu]( acc + bias ) with optional row mask -> 0
    float* C; int ldc; const float* bias; int relu; RowMask rm;
    __device__ __forceinline__ bool row_flag(int m) const { return rm.masked(m); }
    __device__ __forceinline__ void operator()(int m, int n, float acc, bool masked) const {
        float v = acc + (bias ? __ldg(bias + n) : 0.f);
        if (relu) v = fmaxf(v, 0.f);
        if (masked) v = 0.f;
        C[(size_t)m * ldc + n] = v;
    }
};
struct EpiAtomic {  // split-reduction weight gradient; columns >= ncol_w go to the bias gradient (ones column)
    float* W; int ldw; int ncol_w; float* b;
    __device__ __forceinline__ bool row_flag(int) const { return false; }
    __device__ __forceinline__ void operator()(int m, int n, float acc, bool) const {
        if (n < ncol_w) atomicAdd(W + (size_t)m * ldw + n, acc);
        else if (b) atomicAdd(b + m, acc);
    }
};

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
template <int BM, int BN, int TM, int TN, class AMat, bool AKC, class BMat, bool BKC, class Epi>
__global__ void __launch_bounds__(NTHREADS, 2)
sgemm_kernel(AMat amat, BMat bmat, Epi epi, int M, int N, int K, int kchunk) {
    static_assert((BM / TM) * (BN / TN) == NTHREADS, "thread tiling");
    constexpr int TX = BN / TN;
    __shared__ __align__(16) float As[2][BK * (BM + PAD)];
    __shared__ __align__(16) float Bs[2][BK * (BN + PAD)];
    const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);

    TileLoader<BM, AMat, AKC> la{amat};
    TileLoader<BN, BMat, BKC> lb{bmat};
    const bool avec = AMat::kVec && amat.vec_ok(), bvec = BMat::kVec && bmat.vec_ok();

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

    auto mrow = [&](int i) { return (TM == 8) ? ((i < 4) ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4)) : ty * TM + i; };
    auto ncol = [&](int j) { return (TN == 8) ? ((j < 4) ? tx * 4 + j : BN / 2 + tx * 4 + (j - 4)) : tx * TN + j; };

    int buf = 0;
    if (kbeg < kend) {
        la.fetch(m0, kbeg, M, kend, tid, avec);
        lb.fetch(n0, kbeg, N, kend, tid, bvec);
        la.store(As[0], tid);
        lb.store(Bs[0], tid);
    }
    __syncthreads();
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        const bool more = k0 + BK < kend;
        if (more) {
            la.fetch(m0, k0 + BK, M, kend, tid, avec);
            lb.fetch(n0, k0 + BK, N, kend, tid, bvec);
        }
        const float* as = As[buf];
        const float* bs = Bs[buf];
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            float a[TM], b[TN];
            if constexpr (TM == 8) {
                float4 v0 = *reinterpret_cast<const float4*>(&as[kk * (BM + PAD) + ty * 4]);
                float4 v1 = *reinterpret_cast<const float4*>(&as[kk * (BM + PAD) + BM / 2 + ty * 4]);
                a[0] = v0.x; a[1] = v0.y; a[2] = v0.z; a[3] = v0.w;
                a[4] = v1.x; a[5] = v1.y; a[6] = v1.z; a[7] = v1.w;
            } else if constexpr (TM == 4) {
                float4 v0 = *reinterpret_cast<const float4*>(&as[kk * (BM + PAD) + ty * 4]);
                a[0] = v0.x; a[1] = v0.y; a[2] = v0.z; a[3] = v0.w;
            } else {
#pragma unroll
                for (int i = 0; i < TM; i++) a[i] = as[kk * (BM + PAD) + ty * TM + i];
            }
            if constexpr (TN == 8) {
                float4 v0 = *reinterpret_cast<const float4*>(&bs[kk * (BN + PAD) + tx * 4]);
                float4 v1 = *reinterpret_cast<const float4*>(&bs[kk * (BN + PAD) + BN / 2 + tx * 4]);
                b[0] = v0.x; b[1] = v0.y; b[2] = v0.z; b[3] = v0.w;
                b[4] = v1.x; b[5] = v1.y; b[6] = v1.z; b[7] = v1.w;
            } else if constexpr (TN == 4) {
                float4 v0 = *reinterpret_cast<const float4*>(&bs[kk * (BN + PAD) + tx * 4]);
                b[0] = v0.x; b[1] = v0.y; b[2] = v0.z; b[3] = v0.w;
            } else {
#pragma unroll
                for (int j = 0; j < TN; j++) b[j] = bs[kk * (BN + PAD) + tx * TN + j];
            }
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) {
            la.store(As[buf ^ 1], tid);
            lb.store(Bs[buf ^ 1], tid);
        }
        __syncthreads();
        buf ^= 1;
    }
    if (kbeg >= kend) return;
#pragma unroll
    for (int i = 0; i < TM; i++) {
        int m = m0 + mrow(i);
        if (m >= M) continue;
        const bool flag = epi.row_flag(m);
#pragma unroll
        for (int j = 0; j < TN; j++) {
            int n = n0 + ncol(j);
            if (n < N) epi(m, n, acc[i][j], flag);
        }
    }
}

// host-side dispatch on the output width
template <class AMat, bool AKC, class BMat, bool BKC, class Epi>
static inline void launch_sgemm(const AMat& a, const BMat& b, const Epi& e, int M, int N, int K, int splits,
                                cudaStream_t s) {
    int kchunk = ((refil_cdiv(K, splits) + BK - 1) / BK) * BK;
    splits = refil_cdiv(K, kchunk);
    if (N > 64) {
        dim3 g(refil_cdiv(N, 128), refil_cdiv(M, 128), splits);
        sgemm_kernel<128, 128, 8, 8, AMat, AKC, BMat, BKC, Epi><<<g, NTHREADS, 0, s>>>(a, b, e, M, N, K, kchunk);
    } else if (N > 32) {
        dim3 g(refil_cdiv(N, 64), refil_cdiv(M, 128), splits);
        sgemm_kernel<128, 64, 8, 4, AMat, AKC, BMat, BKC, Epi><<<g, NTHREADS, 0, s>>>(a, b, e, M, N, K, kchunk);
    } else if (N > 16) {
        dim3 g(refil_cdiv(N, 32), refil_cdiv(M, 128), splits);
        sgemm_kernel<128, 32, 4, 4, AMat, AKC, BMat, BKC, Epi><<<g, NTHREADS, 0, s>>>(a, b, e, M, N, K, kchunk);
    } else {
        dim3 g(refil_cdiv(N, 16), refil_cdiv(M, 128), splits);
        sgemm_kernel<128, 16, 4, 2, AMat, AKC, BMat, BKC, Epi><<<g, NTHREADS, 0, s>>>(a, b, e, M, N, K, kchunk);
    }
}

}  // namespace rg
