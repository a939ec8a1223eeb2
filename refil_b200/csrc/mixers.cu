// Attention-hypernetwork mixers: the per-(b, t) combine step after the hypernet GEMMs (SURVEY.md §8a rows L5-L7,
// kernel K7) and its backward.
//
// Replaces
//   /root/reference/src/modules/mixers/flex_qmix.py:51-57   (hypernet output modes matrix / vector / alt_vector / scalar)
//   /root/reference/src/modules/mixers/flex_qmix.py:79-121  (FlexQMixer.forward)
//   /root/reference/src/modules/mixers/flex_qmix.py:136-172 (LinearFlexQMixer.forward)
//   /root/reference/src/modules/mixers/vdn.py:9-10          (VDNMixer.forward)
// The plain mix and the REFIL "imagine" mix (2*na agent utilities, w1 = cat[w1(W), w1(I)]) of one (b, t) are
// evaluated together by one warp (lane = mixing-embed index m), because they share b1 / w_final / V.
//
// Inputs are the raw hypernet fc2 outputs, already zeroed for inactive agents (flex_qmix.py:50):
//   W1 [Cw, N, na, me]  (Cw = 1: default mask; Cw = 3: default, W, I)     B1, WF, V [N, na, me]
//   q [N, na] chosen utilities; qW, qI [N, na] utilities of the within / interact copies (imagine only)
// Outputs: q_tot [N], q_tot_im [N].
#include "common.cuh"

#define MIX_FLEX 0
#define MIX_LIN 1
#define MIX_VDN 2
#define MIX_MAX_NA 32

struct MixArgs {
    const float *W1, *B1, *WF, *V;
    const float *q, *qW, *qI;
    float *qtot, *qtot_im;
    float* ingroup;   // optional [N]: sum of the first na imagine mixing weights (flex_qmix.py:167-171, logging only)
    // backward
    const float *g_plain, *g_im;  // dL/dq_tot, dL/dq_tot_im  [N]
    float *dW1, *dB1, *dWF, *dV;  // same shapes as the inputs
    float *dq, *dqW, *dqI;
    int N, na, me, Cw, imagine, softmax_w, tanh_nl;
};

__device__ __forceinline__ float mix_act(float x, int tanh_nl) { return tanh_nl ? tanhf(x) : (x > 0.f ? x : expm1f(x)); }

// flex: one warp per n, lane = m
__global__ void __launch_bounds__(128) flex_mix_fwd_kernel(MixArgs a) {
    const int n = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= a.N) return;
    const int na = a.na, me = a.me;
    const bool on = lane < me;
    // b1 (vector), w_final (vector), v (scalar)
    float b1 = 0.f, wf = 0.f, vs = 0.f;
    for (int i = 0; i < na; i++) {
        const size_t o = ((size_t)n * na + i) * me + lane;
        if (on) { b1 += a.B1[o]; wf += a.WF[o]; vs += a.V[o]; }
    }
    b1 /= (float)na;
    wf /= (float)na;
    const float v = warp_sum(vs) / (float)(na * me);
    if (a.softmax_w) {
        const float mx = warp_max(on ? wf : -INFINITY);
        const float e = on ? expf(wf - mx) : 0.f;
        wf = e / warp_sum(e);
    } else {
        wf = fabsf(wf);
    }
    const int ncomb = a.imagine ? 2 : 1;
    for (int comb = 0; comb < ncomb; comb++) {
        float pre = b1;
        const int nparts = comb ? 2 : 1;
        for (int part = 0; part < nparts; part++) {
            const int c = comb ? 1 + part : 0;
            const float* qs = comb ? (part ? a.qI : a.qW) : a.q;
            for (int i = 0; i < na; i++) {
                float raw = on ? a.W1[(((size_t)c * a.N + n) * na + i) * me + lane] : 0.f;
                float w;
                if (a.softmax_w) {
                    const float mx = warp_max(on ? raw : -INFINITY);
                    const float e = on ? expf(raw - mx) : 0.f;
                    w = e / warp_sum(e);
                } else {
                    w = fabsf(raw);
                }
                pre = fmaf(qs[(size_t)n * na + i], w, pre);
            }
        }
        const float hid = on ? mix_act(pre, a.tanh_nl) : 0.f;
        const float y = warp_sum(hid * wf) + v;
        if (lane == 0) (comb ? a.qtot_im : a.qtot)[n] = y;
    }
}

__global__ void __launch_bounds__(128) flex_mix_bwd_kernel(MixArgs a) {
    const int n = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= a.N) return;
    const int na = a.na, me = a.me;
    const bool on = lane < me;
    float b1 = 0.f, wfr = 0.f;
    for (int i = 0; i < na; i++) {
        const size_t o = ((size_t)n * na + i) * me + lane;
        if (on) { b1 += a.B1[o]; wfr += a.WF[o]; }
    }
    b1 /= (float)na;
    wfr /= (float)na;
    float wf;
    if (a.softmax_w) {
        const float mx = warp_max(on ? wfr : -INFINITY);
        const float e = on ? expf(wfr - mx) : 0.f;
        wf = e / warp_sum(e);
    } else {
        wf = fabsf(wfr);
    }
    const int ncomb = a.imagine ? 2 : 1;
    float db1 = 0.f, dwf = 0.f, dv = 0.f;
    for (int comb = 0; comb < ncomb; comb++) {
        const float g = comb ? a.g_im[n] : a.g_plain[n];
        dv += g;
        const int nparts = comb ? 2 : 1;
        // recompute pre
        float pre = b1;
        for (int part = 0; part < nparts; part++) {
            const int c = comb ? 1 + part : 0;
            const float* qs = comb ? (part ? a.qI : a.qW) : a.q;
            for (int i = 0; i < na; i++) {
                float raw = on ? a.W1[(((size_t)c * a.N + n) * na + i) * me + lane] : 0.f;
                float w;
                if (a.softmax_w) {
                    const float mx = warp_max(on ? raw : -INFINITY);
                    const float e = on ? expf(raw - mx) : 0.f;
                    w = e / warp_sum(e);
                } else {
                    w = fabsf(raw);
                }
                pre = fmaf(qs[(size_t)n * na + i], w, pre);
            }
        }
        const float hid = on ? mix_act(pre, a.tanh_nl) : 0.f;
        dwf += g * hid;
        float dpre = 0.f;
        if (on) {
            const float dh = g * wf;
            dpre = a.tanh_nl ? dh * (1.f - hid * hid) : (pre > 0.f ? dh : dh * (hid + 1.f));
        }
        db1 += dpre;
        for (int part = 0; part < nparts; part++) {
            const int c = comb ? 1 + part : 0;
            const float* qs = comb ? (part ? a.qI : a.qW) : a.q;
            float* dqs = comb ? (part ? a.dqI : a.dqW) : a.dq;
            for (int i = 0; i < na; i++) {
                const size_t o = (((size_t)c * a.N + n) * na + i) * me + lane;
                float raw = on ? a.W1[o] : 0.f;
                const float qv = qs[(size_t)n * na + i];
                float w, draw;
                if (a.softmax_w) {
                    const float mx = warp_max(on ? raw : -INFINITY);
                    const float e = on ? expf(raw - mx) : 0.f;
                    w = e / warp_sum(e);
                    const float dw = dpre * qv;
                    const float t = warp_sum(w * dw);
                    draw = w * (dw - t);
                } else {
                    w = fabsf(raw);
                    draw = dpre * qv * (raw > 0.f ? 1.f : (raw < 0.f ? -1.f : 0.f));
                }
                const float dqv = warp_sum(dpre * w);
                if (on) a.dW1[o] = draw;
                if (lane == 0) dqs[(size_t)n * na + i] = dqv;
            }
        }
    }
    if (!a.imagine && a.Cw > 1) {  // copies that did not take part get a zero gradient
        for (int c = 1; c < a.Cw; c++)
            for (int i = 0; i < na; i++)
                if (on) a.dW1[(((size_t)c * a.N + n) * na + i) * me + lane] = 0.f;
    }
    float dwfr;
    if (a.softmax_w) {
        const float t = warp_sum(on ? wf * dwf : 0.f);
        dwfr = wf * (dwf - t);
    } else {
        dwfr = dwf * (wfr > 0.f ? 1.f : (wfr < 0.f ? -1.f : 0.f));
    }
    const float inv_na = 1.f / (float)na, dvs = dv / (float)(na * me);
    for (int i = 0; i < na; i++) {
        const size_t o = ((size_t)n * na + i) * me + lane;
        if (on) { a.dB1[o] = db1 * inv_na; a.dWF[o] = dwfr * inv_na; a.dV[o] = dvs; }
    }
}

// lin_flex: one warp per n, lane = agent slot (na or 2*na <= 32); W1 rows are reduced over me (alt_vector)
__global__ void __launch_bounds__(128) lin_mix_kernel(MixArgs a, int backward) {
    const int n = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= a.N) return;
    const int na = a.na, me = a.me;
    float vs = 0.f;
    for (int idx = lane; idx < na * me; idx += 32) vs += a.V[(size_t)n * na * me + idx];
    const float v = warp_sum(vs) / (float)(na * me);
    const int ncomb = a.imagine ? 2 : 1;
    float dv = 0.f;
    for (int comb = 0; comb < ncomb; comb++) {
        const int slots = comb ? 2 * na : na;
        const bool on = lane < slots;
        const int part = lane / na, i = lane - part * na;
        const int c = comb ? 1 + part : 0;
        float raw = 0.f, qv = 0.f;
        size_t base = 0;
        if (on) {
            base = (((size_t)c * a.N + n) * na + i) * me;
            for (int m = 0; m < me; m++) raw += a.W1[base + m];
            raw /= (float)me;
            const float* qs = comb ? (part ? a.qI : a.qW) : a.q;
            qv = qs[(size_t)n * na + i];
        }
        float w;
        if (a.softmax_w) {
            const float mx = warp_max(on ? raw : -INFINITY);
            const float e = on ? expf(raw - mx) : 0.f;
            w = e / warp_sum(e);
        } else {
            w = on ? fabsf(raw) : 0.f;
        }
        if (!backward) {
            const float y = warp_sum(qv * w) + v;
            if (lane == 0) (comb ? a.qtot_im : a.qtot)[n] = y;
            if (comb && a.ingroup) {
                const float ing = warp_sum(lane < na ? w : 0.f);
                if (lane == 0) a.ingroup[n] = ing;
            }
        } else {
            const float g = comb ? a.g_im[n] : a.g_plain[n];
            dv += g;
            const float dw = g * qv;
            float draw;
            if (a.softmax_w) {
                const float t = warp_sum(w * dw);
                draw = w * (dw - t);
            } else {
                draw = dw * (raw > 0.f ? 1.f : (raw < 0.f ? -1.f : 0.f));
            }
            if (on) {
                float* dqs = comb ? (part ? a.dqI : a.dqW) : a.dq;
                dqs[(size_t)n * na + i] = g * w;
                const float dm = draw / (float)me;
                for (int m = 0; m < me; m++) a.dW1[base + m] = dm;
            }
        }
    }
    if (backward) {
        if (!a.imagine && a.Cw > 1) {
            for (int c = 1; c < a.Cw; c++)
                for (int idx = lane; idx < na * me; idx += 32) a.dW1[((size_t)c * a.N + n) * na * me + idx] = 0.f;
        }
        const float dvs = dv / (float)(na * me);
        for (int idx = lane; idx < na * me; idx += 32) a.dV[(size_t)n * na * me + idx] = dvs;
    }
}

// vdn: thread per n
__global__ void vdn_mix_kernel(MixArgs a, int backward) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= a.N) return;
    const int na = a.na;
    if (!backward) {
        float s = 0.f;
        for (int i = 0; i < na; i++) s += a.q[(size_t)n * na + i];
        a.qtot[n] = s;
        if (a.imagine) {
            float s2 = 0.f;
            for (int i = 0; i < na; i++) s2 += a.qW[(size_t)n * na + i];
            for (int i = 0; i < na; i++) s2 += a.qI[(size_t)n * na + i];
            a.qtot_im[n] = s2;
        }
    } else {
        const float g = a.g_plain[n];
        for (int i = 0; i < na; i++) a.dq[(size_t)n * na + i] = g;
        if (a.imagine) {
            const float g2 = a.g_im[n];
            for (int i = 0; i < na; i++) { a.dqW[(size_t)n * na + i] = g2; a.dqI[(size_t)n * na + i] = g2; }
        }
    }
}

static int mix_check(const char* name, int kind, const MixArgs& a) {
    REFIL_CHECK_ARG(kind >= MIX_FLEX && kind <= MIX_VDN, "%s: unknown mixer kind %d", name, kind);
    REFIL_CHECK_ARG(a.N > 0 && a.na >= 1 && a.na <= MIX_MAX_NA, "%s: bad N=%d na=%d", name, a.N, a.na);
    REFIL_CHECK_ARG(a.q != nullptr, "%s: q is null", name);
    REFIL_CHECK_ARG(!a.imagine || (a.qW && a.qI), "%s: imagine needs qW and qI", name);
    if (kind != MIX_VDN) {
        REFIL_CHECK_ARG(a.me >= 1 && a.me <= 32, "%s: mixing_embed_dim %d outside [1,32]", name, a.me);
        REFIL_CHECK_ARG(a.W1 && a.V, "%s: hypernet outputs are null", name);
        REFIL_CHECK_ARG(a.Cw == 1 || a.Cw == 3, "%s: w1 copies must be 1 or 3", name);
        REFIL_CHECK_ARG(!a.imagine || a.Cw == 3, "%s: imagine needs 3 w1 copies", name);
        if (kind == MIX_FLEX) REFIL_CHECK_ARG(a.B1 && a.WF, "%s: flex mixer needs b1 and w_final", name);
        if (kind == MIX_LIN) REFIL_CHECK_ARG(2 * a.na <= 32 || !a.imagine, "%s: lin mixer supports 2*na <= 32", name);
    }
    return REFIL_OK;
}

extern "C" int refil_mixer_fwd(int kind, const float* W1, const float* B1, const float* WF, const float* V,
                               const float* q, const float* qW, const float* qI, float* qtot, float* qtot_im,
                               float* ingroup_out, int N, int n_agents, int mixing_embed, int w1_copies, int imagine,
                               int softmax_weights, int tanh_nonlin, cudaStream_t stream) {
    MixArgs a{};
    a.W1 = W1; a.B1 = B1; a.WF = WF; a.V = V; a.q = q; a.qW = qW; a.qI = qI; a.qtot = qtot; a.qtot_im = qtot_im;
    a.ingroup = ingroup_out;
    a.N = N; a.na = n_agents; a.me = mixing_embed; a.Cw = w1_copies; a.imagine = imagine;
    a.softmax_w = softmax_weights; a.tanh_nl = tanh_nonlin;
    int rc = mix_check("mixer_fwd", kind, a);
    if (rc) return rc;
    REFIL_CHECK_ARG(qtot && (!imagine || qtot_im), "mixer_fwd: output is null");
    REFIL_CHECK_ARG(!ingroup_out || (kind == MIX_LIN && imagine), "mixer_fwd: ingroup_out needs the linear mixer in imagine mode");
    if (kind == MIX_FLEX) flex_mix_fwd_kernel<<<refil_cdiv(N, 4), 128, 0, stream>>>(a);
    else if (kind == MIX_LIN) lin_mix_kernel<<<refil_cdiv(N, 4), 128, 0, stream>>>(a, 0);
    else vdn_mix_kernel<<<refil_cdiv(N, 128), 128, 0, stream>>>(a, 0);
    REFIL_CHECK_LAUNCH("mixer_fwd");
    return REFIL_OK;
}

extern "C" int refil_mixer_bwd(int kind, const float* W1, const float* B1, const float* WF, const float* V,
                               const float* q, const float* qW, const float* qI, const float* g_plain,
                               const float* g_im, float* dW1, float* dB1, float* dWF, float* dV, float* dq, float* dqW,
                               float* dqI, int N, int n_agents, int mixing_embed, int w1_copies, int imagine,
                               int softmax_weights, int tanh_nonlin, cudaStream_t stream) {
    MixArgs a{};
    a.W1 = W1; a.B1 = B1; a.WF = WF; a.V = V; a.q = q; a.qW = qW; a.qI = qI;
    a.g_plain = g_plain; a.g_im = g_im; a.dW1 = dW1; a.dB1 = dB1; a.dWF = dWF; a.dV = dV; a.dq = dq; a.dqW = dqW; a.dqI = dqI;
    a.N = N; a.na = n_agents; a.me = mixing_embed; a.Cw = w1_copies; a.imagine = imagine;
    a.softmax_w = softmax_weights; a.tanh_nl = tanh_nonlin;
    int rc = mix_check("mixer_bwd", kind, a);
    if (rc) return rc;
    REFIL_CHECK_ARG(g_plain && dq && (!imagine || (g_im && dqW && dqI)), "mixer_bwd: gradient pointer is null");
    if (kind != MIX_VDN) REFIL_CHECK_ARG(dW1 && dV, "mixer_bwd: hypernet gradient pointer is null");
    if (kind == MIX_FLEX) {
        REFIL_CHECK_ARG(dB1 && dWF, "mixer_bwd: flex needs dB1 and dWF");
        flex_mix_bwd_kernel<<<refil_cdiv(N, 4), 128, 0, stream>>>(a);
    } else if (kind == MIX_LIN) lin_mix_kernel<<<refil_cdiv(N, 4), 128, 0, stream>>>(a, 1);
    else vdn_mix_kernel<<<refil_cdiv(N, 128), 128, 0, stream>>>(a, 1);
    REFIL_CHECK_LAUNCH("mixer_bwd");
    return REFIL_OK;
}
