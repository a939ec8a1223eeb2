// Group Matching environment, batched on device (kernel K0).
//
// Replaces, for E parallel instances per launch:
//   /root/reference/src/envs/group_matching/group_matching.py:19-53  (step)
//   /root/reference/src/envs/group_matching/group_matching.py:91-106 (reset)
//   /root/reference/src/envs/group_matching/group_matching.py:55-76  (get_masks / get_entities)
// and the per-step buffer writes of /root/reference/src/runners/parallel_runner.py:117-197.
//
// Layout in HBM (all owned by the caller, PyTorch tensors):
//   mt_key  u32 [624][E]   MT19937 state, word-major / env-minor  -> thread==env loads are coalesced int32
//   mt_pos  i32 [E]        next word to twist (incremental twist == numpy's block regeneration)
//   loc     i32 [na][E]    ring position per agent
//   grp     u32 [ng][E]    group membership bitmask (bit a = agent a); groups may overlap / be empty
//   est     i32 [4][E]     prev_matches, t, flags(bit0 alive, bit1 in_list, bit2 solved, bit3 hit_limit), ep_len
//   ep_ret  f64 [E]        episode return (python float in the reference)
// Phase 1 of a step: one thread per env consumes the RNG stream sequentially (agents must draw in order) and
// resolves matches with bit-parallel occupancy masks (popc(members & occupancy[s]) == len, the single-thread
// form of a warp ballot over lanes=agents).  Phase 2: the CTA writes the (env, entity, feature) rows of the
// next timestep straight into the EpisodeBatch tensors with lane-contiguous stores.
#include "common.cuh"

#define MT_N 624
#define MT_M 397
#define GM_MAX_AGENTS 32
#define GM_MAX_GROUPS 8
#define GM_THREADS 128   // threads of a CTA: all of them write observation rows (phase 2)
#define GM_ENVS 32       // environments of a CTA: one warp runs their sequential logic (phase 1), four warps share the row writes

// Words of the generator a step may consume are PREFETCHED: the twist of word i reads the old words i, i + 1 and i + 397, none of
// which an earlier draw of the same step can have rewritten (that would need a look-back of 227 words), so the thread issues
// all those loads at once -- 2 W + 1 independent requests instead of a chain of dependent round trips (each draw used to wait
// ~0.5 us on L2; 17 draws per 8-agent step) -- parks them in its shared-memory lane and then runs the sequential draw logic out
// of shared memory.  Draws beyond the window (long rejection runs) fall back to the global path.
#define MT_WIN 32

struct MtRef {
    uint32_t* key;  // points at column e of mt_key
    int E;
    int pos;
    const uint32_t* win_old;   // [MT_WIN + 1] old words pos0 .. pos0 + MT_WIN, stride GM_ENVS (null: no window)
    const uint32_t* win_m;     // [MT_WIN] old words pos0 + 397 ...
    int used, nwin;            // draws taken from the window so far / words it holds
};

__device__ __forceinline__ uint32_t mt_next(MtRef& s) {
    int i = s.pos;
    int i1 = (i + 1 == MT_N) ? 0 : i + 1;
    int im = (i + MT_M >= MT_N) ? i + MT_M - MT_N : i + MT_M;
    uint32_t ki, k1, km;
    if (s.used < s.nwin) {
        ki = s.win_old[s.used * GM_ENVS];
        k1 = s.win_old[(s.used + 1) * GM_ENVS];
        km = s.win_m[s.used * GM_ENVS];
        s.used++;
    } else {
        ki = s.key[(size_t)i * s.E]; k1 = s.key[(size_t)i1 * s.E]; km = s.key[(size_t)im * s.E];
    }
    uint32_t y = (ki & 0x80000000u) | (k1 & 0x7fffffffu);
    y = km ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    s.key[(size_t)i * s.E] = y;
    s.pos = i1;
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

__device__ __forceinline__ double mt_double(MtRef& s) {
    uint32_t a = mt_next(s) >> 5, b = mt_next(s) >> 6;
    return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
}

__device__ __forceinline__ uint32_t mt_bounded(MtRef& s, uint32_t mx) {
    if (mx == 0) return 0;
    uint32_t mask = mx, v;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    do { v = mt_next(s) & mask; } while (v > mx);
    return v;
}

// fill the calling thread's window lane (sm_old / sm_m point at its own column of [..][GM_ENVS] shared arrays) with the first
// `n` <= MT_WIN words it may draw from position pos
__device__ __forceinline__ void mt_prefetch(const uint32_t* key, int E, int pos, int n, uint32_t* sm_old, uint32_t* sm_m) {
#pragma unroll 1
    for (int k0 = 0; k0 < n; k0 += 8) {
        uint32_t o[8], m[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            int i = pos + k0 + k;
            if (i >= MT_N) i -= MT_N;
            int im = i + MT_M;
            if (im >= MT_N) im -= MT_N;
            o[k] = key[(size_t)i * E];
            m[k] = key[(size_t)im * E];
        }
#pragma unroll
        for (int k = 0; k < 8; k++) { sm_old[(k0 + k) * GM_ENVS] = o[k]; sm_m[(k0 + k) * GM_ENVS] = m[k]; }
    }
    int i = pos + n;
    if (i >= MT_N) i -= MT_N;
    sm_old[n * GM_ENVS] = key[(size_t)i * E];
}

// numpy _legacy_seeding(int): init_genrand.  pos=0 in the incremental scheme == numpy's pos=624.
__global__ void gm_seed_kernel(uint32_t* __restrict__ mt_key, int32_t* __restrict__ mt_pos,
                               const uint32_t* __restrict__ seeds, int E) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    uint32_t s = seeds[e];
    for (int i = 0; i < MT_N; i++) {
        mt_key[(size_t)i * E + e] = s;
        s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)i + 1u;
    }
    mt_pos[e] = 0;
}

struct GmParams {
    int E, na, ne, ns, ng, ed, episode_limit, T;
    double rand_trans;
    int fixed_scen;
    int env_offset;  // first row of the rollout tensors this launch writes (rank-local batch index)
};

struct GmBuffers {
    uint32_t* mt_key; int32_t* mt_pos; int32_t* loc; uint32_t* grp; int32_t* est; double* ep_ret;
    // rollout tensors, EpisodeBatch layout [B, T, ...]; any may be null
    float* entities; uint8_t* obs_mask; uint8_t* entity_mask; uint8_t* gt_mask; int32_t* avail;
    float* reward; uint8_t* terminated; long long* filled;
    const long long* actions;  // [B, T, na, 1]
    unsigned long long* step_counter;  // += number of envs stepped
};

__device__ __forceinline__ int gm_matches(const int* loc, const uint32_t* grp, int na, int ns, int ng) {
    // occupancy[s] = bitmask of agents standing on s  (what __ballot_sync(loc==s) yields with lanes=agents)
    int m = 0;
    for (int g = 0; g < ng; g++) {
        uint32_t mem = grp[g];
        int len = __popc(mem), best = 0;
        for (int s = 0; s < ns; s++) {
            uint32_t occ = 0;
            for (int a = 0; a < na; a++) occ |= (loc[a] == s ? 1u : 0u) << a;
            best = max(best, __popc(occ & mem));
        }
        m += (best == len);
    }
    return m;
}

// exact floor(x / d) for 0 <= x < 2^22 with inv = 1.f / d: the runtime integer divisions of the index arithmetic below cost ~25
// instructions each and made the observation writes 80 % of the kernel's instructions (ncu r2o: 15.7 k instructions per warp and step)
__device__ __forceinline__ int gm_div(int x, float inv) { return (int)(((float)x + 0.5f) * inv); }

// Phase 2: write timestep `ts` observation rows for the CTA's envs from shared compact state.
__device__ void gm_write_obs(const GmParams& p, const GmBuffers& b, int ts, int e0, int n_env,
                             const uint8_t* s_loc, const uint32_t* s_grp, const uint8_t* s_write, bool write_gt) {
    const int chunk = p.ne * p.ed;
    const float inv_chunk = 1.f / (float)chunk, inv_ed = 1.f / (float)p.ed, inv_ne = 1.f / (float)p.ne;
    for (int idx = threadIdx.x; idx < n_env * chunk; idx += blockDim.x) {
        const int le = gm_div(idx, inv_chunk), off = idx - le * chunk;
        if (!s_write[le]) continue;
        const int a = gm_div(off, inv_ed), c = off - a * p.ed;
        float v = 0.f;
        if (a < p.na) {
            if (c < p.ns) v = (s_loc[le * GM_MAX_AGENTS + a] == c) ? 1.f : 0.f;
            else if (c < p.ns + p.ng) v = (float)((s_grp[le * GM_MAX_GROUPS + (c - p.ns)] >> a) & 1u);
            else v = (c - p.ns - p.ng == a) ? 1.f : 0.f;
        }
        size_t row = (size_t)(p.env_offset + e0 + le) * p.T + ts;
        b.entities[row * chunk + off] = v;
    }
    if (b.avail) {
        const int ch = p.na * 3;
        const float inv_ch = 1.f / (float)ch;
        for (int idx = threadIdx.x; idx < n_env * ch; idx += blockDim.x) {
            const int le = gm_div(idx, inv_ch), off = idx - le * ch;
            if (!s_write[le]) continue;
            size_t row = (size_t)(p.env_offset + e0 + le) * p.T + ts;
            b.avail[row * ch + off] = 1;
        }
    }
    if (p.ne > p.na) {  // padded entity slots (config 4): masked out exactly like absent SC2 units
        if (b.entity_mask) {
            for (int idx = threadIdx.x; idx < n_env * p.ne; idx += blockDim.x) {
                const int le = gm_div(idx, inv_ne), j = idx - le * p.ne;
                if (!s_write[le]) continue;
                size_t row = (size_t)(p.env_offset + e0 + le) * p.T + ts;
                b.entity_mask[row * p.ne + j] = (j >= p.na);
            }
        }
        if (b.obs_mask) {
            const int ch = p.ne * p.ne;
            const float inv_ch = 1.f / (float)ch;
            for (int idx = threadIdx.x; idx < n_env * ch; idx += blockDim.x) {
                const int le = gm_div(idx, inv_ch), off = idx - le * ch;
                if (!s_write[le]) continue;
                const int i = gm_div(off, inv_ne), j = off - i * p.ne;
                size_t row = (size_t)(p.env_offset + e0 + le) * p.T + ts;
                b.obs_mask[row * ch + off] = (i >= p.na || j >= p.na);
            }
        }
    }
    if (write_gt && b.gt_mask) {
        const int ch = p.na * p.ne;
        const float inv_ch = 1.f / (float)ch;
        for (int idx = threadIdx.x; idx < n_env * ch; idx += blockDim.x) {
            const int le = gm_div(idx, inv_ch), off = idx - le * ch;
            if (!s_write[le]) continue;
            const int ia = gm_div(off, inv_ne), j = off - ia * p.ne;
            uint8_t v = 1;
            if (j < p.na) {
                for (int g = 0; g < p.ng; g++) {  // FIRST group containing ia (group_matching.py:59-63)
                    uint32_t mem = s_grp[le * GM_MAX_GROUPS + g];
                    if ((mem >> ia) & 1u) { v = ((mem >> j) & 1u) ? 0 : 1; break; }
                }
            }
            size_t row = (size_t)(p.env_offset + e0 + le) * p.T + ts;
            b.gt_mask[row * ch + off] = v;
        }
    }
}

__global__ void __launch_bounds__(GM_THREADS) gm_reset_kernel(GmParams p, GmBuffers b) {
    __shared__ uint8_t s_loc[GM_ENVS * GM_MAX_AGENTS];
    __shared__ uint32_t s_grp[GM_ENVS * GM_MAX_GROUPS];
    __shared__ uint8_t s_write[GM_ENVS];
    const int e0 = blockIdx.x * GM_ENVS, le = threadIdx.x, e = e0 + le;
    const int n_env = min(GM_ENVS, p.E - e0);
    if (le < GM_ENVS) s_write[le] = 0;
    if (le < GM_ENVS && e < p.E) {
        MtRef rng{b.mt_key + e, p.E, b.mt_pos[e], nullptr, nullptr, 0, 0};
        int perm[GM_MAX_AGENTS];
        int parts[GM_MAX_GROUPS + 1];
        for (int i = 0; i < p.na; i++) perm[i] = i;
        parts[0] = 0; parts[p.ng] = p.na;
        if (!p.fixed_scen) {
            for (int i = p.na - 1; i >= 1; i--) {
                int j = (int)mt_bounded(rng, (uint32_t)i);
                int tmp = perm[i]; perm[i] = perm[j]; perm[j] = tmp;
            }
            for (int g = 1; g < p.ng; g++) parts[g] = (int)mt_bounded(rng, (uint32_t)(p.na - 1));
        } else {
            for (int g = 1; g < p.ng; g++) parts[g] = (int)rint((double)p.na * g / p.ng);
        }
        int loc[GM_MAX_AGENTS];
        uint32_t grp[GM_MAX_GROUPS];
        for (int g = 0; g < p.ng; g++) {
            uint32_t mem = 0;
            for (int k = parts[g]; k < parts[g + 1]; k++) mem |= 1u << perm[k];
            grp[g] = mem;
            b.grp[(size_t)g * p.E + e] = mem;
            s_grp[le * GM_MAX_GROUPS + g] = mem;
        }
        for (int a = 0; a < p.na; a++) {
            loc[a] = (int)mt_bounded(rng, (uint32_t)(p.ns - 1));
            b.loc[(size_t)a * p.E + e] = loc[a];
            s_loc[le * GM_MAX_AGENTS + a] = (uint8_t)loc[a];
        }
        b.mt_pos[e] = rng.pos;
        b.est[0 * (size_t)p.E + e] = gm_matches(loc, grp, p.na, p.ns, p.ng);
        b.est[1 * (size_t)p.E + e] = 0;
        b.est[2 * (size_t)p.E + e] = 3;  // alive | in_list
        b.est[3 * (size_t)p.E + e] = 0;
        b.ep_ret[e] = 0.0;
        s_write[le] = 1;
        if (b.filled) b.filled[(size_t)(p.env_offset + e) * p.T + 0] = 1;
    }
    __syncthreads();
    if (b.entities) gm_write_obs(p, b, 0, e0, n_env, s_loc, s_grp, s_write, true);
}

__global__ void __launch_bounds__(GM_THREADS) gm_step_kernel(GmParams p, GmBuffers b, int ts) {
    __shared__ uint8_t s_loc[GM_ENVS * GM_MAX_AGENTS];
    __shared__ uint32_t s_grp[GM_ENVS * GM_MAX_GROUPS];
    __shared__ uint8_t s_write[GM_ENVS];
    __shared__ int s_count;
    __shared__ uint32_t s_old[(MT_WIN + 1) * GM_ENVS], s_m[MT_WIN * GM_ENVS];
    const int e0 = blockIdx.x * GM_ENVS, le = threadIdx.x, e = e0 + le;
    const int n_env = min(GM_ENVS, p.E - e0);
    if (le == 0) s_count = 0;
    if (le < GM_ENVS) s_write[le] = 0;
    __syncthreads();
    if (le < GM_ENVS && e < p.E) {
        int flags = b.est[2 * (size_t)p.E + e];
        if (flags & 1) {
            const int pos0 = b.mt_pos[e];
            // window: 2 words per agent + headroom for the rand_trans re-draws; the rest of the state loads overlap it
            const int nwin = min(MT_WIN, ((2 * p.na + (p.rand_trans > 0.0 ? p.na / 2 + 4 : 0)) + 7) & ~7);
            mt_prefetch(b.mt_key + e, p.E, pos0, nwin, s_old + le, s_m + le);
            MtRef rng{b.mt_key + e, p.E, pos0, s_old + le, s_m + le, 0, nwin};
            int loc[GM_MAX_AGENTS];
            uint32_t grp[GM_MAX_GROUPS];
            const size_t row = (size_t)(p.env_offset + e) * p.T + ts;
            for (int g = 0; g < p.ng; g++) { grp[g] = b.grp[(size_t)g * p.E + e]; s_grp[le * GM_MAX_GROUPS + g] = grp[g]; }
            for (int a = 0; a < p.na; a++) {
                int l = b.loc[(size_t)a * p.E + e];
                int ac = (int)b.actions[row * p.na + a];
                if (mt_double(rng) < p.rand_trans) ac = (int)mt_bounded(rng, 2u);
                if (ac == 0) l = (l == 0) ? p.ns - 1 : l - 1;
                else if (ac == 2) l = (l + 1 >= p.ns) ? l + 1 - p.ns : l + 1;
                loc[a] = l;
                b.loc[(size_t)a * p.E + e] = l;
                s_loc[le * GM_MAX_AGENTS + a] = (uint8_t)l;
            }
            b.mt_pos[e] = rng.pos;
            int m = gm_matches(loc, grp, p.na, p.ns, p.ng);
            int prev = b.est[0 * (size_t)p.E + e];
            double rew = -0.1;
            rew += 2.5 * (double)(m - prev);
            b.est[0 * (size_t)p.E + e] = m;
            int t = b.est[1 * (size_t)p.E + e] + 1;
            b.est[1 * (size_t)p.E + e] = t;
            bool solved = (m == p.ng), limit = (t == p.episode_limit);
            bool done = solved || limit;
            // parallel_runner.py:180-183: stored `terminated` excludes time-limit endings
            if (b.reward) b.reward[row] = (float)rew;
            if (b.terminated) b.terminated[row] = (done && !limit) ? 1 : 0;
            if (b.filled) b.filled[row + 1] = 1;
            b.ep_ret[e] += rew;
            b.est[3 * (size_t)p.E + e] += 1;
            int nf = 2;                      // in_list for the next select_actions (alive before this step)
            if (!done) nf |= 1;
            if (solved) nf |= 4;
            if (limit) nf |= 8;
            b.est[2 * (size_t)p.E + e] = nf;
            s_write[le] = 1;
            atomicAdd(&s_count, 1);
        } else if (flags & 2) {
            b.est[2 * (size_t)p.E + e] = flags & ~2;
        }
    }
    __syncthreads();
    // gt_mask at ts + 1 only when the caller passes the tensor: EpisodeRunner stores it every step (episode_runner.py:66-67),
    // ParallelRunner at reset only (parallel_runner.py:140-150 drops the workers' per-step gt_mask)
    if (b.entities) gm_write_obs(p, b, ts + 1, e0, n_env, s_loc, s_grp, s_write, b.gt_mask != nullptr);
    if (threadIdx.x == 0 && b.step_counter && s_count) atomicAdd(b.step_counter, (unsigned long long)s_count);
}

// ------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------
static int gm_check(int E, int na, int ne, int ns, int ng) {
    REFIL_CHECK_ARG(E > 0, "gm_env: n_envs must be > 0");
    REFIL_CHECK_ARG(na >= 1 && na <= GM_MAX_AGENTS, "gm_env: n_agents %d outside [1,%d]", na, GM_MAX_AGENTS);
    REFIL_CHECK_ARG(ng >= 1 && ng <= GM_MAX_GROUPS, "gm_env: n_groups %d outside [1,%d]", ng, GM_MAX_GROUPS);
    REFIL_CHECK_ARG(ns >= 1 && ns <= 255, "gm_env: n_states %d outside [1,255]", ns);
    REFIL_CHECK_ARG(ne >= na, "gm_env: n_entities %d < n_agents %d", ne, na);
    return REFIL_OK;
}

extern "C" int refil_gm_env_seed(uint32_t* mt_key, int32_t* mt_pos, const uint32_t* seeds, int n_envs,
                                 cudaStream_t stream) {
    REFIL_CHECK_ARG(n_envs > 0 && mt_key && mt_pos && seeds, "gm_env_seed: bad arguments");
    gm_seed_kernel<<<refil_cdiv(n_envs, 128), 128, 0, stream>>>(mt_key, mt_pos, seeds, n_envs);
    REFIL_CHECK_LAUNCH("gm_env_seed");
    return REFIL_OK;
}

extern "C" int refil_gm_env_reset(uint32_t* mt_key, int32_t* mt_pos, int32_t* loc, uint32_t* grp, int32_t* est,
                                  double* ep_ret, float* entities, uint8_t* obs_mask, uint8_t* entity_mask,
                                  uint8_t* gt_mask, int32_t* avail_actions, long long* filled, int n_envs,
                                  int n_agents, int n_entities, int n_states, int n_groups, int fixed_scen,
                                  int episode_limit, int T, int env_offset, cudaStream_t stream) {
    int rc = gm_check(n_envs, n_agents, n_entities, n_states, n_groups);
    if (rc) return rc;
    GmParams p{n_envs, n_agents, n_entities, n_states, n_groups, n_states + n_groups + n_agents, episode_limit, T,
               0.0, fixed_scen, env_offset};
    GmBuffers b{mt_key, mt_pos, loc, grp, est, ep_ret, entities, obs_mask, entity_mask, gt_mask, avail_actions,
                nullptr, nullptr, filled, nullptr, nullptr};
    gm_reset_kernel<<<refil_cdiv(n_envs, GM_ENVS), GM_THREADS, 0, stream>>>(p, b);
    REFIL_CHECK_LAUNCH("gm_env_reset");
    return REFIL_OK;
}

extern "C" int refil_gm_env_step(uint32_t* mt_key, int32_t* mt_pos, int32_t* loc, uint32_t* grp, int32_t* est,
                                 double* ep_ret, const long long* actions, float* entities, uint8_t* obs_mask,
                                 uint8_t* entity_mask, uint8_t* gt_mask, int32_t* avail_actions, float* reward,
                                 uint8_t* terminated,
                                 long long* filled, unsigned long long* step_counter, int n_envs, int n_agents,
                                 int n_entities, int n_states, int n_groups, double rand_trans, int episode_limit,
                                 int T, int ts, int env_offset, cudaStream_t stream) {
    int rc = gm_check(n_envs, n_agents, n_entities, n_states, n_groups);
    if (rc) return rc;
    REFIL_CHECK_ARG(ts >= 0 && ts + 1 < T, "gm_env_step: ts=%d needs ts+1 < T=%d", ts, T);
    REFIL_CHECK_ARG(actions != nullptr, "gm_env_step: actions is null");
    GmParams p{n_envs, n_agents, n_entities, n_states, n_groups, n_states + n_groups + n_agents, episode_limit, T,
               rand_trans, 0, env_offset};
    GmBuffers b{mt_key, mt_pos, loc, grp, est, ep_ret, entities, obs_mask, entity_mask, gt_mask, avail_actions,
                reward, terminated, filled, actions, step_counter};
    gm_step_kernel<<<refil_cdiv(n_envs, GM_ENVS), GM_THREADS, 0, stream>>>(p, b, ts);
    REFIL_CHECK_LAUNCH("gm_env_step");
    return REFIL_OK;
}
