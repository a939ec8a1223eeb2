/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Not shipped, not on the product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file's shared object.
 *
 * Plain-C CPU restatement of the reference Group Matching environment
 *   /root/reference/src/envs/group_matching/group_matching.py
 * and of the NumPy legacy RandomState draw model it relies on (third-party dependency,
 * numpy 2.3.5 in this image; algorithm = MT19937 `init_genrand` seeding + legacy
 * `random_uniform`, `random_interval` / masked bounded uint32 rejection).
 *
 * Parity is PINNED: tests/test_oracle_env.py checks this port against
 *   (1) numpy.random.RandomState itself (u32 stream, uniform, randint, shuffle), and
 *   (2) transcripts of the reference GroupMatching class run in the build container,
 *       committed under tests/golden/gm_transcripts.npz (generator: tests/golden/make_golden.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MT_N 624
#define MT_M 397
#define GM_MAX_AGENTS 32
#define GM_MAX_GROUPS 8

typedef struct {
    uint32_t key[MT_N];
    int pos;
} mt_state;

/* numpy _legacy_seeding(int) -> mt19937_seed == Matsumoto/Nishimura init_genrand */
static void mt_seed(mt_state *s, uint32_t seed) {
    for (int i = 0; i < MT_N; i++) {
        s->key[i] = seed;
        seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u;
    }
    s->pos = MT_N;
}

static void mt_gen(mt_state *s) {
    uint32_t y;
    int i;
    for (i = 0; i < MT_N - MT_M; i++) {
        y = (s->key[i] & 0x80000000u) | (s->key[i + 1] & 0x7fffffffu);
        s->key[i] = s->key[i + MT_M] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (; i < MT_N - 1; i++) {
        y = (s->key[i] & 0x80000000u) | (s->key[i + 1] & 0x7fffffffu);
        s->key[i] = s->key[i + (MT_M - MT_N)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    y = (s->key[MT_N - 1] & 0x80000000u) | (s->key[0] & 0x7fffffffu);
    s->key[MT_N - 1] = s->key[MT_M - 1] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    s->pos = 0;
}

static uint32_t mt_u32(mt_state *s) {
    if (s->pos == MT_N) mt_gen(s);
    uint32_t y = s->key[s->pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

/* legacy random_uniform(0,1) == mt19937_next_double: 53-bit from two words */
static double mt_double(mt_state *s) {
    uint32_t a = mt_u32(s) >> 5, b = mt_u32(s) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

/* legacy bounded draw on [0, max]: mask, reject; rng==0 consumes nothing.
 * Used by RandomState.randint (masked uint32 path) and RandomState.shuffle (random_interval). */
static uint32_t mt_bounded(mt_state *s, uint32_t max) {
    if (max == 0) return 0;
    uint32_t mask = max, v;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    while ((v = (mt_u32(s) & mask)) > max) {}
    return v;
}

typedef struct {
    int n_agents, n_states, n_groups, episode_limit;
    double rand_trans;
    mt_state rng;
    int loc[GM_MAX_AGENTS];                 /* agent_locs as index (one-hot in the reference) */
    int grp_start[GM_MAX_GROUPS + 1];       /* partitions (unsorted; may be empty/negative-length) */
    int perm[GM_MAX_AGENTS];                /* shuffled agent ids; group g = perm[start_g:end_g] */
    uint32_t grp_members[GM_MAX_GROUPS];    /* bitmask of members */
    int grp_len[GM_MAX_GROUPS];
    int prev_matches, t;
} gm_env;

/* group_matching.py:6-17,114-118 */
gm_env *gm_create(int n_agents, int n_states, int n_groups, double rand_trans, int episode_limit, uint32_t seed) {
    if (n_agents > GM_MAX_AGENTS || n_groups > GM_MAX_GROUPS) return NULL;
    gm_env *e = (gm_env *)calloc(1, sizeof(gm_env));
    e->n_agents = n_agents; e->n_states = n_states; e->n_groups = n_groups;
    e->rand_trans = rand_trans; e->episode_limit = episode_limit;
    mt_seed(&e->rng, seed);
    return e;
}

void gm_destroy(gm_env *e) { free(e); }

/* group_matching.py:108-109 : sum_g [max_s sum_{a in g} loc[a,s] == len(g)].
 * An empty python slice has len 0 and sum(0).max()==0 -> counts as matched. */
static int gm_calc_group_piles(const gm_env *e) {
    int matches = 0;
    for (int g = 0; g < e->n_groups; g++) {
        int best = 0;
        for (int s = 0; s < e->n_states; s++) {
            int c = 0;
            for (int a = 0; a < e->n_agents; a++)
                if (((e->grp_members[g] >> a) & 1u) && e->loc[a] == s) c++;
            if (c > best) best = c;
        }
        if (best == e->grp_len[g]) matches++;
    }
    return matches;
}

/* group_matching.py:91-106 (fixed_scen=False branch; the fixed_scen branch uses np.int and
 * crashes on numpy >= 1.24, restated here as round(linspace) with no RNG draws for shuffle/partition) */
void gm_reset(gm_env *e, int fixed_scen) {
    int na = e->n_agents;
    for (int i = 0; i < na; i++) e->perm[i] = i;
    int parts[GM_MAX_GROUPS + 1];
    parts[0] = 0; parts[e->n_groups] = na;
    if (!fixed_scen) {
        for (int i = na - 1; i >= 1; i--) {               /* RandomState.shuffle on a list */
            int j = (int)mt_bounded(&e->rng, (uint32_t)i);
            int tmp = e->perm[i]; e->perm[i] = e->perm[j]; e->perm[j] = tmp;
        }
        for (int g = 1; g < e->n_groups; g++)             /* randint(0, na, size=n_groups-1) */
            parts[g] = (int)mt_bounded(&e->rng, (uint32_t)(na - 1));
    } else {
        for (int g = 1; g < e->n_groups; g++) {
            double x = (double)na * g / e->n_groups;      /* np.linspace(0,na,G+1).round(): half-to-even */
            double r = __builtin_rint(x);
            parts[g] = (int)r;
        }
    }
    for (int g = 0; g < e->n_groups; g++) {
        int s = parts[g], t = parts[g + 1];
        e->grp_members[g] = 0; e->grp_len[g] = 0;
        for (int k = s; k < t; k++) { e->grp_members[g] |= 1u << e->perm[k]; e->grp_len[g]++; }
    }
    memcpy(e->grp_start, parts, sizeof(int) * (e->n_groups + 1));
    for (int a = 0; a < na; a++)                          /* randint(0, n_states, size=na) */
        e->loc[a] = (int)mt_bounded(&e->rng, (uint32_t)(e->n_states - 1));
    e->prev_matches = gm_calc_group_piles(e);
    e->t = 0;
}

/* group_matching.py:19-53.  flags: bit0 done, bit1 solved, bit2 episode_limit.
 * reward is returned as the float64 the reference computes. */
int gm_step(gm_env *e, const int64_t *actions, double *reward) {
    for (int ia = 0; ia < e->n_agents; ia++) {
        int ac = (int)actions[ia];
        if (mt_double(&e->rng) < e->rand_trans) ac = (int)mt_bounded(&e->rng, 2u);
        if (ac == 0) e->loc[ia] = (e->loc[ia] == 0) ? e->n_states - 1 : e->loc[ia] - 1;
        else if (ac == 2) e->loc[ia] = (e->loc[ia] + 1 >= e->n_states) ? e->loc[ia] + 1 - e->n_states : e->loc[ia] + 1;
        /* any other value != 1 clears the cell in the reference (agent vanishes); actions are
         * always in {0,1,2} on the product path, so this port treats them as "stay". */
    }
    int m = gm_calc_group_piles(e);
    double rew = -0.1;
    rew += 2.5 * (double)(m - e->prev_matches);
    e->prev_matches = m;
    int flags = 0;
    if (m == e->n_groups) flags |= 1 | 2;
    e->t += 1;
    if (e->t == e->episode_limit) flags |= 1 | 4;
    *reward = rew;
    return flags;
}

/* group_matching.py:66-73: row a = [onehot(loc) | multi-hot(groups containing a) | onehot(a)] */
void gm_get_entities(const gm_env *e, float *out) {
    int ed = e->n_states + e->n_groups + e->n_agents;
    memset(out, 0, sizeof(float) * (size_t)e->n_agents * ed);
    for (int a = 0; a < e->n_agents; a++) {
        float *row = out + (size_t)a * ed;
        row[e->loc[a]] = 1.f;
        for (int g = 0; g < e->n_groups; g++)
            if ((e->grp_members[g] >> a) & 1u) row[e->n_states + g] = 1.f;
        row[e->n_states + e->n_groups + a] = 1.f;
    }
}

/* group_matching.py:55-64: gt_mask[ia, j] = 0 iff j in FIRST group containing ia; obs/entity masks zero */
void gm_get_masks(const gm_env *e, uint8_t *obs_mask, uint8_t *entity_mask, uint8_t *gt_mask) {
    int na = e->n_agents;
    memset(obs_mask, 0, (size_t)na * na);
    memset(entity_mask, 0, (size_t)na);
    memset(gt_mask, 1, (size_t)na * na);
    for (int ia = 0; ia < na; ia++)
        for (int g = 0; g < e->n_groups; g++)
            if ((e->grp_members[g] >> ia) & 1u) {
                for (int j = 0; j < na; j++)
                    if ((e->grp_members[g] >> j) & 1u) gt_mask[ia * na + j] = 0;
                break;
            }
}

void gm_get_locs(const gm_env *e, int32_t *out) { for (int a = 0; a < e->n_agents; a++) out[a] = e->loc[a]; }
int gm_get_t(const gm_env *e) { return e->t; }

/* raw RNG hooks so tests can pin the draw model against numpy.random.RandomState */
mt_state *mt_create(uint32_t seed) { mt_state *s = (mt_state *)malloc(sizeof(mt_state)); mt_seed(s, seed); return s; }
void mt_destroy(mt_state *s) { free(s); }
uint32_t mt_next_u32(mt_state *s) { return mt_u32(s); }
double mt_next_double(mt_state *s) { return mt_double(s); }
uint32_t mt_next_bounded(mt_state *s, uint32_t max) { return mt_bounded(s, max); }

/* Bounded CPU-baseline loop (bench.py cpu_baseline / --impl reference):
 * what env_worker does per step (parallel_runner.py:253-280): step + masks + entities, reset on done.
 * Actions come from a cheap LCG so the loop is self-contained.  Returns env-steps executed. */
long gm_bench_loop(gm_env *e, long n_steps, float *ent_buf, uint8_t *mask_buf) {
    int na = e->n_agents;
    int64_t act[GM_MAX_AGENTS];
    uint32_t lcg = 12345u;
    double rew, acc = 0.0;
    gm_reset(e, 0);
    for (long i = 0; i < n_steps; i++) {
        for (int a = 0; a < na; a++) { lcg = lcg * 1664525u + 1013904223u; act[a] = (lcg >> 16) % 3u; }
        int f = gm_step(e, act, &rew);
        acc += rew;
        gm_get_masks(e, mask_buf, mask_buf + na * na, mask_buf + na * na + na);
        gm_get_entities(e, ent_buf);
        if (f & 1) gm_reset(e, 0);
    }
    ent_buf[0] += (float)(acc * 0.0);
    return n_steps;
}
