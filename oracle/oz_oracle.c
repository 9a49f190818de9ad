/* TEST INFRASTRUCTURE ONLY -- see oz_oracle.h.  CPU restatement of the reference's
 * rules, encoding and batched virtual-loss MCTS, one reference statement per statement.
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (oracle/build_oracle.py).
 */
#include "oz_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ rules */

/* position.h:153 STRIDES; negative stride = left shift (position.h:174-184). */
static const int STRIDES[8] = {-9, -8, -7, -1, 1, 7, 8, 9};
#define MASK_NO_LEFT_RIGHT 0x7E7E7E7E7E7E7E7EULL /* position.h:155-156 */
#define MASK_NO_TOP_BOTTOM 0x00FFFFFFFFFFFF00ULL /* position.h:158-159 */
#define MASK_NO_EDGES (MASK_NO_LEFT_RIGHT & MASK_NO_TOP_BOTTOM)
/* position.h:163-172 */
static const uint64_t MASKS[8] = {MASK_NO_EDGES,      MASK_NO_TOP_BOTTOM, MASK_NO_EDGES,
                                  MASK_NO_LEFT_RIGHT, MASK_NO_LEFT_RIGHT, MASK_NO_EDGES,
                                  MASK_NO_TOP_BOTTOM, MASK_NO_EDGES};

static uint64_t shift_dir(int d, uint64_t m) {
    int s = STRIDES[d];
    return s > 0 ? (m >> s) : (m << -s);
}

/* position.h:186-196 */
static uint64_t potential_flips(int d, uint64_t player, uint64_t opp) {
    opp &= MASKS[d];
    uint64_t f = opp & shift_dir(d, player);
    for (int i = 0; i < 5; ++i) f |= opp & shift_dir(d, f);
    return f;
}

/* position.h:202-229 */
uint64_t ozo_get_legal_moves(uint64_t player, uint64_t opp) {
    uint64_t legal = 0;
    for (int d = 0; d < 8; ++d) legal |= shift_dir(d, potential_flips(d, player, opp));
    return legal & ~(player | opp);
}

/* position.h:231-262 */
uint64_t ozo_get_flips(uint64_t move, uint64_t player, uint64_t opp) {
    uint64_t flips = 0;
    for (int d = 0; d < 8; ++d) {
        uint64_t pf = potential_flips(d, move, opp);
        if ((shift_dir(d, pf) & player) != 0) flips |= pf;
    }
    return flips;
}

/* position.h:264-272 */
void ozo_initial_position(ozo_position *out) {
    out->player = 1;
    out->p1 = 0x0000000810000000ULL;
    out->p2 = 0x0000001008000000ULL;
    out->legal = ozo_get_legal_moves(out->p1, out->p2);
    out->next_legal = 0;
}

/* position.h:328-363 */
void ozo_apply_move(const ozo_position *p, uint64_t move, ozo_position *out) {
    int player = 3 - p->player;
    uint64_t p1 = p->p1, p2 = p->p2, legal, next_legal = 0;
    uint64_t *me = (p->player == 1) ? &p1 : &p2;
    uint64_t *opp = (p->player == 1) ? &p2 : &p1;
    uint64_t flips = ozo_get_flips(move, *me, *opp);
    *me |= move | flips;
    *opp &= ~flips;
    legal = ozo_get_legal_moves(*opp, *me);
    if (legal == 0) {
        next_legal = ozo_get_legal_moves(*me, *opp);
        if (next_legal == 0) player = 0;
    }
    out->player = player; out->p1 = p1; out->p2 = p2;
    out->legal = legal; out->next_legal = next_legal;
}

/* position.h:382-386 */
void ozo_apply_pass(const ozo_position *p, ozo_position *out) {
    ozo_position r;
    r.player = 3 - p->player; r.p1 = p->p1; r.p2 = p->p2;
    r.legal = p->next_legal; r.next_legal = 0;
    *out = r;
}

/* position.h:402-408 */
void ozo_apply_action(const ozo_position *p, int action, ozo_position *out) {
    if (action == 64) { ozo_apply_pass(p, out); return; }
    ozo_position r;
    ozo_apply_move(p, 1ULL << (63 - action), &r);
    *out = r;
}

/* position.h:308-326 */
int ozo_legal_actions(const ozo_position *p, int *actions) {
    if (p->player == 0) return 0;
    if (p->legal == 0) { actions[0] = 64; return 1; }
    int n = 0;
    uint64_t m = 1ULL << 63;
    for (int a = 0; a < 64; ++a) {
        if (m & p->legal) actions[n++] = a;
        m >>= 1;
    }
    return n;
}

/* perft convention of BASELINE.md section 3: pass = one ply, terminal = leaf. */
uint64_t ozo_perft(const ozo_position *p, int depth) {
    if (depth == 0 || p->player == 0) return 1;
    ozo_position c;
    if (p->legal == 0) { ozo_apply_pass(p, &c); return ozo_perft(&c, depth - 1); }
    uint64_t n = 0, moves = p->legal;
    while (moves) {
        uint64_t mv = moves & (~moves + 1);
        moves ^= mv;
        ozo_apply_move(p, mv, &c);
        n += ozo_perft(&c, depth - 1);
    }
    return n;
}

/* --------------------------------------------------------------- encoding */

/* transformation.h:40-57 */
int ozo_transform_action(int action, int t) {
    if (action == 64) return action;
    int row = action / 8, col = action % 8;
    if (t % 2 == 1) col = 7 - col;
    int rot = t / 2;
    for (int i = 0; i < rot; ++i) { int old = row; row = col; col = 7 - old; }
    return row * 8 + col;
}

/* transformation.h:83-116 */
void ozo_positions_to_features(const ozo_position *hist, int n_hist, float *f, int H, int t) {
    float side = (float)hist[0].player - 1.0f;
    for (int i = 0; i < 64; ++i) f[i] = side;
    f += 64;
    for (int k = 0; k < H; ++k) {
        if (k >= n_hist) {
            for (int i = 0; i < 128; ++i) f[i] = 0.0f;
        } else {
            uint64_t b = hist[k].p1, w = hist[k].p2, m = 1ULL << 63;
            for (int i = 0; i < 64; ++i) {
                int j = ozo_transform_action(i, t);
                f[j] = (b & m) != 0;
                f[64 + j] = (w & m) != 0;
                m >>= 1;
            }
        }
        f += 128;
    }
}

/* -------------------------------------------------------------------- RNG */

void ozo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t a = (uint64_t)0xD2511F53u * c0, b = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(b >> 32) ^ c1 ^ k0, n1 = (uint32_t)b;
        uint32_t n2 = (uint32_t)(a >> 32) ^ c3 ^ k1, n3 = (uint32_t)a;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

/* ----------------------------------------------------------------- search */

/* search_node.h:16-46, as an index-linked pool (parent links outlive re-rooting). */
typedef struct {
    ozo_position pos;
    int parent;       /* -1 above the initial position */
    int first_child;  /* children are contiguous, ascending action order; -1 = none */
    int n_children;
    int N;            /* visit_count */
    float W, Q, P;    /* total / mean action value, prior (default 1, search_node.h:45) */
} node_t;

struct ozo_mcts {
    int H, sims, T, B;
    float c_base, c_init, eps, alpha;
    uint64_t seed; uint32_t game_id, search_idx;
    const float *gamma_table; int gamma_len; int sym_mode;
    int eval_kind; ozo_eval_fn eval; void *eval_ctx;
    node_t *nodes; long n_nodes, cap; int root;
    /* per-search slot buffers (search_thread.cpp:25-44) */
    int inflight; int *leaves; int *tf; float *feat, *pol, *val;
};

static int new_node(ozo_mcts *m, const ozo_position *p, int parent, float prior) {
    if (m->n_nodes == m->cap) {
        m->cap = m->cap ? m->cap * 2 : 4096;
        m->nodes = (node_t *)realloc(m->nodes, (size_t)m->cap * sizeof(node_t));
    }
    node_t *n = &m->nodes[m->n_nodes];
    n->pos = *p; n->parent = parent; n->first_child = -1; n->n_children = 0;
    n->N = 0; n->W = 0.0f; n->Q = 0.0f; n->P = prior;
    return (int)m->n_nodes++;
}

ozo_mcts *ozo_mcts_create(int H, int sims, int T, int B, float c_base, float c_init,
                          float eps, float alpha) {
    ozo_mcts *m = (ozo_mcts *)calloc(1, sizeof(ozo_mcts));
    m->H = H; m->sims = sims; m->T = T; m->B = B;
    m->c_base = c_base; m->c_init = c_init; m->eps = eps; m->alpha = alpha;
    m->eval_kind = OZO_EVAL_SYNTH;
    ozo_mcts_reset_position(m);
    return m;
}

void ozo_mcts_destroy(ozo_mcts *m) {
    if (!m) return;
    free(m->nodes); free(m->leaves); free(m->tf); free(m->feat); free(m->pol); free(m->val);
    free(m);
}

void ozo_mcts_set_rng(ozo_mcts *m, uint64_t seed, uint32_t game_id, const float *table,
                      int len, int sym_mode) {
    m->seed = seed; m->game_id = game_id; m->gamma_table = table; m->gamma_len = len;
    m->sym_mode = sym_mode;
}

void ozo_mcts_set_evaluator(ozo_mcts *m, int kind, ozo_eval_fn fn, void *ctx) {
    m->eval_kind = kind; m->eval = fn; m->eval_ctx = ctx;
}

/* mcts.cpp:40-43 */
void ozo_mcts_reset_position(ozo_mcts *m) {
    ozo_position p;
    ozo_initial_position(&p);
    m->n_nodes = 0;
    m->root = new_node(m, &p, -1, 1.0f);
    m->search_idx = 0;
}

void ozo_mcts_position(const ozo_mcts *m, ozo_position *out) { *out = m->nodes[m->root].pos; }
long ozo_mcts_num_nodes(const ozo_mcts *m) { return m->n_nodes; }
int ozo_mcts_root_visit_count(const ozo_mcts *m) { return m->nodes[m->root].N; }

static void rng(const ozo_mcts *m, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) {
    uint32_t ctr[4] = {m->game_id, c1, c2, c3};
    uint32_t key[2] = {(uint32_t)m->seed, (uint32_t)(m->seed >> 32)};
    ozo_philox4x32_10(ctr, key, out);
}

/* Stand-in for std::gamma_distribution<float>(alpha, 1) (search_thread.cpp:23,233):
 * injected table, else Marsaglia-Tsang with the alpha < 1 boost. */
static float gamma_draw(const ozo_mcts *m, uint32_t sim, int child) {
    uint32_t r[4];
    if (m->gamma_table) {
        rng(m, m->search_idx, sim, 0x100u + (uint32_t)child, r);
        return m->gamma_table[r[0] % (uint32_t)m->gamma_len];
    }
    if (!(m->alpha > 0.0f)) return 0.0f;
    float a = m->alpha < 1.0f ? m->alpha + 1.0f : m->alpha;
    float d = a - 1.0f / 3.0f, c = 1.0f / sqrtf(9.0f * d);
    for (uint32_t attempt = 0;; ++attempt) {
        rng(m, m->search_idx, sim, 0x100u + (uint32_t)child + 64u * attempt, r);
        float u1 = u01(r[0]), u2 = u01(r[1]), u3 = u01(r[2]), u4 = u01(r[3]);
        float x = sqrtf(-2.0f * logf(u1)) * cosf(6.28318530717958647692f * u2);
        float v = 1.0f + c * x;
        if (v <= 0.0f) continue;
        v = v * v * v;
        if (logf(u3) < 0.5f * x * x + d - d * v + d * logf(v) || attempt >= 15) {
            float g = d * v;
            if (m->alpha < 1.0f) g *= powf(u4, 1.0f / m->alpha);
            return g;
        }
    }
}

/* search_thread.cpp:192-260 */
static int choose_best_child(ozo_mcts *m, int ni, uint32_t sim) {
    node_t *node = &m->nodes[ni];
    node_t *ch = &m->nodes[node->first_child];
    int n = node->n_children;
    if (n == 1) return node->first_child; /* :194-196 */

    /* :198-203  (1 + N + c_base): the int sum converts to float before the add */
    float rate = logf(((float)(1 + node->N) + m->c_base) / m->c_base) + m->c_init;
    int total = 0;
    for (int i = 0; i < n; ++i) total += ch[i].N; /* :204-207 */
    float mult = rate * sqrtf((float)total);       /* :208-209 */

    if (!(ni == m->root && m->eps > 0.0f)) { /* :211-228 */
        int best = 0;
        float best_ucb = ch[0].Q + mult * ch[0].P / (1.0f + (float)ch[0].N);
        for (int i = 1; i < n; ++i) {
            float ucb = ch[i].Q + mult * ch[i].P / (1.0f + (float)ch[i].N);
            if (ucb > best_ucb) { best = i; best_ucb = ucb; }
        }
        return node->first_child + best;
    }

    /* :230-259 root with Dirichlet noise, re-drawn on every call */
    float noises[40];
    float noise_sum = 0.0f;
    for (int i = 0; i < n; ++i) { noises[i] = gamma_draw(m, sim, i); noise_sum += noises[i]; }
    if (noise_sum == 0.0f) noise_sum = 1.0f;
    float pm = 1.0f - m->eps, nm = m->eps / noise_sum;
    int best = 0;
    float best_ucb = 0.0f;
    for (int i = 0; i < n; ++i) {
        float prob = ch[i].P * pm + noises[i] * nm;
        float ucb = ch[i].Q + mult * prob / (1.0f + (float)ch[i].N);
        if (i == 0 || ucb > best_ucb) { best = i; best_ucb = ucb; }
    }
    return node->first_child + best;
}

/* position_iterator.h:24-71 + transformation.h:83-116 */
static void node_features(const ozo_mcts *m, int ni, float *f, int t) {
    ozo_position hist[64];
    int n = 0;
    for (int k = ni; k >= 0 && n < m->H; k = m->nodes[k].parent) hist[n++] = m->nodes[k].pos;
    ozo_positions_to_features(hist, n, f, m->H, t);
}

/* search_thread.cpp:130-190 */
static void expand_and_backward(ozo_mcts *m, int li, int t, const float *policy, const float *value) {
    if (m->nodes[li].pos.player != 0 && m->nodes[li].n_children == 0) { /* :135 */
        int acts[40];
        ozo_position lp = m->nodes[li].pos;
        int n = ozo_legal_actions(&lp, acts);
        int first = (int)m->n_nodes;
        for (int i = 0; i < n; ++i) {
            ozo_position cp;
            ozo_apply_action(&lp, acts[i], &cp);
            new_node(m, &cp, li, policy[ozo_transform_action(acts[i], t)]); /* :139-149 */
        }
        m->nodes[li].first_child = first;
        m->nodes[li].n_children = n;
    }
    node_t *leaf = &m->nodes[li];
    float v;
    if (leaf->pos.player != 0) {
        v = -value[0]; /* :156-157 */
    } else {           /* :158-177 */
        uint64_t pd, od;
        if (m->nodes[leaf->parent].pos.player == 1) { pd = leaf->pos.p1; od = leaf->pos.p2; }
        else { pd = leaf->pos.p2; od = leaf->pos.p1; }
        int pc = __builtin_popcountll(pd), oc = __builtin_popcountll(od);
        v = pc > oc ? 1.0f : (pc < oc ? -1.0f : 0.0f);
    }
    for (int c = li; c != m->root; c = m->nodes[c].parent) { /* :180-189 */
        node_t *nd = &m->nodes[c];
        nd->W += 1.0f + v;
        nd->Q = nd->W / (float)nd->N;
        v = -v;
    }
}

static void run_eval(ozo_mcts *m, int planes) {
    if (m->eval_kind == OZO_EVAL_SYNTH) ozo_eval_synth(m->feat, m->inflight, planes, m->pol, m->val);
    else if (m->eval_kind == OZO_EVAL_HASH) ozo_eval_hash(m->feat, m->inflight, planes, m->pol, m->val);
    else m->eval(m->eval_ctx, m->feat, m->inflight, planes, m->pol, m->val);
}

/* search_thread.cpp:47-128 and mcts.h:220-256.  The T worker threads of the reference
 * are serialised as "thread 0 selects B, thread 1 selects B, ..., thread 0 backs up B,
 * ..." -- one round of T*B slots; for T = 1 this is exactly _simulate_batch. */
void ozo_mcts_search(ozo_mcts *m) {
    int planes = 1 + 2 * m->H, nf = planes * 64;
    int inflight = m->T * m->B;
    if (inflight != m->inflight) {
        m->inflight = inflight;
        m->leaves = (int *)realloc(m->leaves, sizeof(int) * inflight);
        m->tf = (int *)realloc(m->tf, sizeof(int) * inflight);
        m->pol = (float *)realloc(m->pol, sizeof(float) * 65 * inflight);
        m->val = (float *)realloc(m->val, sizeof(float) * inflight);
        free(m->feat);
        m->feat = (float *)calloc((size_t)nf * inflight, sizeof(float)); /* :37-39 zeros */
    }
    memset(m->feat, 0, sizeof(float) * (size_t)nf * inflight); /* new SearchThread per search */
    int rounds = (m->sims + inflight - 1) / inflight;           /* :48-50 */
    for (int r = 0; r < rounds; ++r) {
        for (int i = 0; i < inflight; ++i) { /* :62-79 */
            uint32_t sim = (uint32_t)(r * inflight + i);
            int ni = m->root;
            while (!(m->nodes[ni].pos.player == 0 || m->nodes[ni].n_children == 0))
                ni = choose_best_child(m, ni, sim);
            m->leaves[i] = ni;
            for (int c = ni; c != m->root; c = m->nodes[c].parent) {
                node_t *nd = &m->nodes[c];
                nd->N += 1;
                nd->W -= 1.0f;
                nd->Q = nd->W / (float)nd->N;
            }
            m->nodes[m->root].N += 1;
        }
        int all_terminal = 1;
        for (int i = 0; i < inflight; ++i) { /* :87-100 */
            if (m->nodes[m->leaves[i]].pos.player == 0) continue;
            all_terminal = 0;
            uint32_t sim = (uint32_t)(r * inflight + i);
            if (m->sym_mode == OZO_SYM_IDENTITY) m->tf[i] = 0;
            else { uint32_t w[4]; rng(m, m->search_idx, sim, 0u, w); m->tf[i] = (int)(w[0] & 7u); }
            node_features(m, m->leaves[i], m->feat + (size_t)i * nf, m->tf[i]);
        }
        if (!all_terminal) run_eval(m, planes); /* :102-111 */
        for (int i = 0; i < inflight; ++i)      /* :118-125 */
            expand_and_backward(m, m->leaves[i], m->tf[i], m->pol + i * 65, m->val + i);
    }
    m->search_idx += 1;
}

/* mcts.cpp:45-52 */
int ozo_mcts_visit_counts(const ozo_mcts *m, int *counts) {
    const node_t *r = &m->nodes[m->root];
    for (int i = 0; i < r->n_children; ++i) counts[i] = m->nodes[r->first_child + i].N;
    return r->n_children;
}

/* mcts.cpp:54-61 */
int ozo_mcts_mean_action_values(const ozo_mcts *m, float *values) {
    const node_t *r = &m->nodes[m->root];
    for (int i = 0; i < r->n_children; ++i) values[i] = m->nodes[r->first_child + i].Q;
    return r->n_children;
}

/* mcts.cpp:63-112 */
int ozo_mcts_self_play_data(const ozo_mcts *m, float *features, float *policy) {
    const node_t *r = &m->nodes[m->root];
    if (r->pos.player == 0) return -1;
    if (r->n_children == 0) return -2;
    int acts[40];
    int n = ozo_legal_actions(&r->pos, acts);
    int sum = 0;
    for (int i = 0; i < n; ++i) sum += m->nodes[r->first_child + i].N;
    if (sum == 0) sum = 1;
    int nf = (1 + 2 * m->H) * 64;
    for (int t = 0; t < 8; ++t) {
        node_features(m, m->root, features + (size_t)t * nf, t);
        float *p = policy + t * 65;
        for (int i = 0; i < 65; ++i) p[i] = 0.0f;
        for (int i = 0; i < n; ++i)
            p[ozo_transform_action(acts[i], t)] =
                (float)m->nodes[r->first_child + i].N / (float)sum;
    }
    return 0;
}

/* mcts.cpp:114-165 */
int ozo_mcts_apply_action(ozo_mcts *m, int action) {
    if (!(0 <= action && action < 65)) return -1;
    node_t *r = &m->nodes[m->root];
    if (action == 64) {
        if (r->pos.player == 0 || r->pos.legal != 0) return -2;
    } else if (((1ULL << (63 - action)) & r->pos.legal) == 0) {
        return -2;
    }
    if (r->n_children == 0) { /* :140-148 */
        ozo_position np;
        ozo_apply_action(&r->pos, action, &np);
        m->root = new_node(m, &np, m->root, 1.0f);
        return 0;
    }
    int idx = 0;
    if (action != 0 && r->n_children > 1) { /* :150-157 */
        uint64_t prev = ~0ULL << (64 - action);
        idx = __builtin_popcountll(r->pos.legal & prev);
    }
    m->root = r->first_child + idx; /* siblings become unreachable (:161-164) */
    return 0;
}

int ozo_mcts_choose_action(const ozo_mcts *m, int ply, int temperature_plies) {
    const node_t *r = &m->nodes[m->root];
    int acts[40], n = ozo_legal_actions(&r->pos, acts);
    if (n == 0) return -1;
    uint32_t w[4];
    rng(m, (uint32_t)ply, 0xFFFFFFFFu, 2u, w);
    if (r->n_children == 0) return acts[0];
    if (ply < temperature_plies) {
        uint32_t sum = 0;
        for (int i = 0; i < n; ++i) sum += (uint32_t)m->nodes[r->first_child + i].N;
        if (sum == 0) return acts[(uint32_t)(((uint64_t)w[0] * (uint32_t)n) >> 32)];
        uint32_t pick = (uint32_t)(((uint64_t)w[0] * sum) >> 32), acc = 0;
        for (int i = 0; i < n; ++i) {
            acc += (uint32_t)m->nodes[r->first_child + i].N;
            if (pick < acc) return acts[i];
        }
        return acts[n - 1];
    }
    int best = -1, ties = 0;
    for (int i = 0; i < n; ++i) {
        int c = m->nodes[r->first_child + i].N;
        if (c > best) { best = c; ties = 1; } else if (c == best) ++ties;
    }
    uint32_t pick = (uint32_t)(((uint64_t)w[0] * (uint32_t)ties) >> 32);
    for (int i = 0; i < n; ++i)
        if (m->nodes[r->first_child + i].N == best && pick-- == 0) return acts[i];
    return acts[0];
}

/* ------------------------------------------------------------- evaluators */

/* SURVEY.md section 8(c) "exact synthetic net": D4-equivariant, every operation exactly
 * representable in fp32 except one correctly rounded division per output. */
void ozo_eval_synth(const float *f, int batch, int planes, float *policy, float *value) {
    for (int b = 0; b < batch; ++b) {
        const float *x = f + (size_t)b * planes * 64;
        const float *bl = x + 64, *wh = x + 128;
        float w[65], sum = 0.0f, s1 = 0.0f, s2 = 0.0f;
        for (int r = 0; r < 8; ++r)
            for (int c = 0; c < 8; ++c) {
                float box = 0.0f;
                for (int dr = -1; dr <= 1; ++dr)
                    for (int dc = -1; dc <= 1; ++dc) {
                        int rr = r + dr, cc = c + dc;
                        if (rr >= 0 && rr < 8 && cc >= 0 && cc < 8)
                            box += bl[rr * 8 + cc] + wh[rr * 8 + cc];
                    }
                w[r * 8 + c] = 1.0f + box;
            }
        w[64] = 1.0f;
        for (int i = 0; i < 65; ++i) sum += w[i];
        for (int i = 0; i < 64; ++i) { s1 += bl[i]; s2 += wh[i]; }
        for (int i = 0; i < 65; ++i) policy[b * 65 + i] = w[i] / sum;
        value[b] = (1.0f - 2.0f * x[0]) * (s1 - s2) / 64.0f;
    }
}

static uint64_t mix64(uint64_t z) {
    z ^= z >> 32; z *= 0x9E3779B97F4A7C15ULL; z ^= z >> 29;
    z *= 0xBF58476D1CE4E5B9ULL; z ^= z >> 32;
    return z;
}

/* Pseudo-random priors/values from an integer hash of the feature planes: integer
 * weights, one correctly rounded division per output.  Not D4-equivariant. */
void ozo_eval_hash(const float *f, int batch, int planes, float *policy, float *value) {
    for (int b = 0; b < batch; ++b) {
        const float *x = f + (size_t)b * planes * 64;
        uint64_t h = 0x243F6A8885A308D3ULL;
        for (int c = 0; c < planes; ++c) {
            uint64_t m = 0;
            for (int i = 0; i < 64; ++i) if (x[c * 64 + i] != 0.0f) m |= 1ULL << (63 - i);
            h = mix64(h ^ m);
        }
        uint32_t w[65], sum = 0;
        for (int i = 0; i < 65; ++i) {
            w[i] = 1u + (uint32_t)(mix64(h + (uint64_t)(i + 1) * 0xD1B54A32D192ED03ULL) % 997u);
            sum += w[i];
        }
        for (int i = 0; i < 65; ++i) policy[b * 65 + i] = (float)w[i] / (float)sum;
        int v = (int)(mix64(h ^ 0xA5A5A5A5A5A5A5A5ULL) % 2001u) - 1000;
        value[b] = (float)v / 1000.0f;
    }
}
