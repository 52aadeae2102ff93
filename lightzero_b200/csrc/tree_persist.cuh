// tree_persist.cuh -- the tree phase of the persistent search kernel (net_tc.cu): the same arithmetic as tree.cuh's
// tree_traverse<false,false> / tree_backprop<false,false> (cnode.cpp:419-449, 480-500, 551-595, 654-698, 754-825), restated for a
// warp that owns ONE tree for the whole search, so that the dependent global round trips shrink to one per tree level:
//   * per-tree scalars (path length, virtual to_play, MinMax bounds, the root's visit count / value sum, the root's legal list)
//     and the first 32 path entries live in the owning warp's registers across simulations (the global copies are still
//     written, fire-and-forget, so the step-wise entry points and lz_tree_results see the same state afterwards);
//   * one level of the descent = ONE batch of five coalesced loads (the node's whole edge block: every child's prior, value sum,
//     reward, visit count, child slot); compute_mean_q, the PUCT scores, the arg-max and the chosen child's (visit, slot) all
//     come out of those registers by shuffles;
//   * the exploration-rate table pbc[] sits in shared memory.
// Every fp32 operation is the same explicit round-to-nearest intrinsic in the same order as tree.cuh (the bit-exact parity
// tests drive both and compare the trees word for word).  A <= 32 only (one lane per child); larger action spaces use tree.cuh.
#pragma once
#include "tree.cuh"

namespace lz {

struct PTree {                 // registers of the owning warp (uniform across lanes unless noted)
    int nl, my_legal;          // root legal list; lane k holds legal[k] (per lane)
    int plen, vtp;
    float mmax, mmin;
    int root_visit;
    float root_vsum, root_reward;
    int root_to_play, tp0, players;
    int my_pslot, my_pact;     // path entry `lane` (per lane); entries >= 32 only in global memory
};

__device__ __forceinline__ void ptree_init(const TreeParams &p, PTree &T, int b, int lane)
{
    T.nl = p.nlegal[b];
    T.my_legal = (lane < T.nl && lane < p.A) ? p.legal[(size_t)b * p.A + lane] : 0;
    T.plen = p.path_len[b];
    T.vtp = p.vtp[b];
    T.mmax = p.mm_max[b];
    T.mmin = p.mm_min[b];
    T.root_visit = p.root_visit[b];
    T.root_vsum = p.root_vsum[b];
    T.root_reward = p.root_reward[b];
    T.root_to_play = p.n_to_play[(size_t)b * p.N];
    T.tp0 = p.to_play[b];
    T.players = (*p.players_max == -1) ? 1 : 2;
    T.my_pslot = 0;
    T.my_pact = 0;
}

// cucb_score (cnode.cpp:654-698) on register operands; identical operation order to tree.cuh's ucb_score<false>
__device__ __forceinline__ float ptree_ucb(bool active, int vis, float prior, float rw, float vsum, float pbc, float sq, float mean_q,
                                           float discount, int players, float mmax, float mmin, float delta_max)
{
    if (!active) return -INFINITY;
    float pb_c = __fmul_rn(pbc, __fdiv_rn(sq, (float)(vis + 1)));
    float prior_score = __fmul_rn(pb_c, prior);
    float value_score;
    if (vis == 0) {
        value_score = mean_q;
    } else {
        float v = __fdiv_rn(vsum, (float)vis);
        value_score = __fadd_rn(rw, __fmul_rn(discount, players == 1 ? v : -v));
    }
    value_score = mm_normalize(value_score, mmax, mmin, delta_max);
    if (value_score < 0.0f) value_score = 0.0f;
    if (value_score > 1.0f) value_score = 1.0f;
    return __fadd_rn(prior_score, value_score);
}

// One PUCT descent (cbatch_traverse body, cnode.cpp:783-824).  pbc_tab: the table in shared memory (or p.pbc).
__device__ __forceinline__ void ptree_traverse(const TreeParams &p, PTree &T, int b, int lane, int deterministic, unsigned step,
                                               const float *pbc_tab, int *out_ix, int *out_action, int *leaf_slot = nullptr,
                                               int *leaf_action = nullptr)
{
    const int A = p.A, N = p.N;
    const uint32_t *tree_edges = p.edges + (size_t)b * N * kEdgeFields * A;
    const int players = T.players;
    const float discount = p.discount, delta_max = p.delta;
    const float mmax = T.mmax, mmin = T.mmin;
    int *pslot = p.path_slot + (size_t)b * N, *pact = p.path_action + (size_t)b * N;

    int slot = 0, node_visit = T.root_visit, plen = 0, last_action = -1;
    int vtp = T.tp0;
    bool is_root = true;
    float parent_q = 0.0f;

    while (true) {
        const uint32_t *nb = tree_edges + (size_t)slot * kEdgeFields * A;
        const int n = is_root ? T.nl : A;
        const bool act = lane < n;
        const int a = act ? (is_root ? T.my_legal : lane) : 0;
        // the node's whole edge block in one round trip
        uint32_t w_vis = 0, w_vsum = 0, w_rew = 0, w_prior = 0, w_cs = 0xffffffffu;
        if (act) {
            w_vis = nb[F_VISIT * A + a];
            w_vsum = nb[F_VSUM * A + a];
            w_rew = nb[F_REWARD * A + a];
            w_prior = nb[F_PRIOR * A + a];
            w_cs = nb[F_CSLOT * A + a];
        }
        const int vis = (int)w_vis;
        const float vsum = u2f(w_vsum), rw = u2f(w_rew), prior = u2f(w_prior);
        // ---- compute_mean_q: sequential fp32 sum over visited children in legal order (cnode.cpp:169-203)
        float q = 0.0f;
        if (vis > 0) q = __fadd_rn(rw, __fmul_rn(discount, __fdiv_rn(vsum, (float)vis)));
        // every lane runs the same sum; the broadcasts do not depend on the running total, so they pipeline ahead of the add chain
        float total = 0.0f;
        const unsigned m = __ballot_sync(0xffffffffu, vis > 0);
        const int tv = __popc(m);
#pragma unroll 6
        for (int l = 0; l < n; ++l) {
            const float ql = __shfl_sync(0xffffffffu, q, l);
            if ((m >> l) & 1u) total = __fadd_rn(total, ql);
        }
        float mean_q;
        if (is_root && tv > 0) mean_q = __fdiv_rn(total, (float)tv);
        else mean_q = __fdiv_rn(__fadd_rn(parent_q, total), (float)(tv + 1));

        // ---- cselect_child: first legal position attaining the exact maximum (cnode.cpp:551-595)
        const float total_children = (float)(node_visit - 1);   // cnode.cpp:574
        const float pbc = pbc_tab[node_visit - 1];
        const float sq = __fsqrt_rn(total_children);
        const float sc = ptree_ucb(act, vis, prior, rw, vsum, pbc, sq, mean_q, discount, players, mmax, mmin, delta_max);
        float best = kFloatMin;
        int best_k = -1;
        {
            const float cmax = warp_max_exact(sc);
            const unsigned eq = __ballot_sync(0xffffffffu, act && sc == cmax);
            if (best < cmax) {
                best = cmax;
                best_k = __ffs(eq) - 1;
            }
        }
        if (!deterministic && best_k >= 0) {
            // tie list of cnode.cpp:576-586: the arg-max position, then every LATER position whose score >= max - 1e-6;
            // drawn uniformly with the same counter-based hash as tree.cuh
            const float thr = __fsub_rn(best, 0.000001f);
            const unsigned later = __ballot_sync(0xffffffffu, act && lane > best_k && sc >= thr);
            const int count = 1 + __popc(later);
            if (count > 1) {
                unsigned long long h = mix64(p.rng_seed ^ mix64(*p.rng_epoch) ^ mix64(((unsigned long long)b << 32) | step) ^ (unsigned)plen);
                const int r = (int)(h % (unsigned)count);
                if (r > 0) {
                    unsigned mm2 = later;
                    for (int j = 1; j < r; ++j) mm2 &= mm2 - 1;
                    best_k = __ffs(mm2) - 1;
                }
            }
        }
        int action = 0;
        {
            const int la = __shfl_sync(0xffffffffu, T.my_legal, best_k >= 0 ? best_k : 0);
            if (best_k >= 0) action = is_root ? la : best_k;
        }
        if (players > 1) vtp = (vtp == 1) ? 2 : 1;   // cnode.cpp:798-805

        if (lane == 0) {
            p.n_best[(size_t)b * N + slot] = action;   // cnode.cpp:807
            pslot[plen] = slot;
            pact[plen] = action;
        }
        if (lane == plen) { T.my_pslot = slot; T.my_pact = action; }
        ++plen;
        last_action = action;
        int cs;
        if (best_k >= 0) {           // the chosen child's statistics are in lane best_k's registers
            node_visit = __shfl_sync(0xffffffffu, vis, best_k);
            cs = (int)__shfl_sync(0xffffffffu, w_cs, best_k);
        } else {                     // no child scored above FLOAT_MIN: action 0 (cnode.cpp:589), whatever its block holds
            node_visit = (int)nb[F_VISIT * A];
            cs = (int)nb[F_CSLOT * A];
        }
        is_root = false;
        parent_q = mean_q;
        if (cs < 0 || plen >= N) break;
        slot = cs;
    }
    T.plen = plen;
    T.vtp = vtp;
    if (lane == 0) {
        p.path_len[b] = plen;
        p.vtp[b] = vtp;
        p.search_len[b] = plen;
        if (out_ix) out_ix[b] = slot;     // parent of the leaf: its slot == current_latent_state_index
        if (out_action) out_action[b] = last_action;
        if (leaf_slot) *leaf_slot = slot;             // e.g. a shared-memory hand-off to the network phase of the same CTA
        if (leaf_action) *leaf_action = last_action;
    }
    __syncwarp();
}

// path entry i (uniform call; i may differ per lane): registers for i < 32, global beyond
__device__ __forceinline__ int ptree_path(int reg, const int *glob, int i)
{
    const int r = __shfl_sync(0xffffffffu, reg, i & 31);
    return (i >= 32) ? glob[i] : r;
}

// cbatch_backpropagate body (cnode.cpp:495-499): expand the leaf into slot `latent_index`, then cbackpropagate (cnode.cpp:419-478)
__device__ __forceinline__ void ptree_backprop(const TreeParams &p, PTree &T, int b, int lane, int latent_index, float reward, float value,
                                               const float *logits)
{
    const int A = p.A, N = p.N;
    const int plen = T.plen;
    if (plen == 0 || latent_index >= N) return;
    uint32_t *tree_edges = p.edges + (size_t)b * N * kEdgeFields * A;
    const int *pslot = p.path_slot + (size_t)b * N, *pact = p.path_action + (size_t)b * N;
    const int tp = T.vtp;
    const float discount = p.discount;

    expand_block(tree_edges + (size_t)latent_index * kEdgeFields * A, A, logits, nullptr, A, lane);
    const int leaf_ps = ptree_path(T.my_pslot, pslot, plen - 1), leaf_pa = ptree_path(T.my_pact, pact, plen - 1);
    uint32_t *leaf_nb = tree_edges + (size_t)leaf_ps * kEdgeFields * A;
    if (lane == 0) {
        p.n_batch[(size_t)b * N + latent_index] = b;
        p.n_to_play[(size_t)b * N + latent_index] = tp;
        p.n_best[(size_t)b * N + latent_index] = -1;
        leaf_nb[F_CSLOT * A + leaf_pa] = (uint32_t)latent_index;
        leaf_nb[F_REWARD * A + leaf_pa] = f2u(reward);
    }

    float mmax = T.mmax, mmin = T.mmin;
    float G = value;   // bootstrap_value
    // path nodes i = plen (leaf) ... 0 (root); node i>=1 hangs on edge (pslot[i-1], pact[i-1]).
    for (int hi = plen; hi >= 0; hi -= 32) {
        const int i = hi - lane;
        const bool act = i >= 0;
        const int ps_prev = ptree_path(T.my_pslot, pslot, act && i >= 1 ? i - 1 : 0);
        const int pa_prev = ptree_path(T.my_pact, pact, act && i >= 1 ? i - 1 : 0);
        const int ps_self = ptree_path(T.my_pslot, pslot, act && i < plen ? i : 0);
        float vs = 0.0f, rw = 0.0f;
        int vc = 0, ntp = 0;
        uint32_t *enb = nullptr;
        int ea = 0;
        if (act) {
            if (i == plen) {               // the leaf: unvisited edge, reward just predicted
                rw = reward; ntp = tp;
                enb = leaf_nb; ea = leaf_pa;
            } else if (i == 0) {
                vs = T.root_vsum; vc = T.root_visit; rw = T.root_reward;
                ntp = T.root_to_play;
            } else {
                enb = tree_edges + (size_t)ps_prev * kEdgeFields * A;
                ea = pa_prev;
                vs = u2f(enb[F_VSUM * A + ea]);
                vc = (int)enb[F_VISIT * A + ea];
                rw = u2f(enb[F_REWARD * A + ea]);
                if (tp != -1) ntp = p.n_to_play[(size_t)b * N + ps_self];    // only the two-player recurrence reads it
            }
        }
        const int cnt = min(32, hi + 1);
        float my_vs = vs;
        int my_vc = vc;
        for (int l = 0; l < cnt; ++l) {
            float vs_l = __shfl_sync(0xffffffffu, vs, l);
            int vc_l = __shfl_sync(0xffffffffu, vc, l);
            float rw_l = __shfl_sync(0xffffffffu, rw, l);
            int ntp_l = __shfl_sync(0xffffffffu, ntp, l);
            float nvs, upd;
            if (tp == -1) {                // cnode.cpp:432-449
                nvs = __fadd_rn(vs_l, G);
                float nodeval = __fdiv_rn(nvs, (float)(vc_l + 1));
                upd = __fadd_rn(rw_l, __fmul_rn(discount, nodeval));
                G = __fadd_rn(rw_l, __fmul_rn(discount, G));
            } else {                       // cnode.cpp:450-477
                bool same = (ntp_l == tp);
                nvs = __fadd_rn(vs_l, same ? G : -G);
                float nodeval = __fdiv_rn(nvs, (float)(vc_l + 1));
                upd = __fadd_rn(rw_l, __fmul_rn(discount, -nodeval));
                G = same ? __fadd_rn(-rw_l, __fmul_rn(discount, G)) : __fadd_rn(rw_l, __fmul_rn(discount, G));
            }
            if (upd > mmax) mmax = upd;    // cminimax.cpp:19-26
            if (upd < mmin) mmin = upd;
            if (lane == l) { my_vs = nvs; my_vc = vc_l + 1; }
        }
        if (act) {
            if (i == 0) {
                p.root_vsum[b] = my_vs;
                p.root_visit[b] = my_vc;
            } else {
                enb[F_VSUM * A + ea] = f2u(my_vs);
                enb[F_VISIT * A + ea] = (uint32_t)my_vc;
            }
        }
        if (hi < 32) {                     // the chunk that holds the root (i == 0 in lane hi)
            T.root_vsum = __shfl_sync(0xffffffffu, my_vs, hi);
            T.root_visit = __shfl_sync(0xffffffffu, my_vc, hi);
        }
    }
    T.mmax = mmax;
    T.mmin = mmin;
    if (lane == 0) {
        p.mm_max[b] = mmax;
        p.mm_min[b] = mmin;
    }
    __syncwarp();
}

}  // namespace lz
