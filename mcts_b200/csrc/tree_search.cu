// tree_search.cu -- batched leaf-gather MCTS driver over the GPU rollout engine (host code only).
//
// What it mirrors in the reference (thomasmarsh/mcts):
//   descent ............ select_step, search/shared.rs:569-736 (expand :220-289, one virtual loss per edge :665-671)
//   child choice ....... UCB1 / UCB1-Tuned, select/ucb.rs:9-146, unvisited value node.rs:139-159,
//                        arg-max with random offset + prime stride, strict '>', select/mod.rs:343-386
//   leaf fan-out ....... num_rollouts_per_leaf with k-1 extra virtual loss, search_impl.rs:101-113, shared.rs:787-805
//   backprop ........... backprop.rs:704-731 with ChildArray::update / NodeStats::update (node.rs:295-302,673-681)
//                        summed over the k trials of a leaf (exact: utilities are +1 / 0 / -1)
//   GRAVE / RAVE ....... select/rave.rs:60-75 (beta schedules HandSelected, MinMSE, Threshold), :157-184 (get_ref),
//                        :196-230 (score); tables written by update_grave, backprop.rs:619-640
//   AMAF ............... select/amaf.rs:48-86 over per-edge AMAF rows written by update_amaf, backprop.rs:565-617
//   LGR table .......... TreeStats.player_replies, last write wins over (trial, ply) order (backprop.rs:908-923): the
//                        kernel's stamps for the playout part, the tree-path pairs added here
//   LGRF-2 table ....... TreeStats.player_replies2 with forgetting (backprop.rs:926-966): the kernel's last insert per
//                        context and its replay pass (mr_reply2_inserts / mr_reply2_resolve), merged here with the
//                        tree-path windows of every leaf (inserted by its trials the player did not lose, removed by
//                        the ones it lost)
//   NST table .......... TreeStats.player_bigram_actions over consecutive (preceding action, action) pairs of tree path
//                        and playout (backprop.rs:885-899): the kernel's bigram delta plus the tree-path pairs added here
//   MAST table ......... TreeStats.player_actions, written for playout trace AND tree path (backprop.rs:833-870)
//   final move ......... RobustChild: (visits, expected score) lexicographic max, select/basic.rs:13-45
//   root parallel ...... choose_action_root_parallel, search/parallel.rs:56-115: one searcher per host thread, each with
//                        its own seed, tree and ctx (streams); root-child totals summed, most merged visits wins
// The tree (arena, edge statistics in the parent's child arrays, root_stats) stays on the host; only the
// simulation phase crosses the C ABI (mr_submit / mr_wait).  Tie-break draws of the tree come from the same
// Philox4x32-10 generator as the playouts, on their own stream: ctr = (descent id, depth, 0, 2).
//
// Two things are new relative to the reference's loop, both only visible with batch_leaves > 1 (with one leaf per batch
// and one batch in flight the driver IS the sequential loop of search_impl.rs:76-114):
//   * a batch of B descents is gathered before any of their results exist, and up to `pipeline_depth` batches are in
//     flight (the staleness the reference's tree-parallel workers live with, parallel.rs:188-251, made deterministic
//     by the fixed order gather(n), submit(n), apply(n - depth + 1));
//   * the expansion test counts IN-FLIGHT trials: a node is simulated from when `num_visits + trials in flight through
//     it < expand_threshold`, so the second descent that reaches a node whose first simulation is still on the GPU
//     expands it and goes on instead of queueing the same leaf again.  Without this every descent of the first batch
//     stops at the unvisited root, and a fresh node swallows as many descents as arrive before its first result.
#include <chrono>
#include <cmath>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "host_games.h"
#include "mcts_rollout.h"

namespace {

const uint32_t kPrimes[16] = {14323, 18713, 19463, 30553, 33469, 45343, 50221, 51991,
                              53201, 56923, 64891, 72763, 74471, 81647, 92581, 94693};

void philox4(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}
uint32_t philox_word0(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    uint32_t o[4];
    philox4(seed, c0, c1, c2, c3, o);
    return o[0];
}
inline uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

// NodeStats / one ChildArray row.  Utilities are +1 / -1 / 0 in a two-player zero-sum game, so score[1] == -score[0]
// and sum_squared_score is the number of decisive trials for both players: one number each.
struct Stats {
    uint32_t visits = 0, vloss = 0;
    int64_t score0 = 0, sumsq = 0;
    int64_t score(int player) const { return player == 0 ? score0 : -score0; }
};
struct AmafRow { // ChildArray amaf rows (node.rs:392-396,711-719): every player's count moves together
    uint32_t visits = 0;
    int64_t score0 = 0;
};
// node.rs:126-133
inline double expected_score(const Stats &s, int player) {
    if (s.visits == 0) return 0.0;
    const double lv = (double)s.vloss;
    return ((double)s.score(player) - lv) / ((double)s.visits + lv);
}

// select/rave.rs:60-75
inline double rave_beta(const mr_search_desc *d, uint32_t n, uint32_t amaf_n) {
    if (d->rave_schedule == MR_RAVE_THRESHOLD) {
        const double rave = (double)d->rave_param;
        return std::max(0.0, rave - (double)n) / rave;
    }
    if (d->rave_schedule == MR_RAVE_MIN_MSE) {
        const double nn = (double)n, an = (double)amaf_n, bias = d->rave_bias;
        return an / (nn + an + 4.0 * nn * an * bias * bias);
    }
    const double k = (double)d->rave_param;
    return std::sqrt(k / (3.0 * (double)n + k));
}

// The value an unvisited (not yet created) child competes with: node.rs:139-159 value_estimate_unvisited, then the
// strategy's own exploration term (ucb.rs:55-61, :134-146; none for Rave / Amaf, rave.rs:226-230, amaf.rs:76-79).
inline double unvisited_child_score(const mr_search_desc *d, const Stats &cur, int player, double parent_log) {
    double q;
    if (d->q_init == MR_QINIT_INFINITY) q = 10000.0;
    else q = cur.visits == 0 ? 10000.0 : expected_score(cur, player);
    const double c = d->exploration_c;
    if (d->select == MR_SELECT_UCB1_TUNED) return q + (parent_log * std::min(1.0, 1.0) + c * std::sqrt(parent_log));
    if (d->select == MR_SELECT_GRAVE || d->select == MR_SELECT_AMAF) return q;
    return q + c * std::sqrt(parent_log);
}

// The score of a created child: ucb.rs:40-52 (UCB1), :90-132 (UCB1-Tuned), rave.rs:196-224 (GRAVE; grave_n / grave_q are
// the reference node's AMAF count and score sum of this action for the mover), amaf.rs:58-74 (AMAF).
inline double created_child_score(const mr_search_desc *d, const Stats &e, const AmafRow *amaf_row, int player, double parent_log,
                                  uint32_t grave_n, double grave_q) {
    const double c = d->exploration_c;
    const double exploit = expected_score(e, player);
    const double total = (double)(e.visits + e.vloss);
    switch (d->select) {
    case MR_SELECT_UCB1_TUNED: {
        const double var = std::max(0.0, (double)e.sumsq / total - exploit * exploit);
        const double frac = parent_log / total;
        return exploit + (frac * std::min(1.0, var) + c * std::sqrt(frac));
    }
    case MR_SELECT_GRAVE: {
        const uint32_t n_tot = e.visits + e.vloss;
        const double explore = c * std::sqrt(parent_log / (double)n_tot);
        const double b = rave_beta(d, n_tot, grave_n);
        const double amaf = grave_n == 0 ? 0.0 : grave_q / (double)grave_n;
        return (1.0 - b) * exploit + b * amaf + explore;
    }
    case MR_SELECT_AMAF: {
        const double amaf_n = (double)std::max<uint32_t>(1u, amaf_row->visits);
        const double amaf = (double)(player == 0 ? amaf_row->score0 : -amaf_row->score0) / amaf_n;
        const double ucb1 = exploit + c * std::sqrt(parent_log / total);
        return d->amaf_alpha * amaf + (1.0 - d->amaf_alpha) * ucb1;
    }
    default: return exploit + c * std::sqrt(parent_log / total);
    }
}

enum : uint8_t { UNEXPANDED = 0, EXPANDED = 1, TERMINAL = 2 };

struct Node {
    mr_leaf state;
    uint8_t status = UNEXPANDED;
    uint8_t player = 0;
    uint16_t n_children = 0;
    uint32_t first = 0;     // index of child 0 in the edge arrays
    int32_t grave_tab = -1; // cached index of this state's GRAVE table (-1 = not looked up successfully yet)
};

struct PathEntry {
    uint32_t node;
    uint32_t idx; // this node's slot in its parent's child array (unused for the root)
};

struct StateKey {
    uint64_t w[5]; // the 40-byte record
    bool operator==(const StateKey &o) const { return memcmp(w, o.w, sizeof w) == 0; }
};
struct StateKeyHash {
    size_t operator()(const StateKey &k) const {
        uint64_t h = 0xcbf29ce484222325ull;
        for (int i = 0; i < 5; ++i) {
            h ^= k.w[i];
            h *= 0x100000001b3ull;
            h ^= h >> 29;
        }
        return (size_t)h;
    }
};
inline StateKey key_of(const mr_leaf &s) {
    StateKey k;
    memcpy(k.w, &s, sizeof k.w);
    return k;
}

struct Tree {
    int game;
    bool use_table = false;
    bool symmetry = false; // MR_SEARCH_SYMMETRY: the table is keyed by Othello's canonical orientation (see Search::select)
    bool want_amaf_rows = false;
    std::unordered_map<StateKey, uint32_t, StateKeyHash> table; // use_transpositions: state -> node
    std::vector<Node> nodes;
    // the edges, as parallel arrays (the selection loop reads child + edge, the rest is touched on expansion / backprop)
    std::vector<uint16_t> action;
    std::vector<uint16_t> slot; // compact per-player slot of the action, in the numbering of the node's mover
    std::vector<int32_t> child; // -1 = not created yet
    std::vector<Stats> edge;
    std::vector<AmafRow> amaf_row; // MR_SELECT_AMAF only
    Stats root_stats;
    uint64_t root_inflight = 0;    // trials gathered through the root and not applied yet (the root has no edge to carry
                                   // virtual losses; root_stats stays as the reference keeps it)
    uint32_t max_depth = 0;

    uint32_t add_node(const mr_leaf &s) {
        Node n;
        n.state = s;
        n.player = s.turn;
        nodes.push_back(n);
        return (uint32_t)nodes.size() - 1;
    }
    // search/shared.rs:220-289: terminal_status first; otherwise children in generate_actions order
    void expand(uint32_t id) {
        Node &n = nodes[id];
        if (mr_host::terminal_status(game, n.state) != MR_NOT_TERMINAL) {
            n.status = TERMINAL;
            return;
        }
        uint16_t acts[MR_MAX_ROOT_CHILDREN];
        const int cnt = mr_host::generate_actions(game, n.state, acts);
        if (cnt == 0) { // cannot be expanded further; the playout loop scores it as a draw (simulate.rs:112-116)
            n.status = TERMINAL;
            return;
        }
        n.first = (uint32_t)action.size();
        n.n_children = (uint16_t)cnt;
        for (int i = 0; i < cnt; ++i) {
            action.push_back(acts[i]);
            slot.push_back((uint16_t)mr_host::action_slot(game, n.player, acts[i]));
            child.push_back(-1);
            edge.push_back(Stats());
            if (want_amaf_rows) amaf_row.push_back(AmafRow());
        }
        n.status = EXPANDED;
    }
};

struct GraveTable { // per state: [2][S] ActionStats (node.rs:51-55) over compact action slots, integer scores
    std::vector<uint32_t> visits;
    std::vector<int64_t> score;
    explicit GraveTable(size_t n) : visits(n, 0u), score(n, 0) {}
};

struct Search {
    mr_ctx *ctx;
    const mr_search_desc *d;
    Tree t;
    std::vector<mr_action_stat> mast;   // [2][NA], TreeStats.player_actions
    std::vector<mr_action_stat> bigram; // [2][S][S], TreeStats.player_bigram_actions over compact action slots (MR_EPS_NST)
    std::vector<uint16_t> replies2;     // [2][S][S], TreeStats.player_replies2 over action slots: 1 + reply code (MR_EPS_LGR2)
    std::vector<uint16_t> replies;      // [2][NA], TreeStats.player_replies: 1 + reply code by (player, preceding action id)
    int na;
    int S; // compact action slots per player
    // TreeStats.grave (search/shared.rs:52): per reference-node state, per player, per action.  The reference keys
    // it by node.hash; here the key is the full state record (no collisions), so equal states share one table.
    std::unordered_map<StateKey, uint32_t, StateKeyHash> grave_index;
    std::vector<GraveTable> grave; // one per distinct state

    // (the table of a node is found once through the state-keyed index and then remembered in the node: the selection
    // formula asks for it once per CHILD, 121 times per visited Hex node)
    GraveTable *grave_of(Node &n, bool create) {
        if (n.grave_tab >= 0) return &grave[(size_t)n.grave_tab];
        auto it = grave_index.find(key_of(n.state));
        if (it != grave_index.end()) {
            n.grave_tab = (int32_t)it->second;
            return &grave[it->second];
        }
        if (!create) return nullptr;
        n.grave_tab = (int32_t)grave.size();
        grave_index.emplace(key_of(n.state), (uint32_t)grave.size());
        grave.emplace_back((size_t)2 * S);
        return &grave.back();
    }

    // select/rave.rs:157-184 get_ref: the child itself when its incoming edge has total_visits >= threshold (decided in
    // best_child), else the nearest such node on the path, else the root
    uint32_t grave_ref_ancestor(const std::vector<PathEntry> &path, size_t len) const {
        for (size_t i = len; i-- > 1;) {
            const Stats &e = t.edge[t.nodes[path[i - 1].node].first + path[i].idx];
            if (e.visits + e.vloss >= d->grave_threshold) return path[i].node;
        }
        return 0;
    }

    const Stats &current_stats(const std::vector<PathEntry> &path) const {
        if (path.size() == 1) return t.root_stats;
        const PathEntry &me = path.back();
        const Node &parent = t.nodes[path[path.size() - 2].node];
        return t.edge[parent.first + me.idx];
    }

    // ---- per-edge score cache ----
    // Between two backprops the only statistics that change are the virtual losses of the edges a descent walks, so the
    // score of every OTHER child of a node stays what it was when the node was last evaluated.  A node's children are
    // scored in full on its first visit of an epoch (epoch = the number of backprop batches applied, bumped too whenever
    // an edge crosses the GRAVE threshold, which changes reference nodes below it) and the walked edge is re-scored when
    // its virtual loss changes; the arg-max then reads cached numbers.  The numbers are the ones a fresh evaluation
    // would produce (same inputs, same formula), so the search is unchanged -- the oracle re-evaluates everything on
    // every descent and the parity tests compare.  Off with transpositions: a shared node is reached along different
    // edges, and the "current" statistics its scores depend on are the incoming edge's.
    struct LevelCtx { // what scoring a child of one path node needs, kept per level of the current descent
        double parent_log = 0.0;
        int player = 0;
        bool anc_done = false;
        const GraveTable *anc_tab = nullptr;
    };
    std::vector<double> score_cache;  // per edge
    std::vector<uint32_t> node_epoch; // per node: the epoch its children's cached scores belong to (0 = never)
    uint32_t epoch = 1;
    bool cache_ok = false;
    std::vector<LevelCtx> lctx; // parallel to the path of the descent in progress

    double edge_score(const std::vector<PathEntry> &path, size_t level, uint32_t ei, LevelCtx &c) {
        const Stats &e = t.edge[ei];
        uint32_t grave_n = 0;
        double grave_q = 0.0;
        if (d->select == MR_SELECT_GRAVE) {
            // get_ref: the child itself once its edge has the threshold visits, else the nearest such ancestor --
            // which is the same node for every child of this node, so it is resolved once per visit (anc_tab)
            const bool own_ref = e.visits + e.vloss >= d->grave_threshold;
            if (!own_ref && !c.anc_done) {
                c.anc_tab = grave_of(t.nodes[grave_ref_ancestor(path, level + 1)], false);
                c.anc_done = true;
            }
            if (const GraveTable *tab = own_ref ? grave_of(t.nodes[(uint32_t)t.child[ei]], false) : c.anc_tab) {
                const size_t g = (size_t)c.player * S + t.slot[ei];
                grave_n = tab->visits[g];
                grave_q = (double)tab->score[g];
            }
        }
        return created_child_score(d, e, t.want_amaf_rows ? &t.amaf_row[ei] : nullptr, c.player, c.parent_log, grave_n, grave_q);
    }

    // an edge's virtual loss changes from `before` to its current value: keep the cache exact
    void edge_changed(const std::vector<PathEntry> &path, size_t level, uint32_t ei, uint32_t total_before) {
        if (!cache_ok) return;
        const Stats &e = t.edge[ei];
        // crossing the GRAVE threshold makes the child the reference node of everything BELOW it: cached scores in its
        // subtree are stale (a child that has not been expanded has no subtree -- the common case, a fresh leaf whose
        // first k trials already reach the threshold)
        if (d->select == MR_SELECT_GRAVE && total_before < d->grave_threshold && e.visits + e.vloss >= d->grave_threshold &&
            t.child[ei] >= 0 && t.nodes[(uint32_t)t.child[ei]].status == EXPANDED)
            ++epoch;
        if (t.child[ei] >= 0) score_cache[ei] = edge_score(path, level, ei, lctx[level]);
    }

    // select/mod.rs:343-386 with the strategies' scores; `path.back()` is the node, `lctx.back()` its context
    uint32_t best_child(const std::vector<PathEntry> &path, uint32_t descent) {
        const uint32_t node_id = path.back().node;
        const Node &node = t.nodes[node_id];
        const size_t level = path.size() - 1;
        LevelCtx &c = lctx[level];
        const int player = node.player;
        const Stats &cur = current_stats(path);
        const double parent_log = std::log(std::max(1.0, (double)cur.visits)); // ucb.rs:34-37: without virtual loss
        c.parent_log = parent_log;
        c.player = player;
        c.anc_done = false;
        c.anc_tab = nullptr;
        const double unvisited_value = unvisited_child_score(d, cur, player, parent_log);
        const uint32_t n = node.n_children;
        const uint32_t r = mulhi32(philox_word0(d->seed, descent, (uint32_t)path.size() - 1, 0, 2), 16u * n);
        uint32_t i = r / 16u;
        const uint32_t stride = kPrimes[r % 16u] % n;
        // No created child can score above 1 + ln N + c sqrt(ln N): exploit, AMAF mean and variance bound are <= 1, every
        // exploration term is at most its value at n_total = 1 (for Amaf with alpha in [0, 1] the blend of two such
        // numbers).  When the unvisited value is beyond that (QInit::Infinity, or a parent without visits), the arg-max
        // is the first child in visiting order that does not exist yet -- found without evaluating a single formula.
        // Exactly the result of the full loop below (strict '>' keeps the first maximum), only cheaper.
        if (unvisited_value > 2.0 + parent_log + std::fabs(d->exploration_c) * std::sqrt(parent_log) &&
            (d->select != MR_SELECT_AMAF || (d->amaf_alpha >= 0.0 && d->amaf_alpha <= 1.0))) {
            uint32_t j = i;
            for (uint32_t k = 0; k < n; ++k) {
                if (t.child[node.first + j] < 0) return j;
                j += stride;
                if (j >= n) j -= n;
            }
        }
        if (cache_ok && node_epoch[node_id] != epoch) { // first visit of this epoch: score every existing child once
            for (uint32_t j = 0; j < n; ++j)
                if (t.child[node.first + j] >= 0) score_cache[node.first + j] = edge_score(path, level, node.first + j, c);
            node_epoch[node_id] = epoch;
        }
        bool have = false;
        double best = 0.0;
        uint32_t best_i = i;
        for (uint32_t k = 0; k < n; ++k) {
            const uint32_t ei = node.first + i;
            const double score = t.child[ei] < 0 ? unvisited_value : (cache_ok ? score_cache[ei] : edge_score(path, level, ei, c));
            if (!have || score > best) {
                have = true;
                best = score;
                best_i = i;
            }
            i += stride;
            if (i >= n) i -= n;
        }
        return best_i;
    }

    // search/shared.rs:569-736, with the in-flight trials counted by the expansion test (file header).  `leaf` receives
    // the state the playouts start from.  With MR_SEARCH_SYMMETRY the descent carries the REAL state: every node but the
    // root holds its state, and so its action list, in the canonical orientation (shared.rs:233-252), and a chosen action is
    // translated back through the symmetry of the real state at that node (symmetry.rs:38-50 incoming_sym, othello/lib.rs:
    // 658-665 invert_action) -- recomputed on every visit, because a shared node is reached in different orientations.
    void select(uint32_t descent, std::vector<PathEntry> &path, mr_leaf &leaf) {
        path.clear();
        lctx.clear();
        uint32_t cur = 0, incoming = 0;
        mr_leaf real = t.nodes[0].state;
        int real_sym = 0;
        for (;;) {
            path.push_back({cur, incoming});
            lctx.emplace_back();
            leaf = t.symmetry ? real : t.nodes[cur].state; // (for every return below)
            const Stats &cs = current_stats(path);
            // trials in flight through this node from OTHER descents: every gathered descent holds k virtual losses on
            // each of its edges (root_inflight for the root); this one has added 1 to the edge it just came down
            const uint64_t others = path.size() == 1 ? t.root_inflight : (uint64_t)cs.vloss - 1u;
            if (t.nodes[cur].status == TERMINAL) return;
            if ((uint64_t)cs.visits + others < d->expand_threshold) return;
            if (t.nodes[cur].status == UNEXPANDED) {
                t.expand(cur);
                if (t.nodes[cur].status == TERMINAL) return;
                score_cache.resize(t.edge.size(), 0.0);
            }
            const uint32_t best = best_child(path, descent);
            incoming = best;
            const uint32_t ei = t.nodes[cur].first + best;
            const uint32_t total_before = t.edge[ei].visits + t.edge[ei].vloss;
            t.edge[ei].vloss += 1; // one virtual loss per traversed edge
            mr_leaf key; // MR_SEARCH_SYMMETRY: what the child node is keyed by
            if (t.symmetry) {
                uint16_t a = t.action[ei];
                // the root's list is literal; PASS maps to itself; real_sym = the symmetry of `real` at this node
                if (cur != 0 && a < 64) a = (uint16_t)mr_host::invert_symmetry8((int)a, real_sym);
                mr_host::apply(t.game, real, a);
                real_sym = mr_host::othello_canonical(real, key);
            }
            if (t.child[ei] >= 0) {
                edge_changed(path, path.size() - 1, ei, total_before);
                cur = (uint32_t)t.child[ei];
                // shared.rs:382-405 verified_child_id: the cached slot belongs to the orientation that created it; past
                // the disc limit another orientation's child is a different node, found (or made) through the table
                if (t.symmetry && !(key_of(t.nodes[cur].state) == key_of(key))) {
                    auto it = t.table.find(key_of(key));
                    if (it != t.table.end()) {
                        cur = it->second;
                    } else {
                        cur = t.add_node(key);
                        t.table.emplace(key_of(key), cur);
                        if (node_epoch.size() < t.nodes.size()) node_epoch.resize(t.nodes.size(), 0u);
                    }
                }
            } else {
                mr_leaf s = t.nodes[cur].state;
                if (t.symmetry) s = key;
                else mr_host::apply(t.game, s, t.action[ei]);
                uint32_t id;
                if (t.use_table) { // resolve_child_id: reuse the node of an identical state (shared.rs:412-436)
                    auto it = t.table.find(key_of(s));
                    if (it != t.table.end()) {
                        id = it->second;
                    } else {
                        id = t.add_node(s);
                        t.table.emplace(key_of(s), id);
                    }
                } else {
                    id = t.add_node(s);
                }
                t.child[ei] = (int32_t)id;
                if (node_epoch.size() < t.nodes.size()) node_epoch.resize(t.nodes.size(), 0u);
                edge_changed(path, path.size() - 1, ei, total_before);
                cur = id;
                if (d->expand_threshold > 0) {
                    path.push_back({cur, incoming});
                    lctx.emplace_back();
                    leaf = t.symmetry ? real : t.nodes[cur].state;
                    return;
                }
            }
        }
    }

    // shared.rs:787-805: k-1 extra virtual losses on every edge of the path; the root counts the k trials in flight
    void add_path_virtual_loss(const std::vector<PathEntry> &path, uint32_t k) {
        if (k > 1)
            for (size_t i = 1; i < path.size(); ++i) {
                const uint32_t ei = t.nodes[path[i - 1].node].first + path[i].idx;
                const uint32_t total_before = t.edge[ei].visits + t.edge[ei].vloss;
                t.edge[ei].vloss += k - 1;
                edge_changed(path, i - 1, ei, total_before);
            }
        t.root_inflight += k;
    }

    // update_grave (backprop.rs:619-640) for the k trials of one leaf: every non-root node X of the path gets, per
    // (action, player) occurrence of [playout trace + the path's edges BELOW X], visits += 1 and score += u[player].
    // `amaf` is the kernel's per-leaf occurrence table of the playout traces, [2][S] over action slots.
    void backprop_grave(const std::vector<PathEntry> &path, const mr_action_stat *amaf, const int64_t u[2], uint32_t k) {
        below.clear(); // (slot, player) of the edges below the node being updated
        for (size_t i = path.size(); i-- > 1;) {
            GraveTable &tab = *grave_of(t.nodes[path[i].node], true);
            uint32_t *__restrict__ tv = tab.visits.data();
            int64_t *__restrict__ ts = tab.score.data();
            const size_t cnt = (size_t)2 * S;
            for (size_t e = 0; e < cnt; ++e) tv[e] += amaf[e].visits;
            for (size_t e = 0; e < cnt; ++e) ts[e] += amaf[e].score;
            for (const auto &b : below) {
                const size_t e = (size_t)b.second * S + b.first;
                tv[e] += k;
                ts[e] += u[b.second];
            }
            const Node &parent = t.nodes[path[i - 1].node];
            below.emplace_back((uint32_t)t.slot[parent.first + path[i].idx], (int)parent.player);
        }
    }

    // update_amaf (backprop.rs:565-617) for the k trials of one leaf: at every non-root path node X with parent P,
    // each occurrence of (action, P's mover) in [playout trace + the path's edges BELOW X] whose action is one of
    // P's created children adds one AMAF visit and the trial's utilities to that child's row.  The playout part
    // arrives pre-summed per (player, slot) from the kernel; utilities are +1/-1/0, so the other player's sum
    // is the negation of the mover's.
    void backprop_amaf(const std::vector<PathEntry> &path, const mr_action_stat *amaf, const int64_t u[2], uint32_t k) {
        below.clear();
        for (size_t i = path.size(); i-- > 1;) {
            const Node &parent = t.nodes[path[i - 1].node];
            const int mover = parent.player;
            for (uint32_t j = 0; j < parent.n_children; ++j) {
                if (t.child[parent.first + j] < 0) continue;
                const uint32_t a = t.slot[parent.first + j];
                uint32_t cnt = amaf[(size_t)mover * S + a].visits;
                int64_t sc = amaf[(size_t)mover * S + a].score;
                for (const auto &b : below)
                    if (b.second == mover && b.first == a) {
                        cnt += k;
                        sc += u[mover];
                    }
                AmafRow &e = t.amaf_row[parent.first + j];
                e.visits += cnt;
                e.score0 += mover == 0 ? sc : -sc;
            }
            below.emplace_back((uint32_t)t.slot[parent.first + path[i].idx], mover);
        }
    }
    std::vector<std::pair<uint32_t, int>> below;

    // backprop.rs:908-923 for one batch.  A trial writes replies[p][prev] = action for every consecutive pair of its
    // chronological action list (tree path, then playout) whose second action was made by a player p that did not
    // lose the trial; later writes win.  `upd` holds, per key, the last such write among the PLAYOUT pairs of the
    // batch (order = ((g + 1) << 13) | (4096 + ply), g = leaf * k + playout).  The tree-path pairs of leaf b are
    // written by every trial of b that p did not lose -- last of them playout last_win[b][p] - 1 -- before that
    // trial's playout pairs: order ((g + 1) << 13) | depth index.
    void merge_replies(const std::vector<std::vector<PathEntry>> &paths, uint32_t nb, uint32_t k,
                       std::vector<mr_reply_update> &upd, const std::vector<uint32_t> &last_win) {
        for (uint32_t b = 0; b < nb; ++b) {
            const std::vector<PathEntry> &path = paths[b];
            for (size_t i = 2; i < path.size(); ++i) {
                const Node &grand = t.nodes[path[i - 2].node], &parent = t.nodes[path[i - 1].node];
                const int p = parent.player;
                const uint32_t lw = last_win[(size_t)b * 2 + p];
                if (lw == 0) continue;
                const uint16_t prev = t.action[grand.first + path[i - 1].idx], act = t.action[parent.first + path[i].idx];
                const uint64_t order = ((((uint64_t)b * k + (lw - 1)) + 1) << 13) | (uint64_t)std::min<size_t>(i, 4095);
                mr_reply_update &u = upd[(size_t)p * na + mr_host::action_id(t.game, prev)];
                if (order > u.order) {
                    u.order = order;
                    u.action_plus1 = (uint16_t)(act + 1);
                }
            }
        }
        for (size_t i = 0; i < upd.size(); ++i)
            if (upd[i].order) replies[i] = upd[i].action_plus1;
    }

    // backprop.rs:885-899, tree-path part, for the k trials of one leaf: every pair of consecutive edges of the descent
    // (the pair (last edge, first playout action) and the pairs inside the playouts arrive from the kernel)
    void backprop_bigrams(const std::vector<PathEntry> &path, const mr_leaf_result &r, uint32_t k) {
        const int64_t nonzero = (int64_t)k - r.draws;
        const int64_t u[2] = {(int64_t)r.wins[0] - (nonzero - r.wins[0]), (int64_t)r.wins[1] - (nonzero - r.wins[1])};
        const size_t SS = (size_t)S;
        for (size_t i = 2; i < path.size(); ++i) {
            const Node &grand = t.nodes[path[i - 2].node], &parent = t.nodes[path[i - 1].node];
            const int p = parent.player;
            mr_action_stat &e = bigram[((size_t)p * SS + t.slot[grand.first + path[i - 1].idx]) * SS + t.slot[parent.first + path[i].idx]];
            e.visits += k;
            e.score += (int32_t)u[p];
        }
    }

    // backprop.rs:926-966 for one batch.  Per context (p, own previous action, opponent's answer) the table ends up as:
    // the LAST insert (trials p did not lose), unless a later trial that p lost played the same action in that context;
    // with no insert, the old entry unless some trial that p lost played it.  `ins` arrives with the playout part's last
    // inserts; the tree-path windows of leaf b (three consecutive edges) are inserted by its trials up to
    // last_win[b][p] - 1 and removed by its trials up to last_lose[b][p] - 1, always before that trial's playout windows.
    int merge_replies2(int slot, const std::vector<std::vector<PathEntry>> &paths, uint32_t nb, uint32_t k,
                       std::vector<mr_reply_update> &ins, const std::vector<uint32_t> &last_win,
                       const std::vector<uint32_t> &last_lose, std::vector<uint8_t> &removed) {
        const size_t SS = (size_t)S;
        auto key_of_window = [&](const std::vector<PathEntry> &path, size_t i, int &p, uint16_t &act) -> size_t {
            const Node &n2 = t.nodes[path[i - 3].node], &n1 = t.nodes[path[i - 2].node], &n0 = t.nodes[path[i - 1].node];
            p = n0.player;
            act = t.action[n0.first + path[i].idx];
            return ((size_t)p * SS + t.slot[n2.first + path[i - 2].idx]) * SS + t.slot[n1.first + path[i - 1].idx];
        };
        for (uint32_t b = 0; b < nb; ++b)
            for (size_t i = 3; i < paths[b].size(); ++i) {
                int p;
                uint16_t act;
                const size_t key = key_of_window(paths[b], i, p, act);
                const uint32_t lw = last_win[(size_t)b * 2 + p];
                if (lw == 0) continue;
                const uint64_t order = ((((uint64_t)b * k + (lw - 1)) + 1) << 13) | (uint64_t)std::min<size_t>(i, 4095);
                if (order > ins[key].order) {
                    ins[key].order = order;
                    ins[key].action_plus1 = (uint16_t)(act + 1);
                }
            }
        for (size_t i = 0; i < ins.size(); ++i) // contexts without an insert: the current entry is the candidate
            if (ins[i].order == 0) ins[i].action_plus1 = replies2[i];
        int rc = mr_reply2_resolve(ctx, slot, ins.data(), removed.data());
        if (rc != MR_OK) return rc;
        for (uint32_t b = 0; b < nb; ++b)
            for (size_t i = 3; i < paths[b].size(); ++i) {
                int p;
                uint16_t act;
                const size_t key = key_of_window(paths[b], i, p, act);
                const uint32_t ll = last_lose[(size_t)b * 2 + p];
                if (ll && ins[key].action_plus1 == (uint16_t)(act + 1) && (uint64_t)b * k + ll > (ins[key].order >> 13)) removed[key] = 1;
            }
        for (size_t i = 0; i < ins.size(); ++i) replies2[i] = removed[i] ? (uint16_t)0 : ins[i].action_plus1;
        return MR_OK;
    }

    // k trials of one leaf, aggregated
    void backprop(const std::vector<PathEntry> &path, const mr_leaf_result &r, uint32_t k, const mr_action_stat *amaf) {
        const int64_t nonzero = (int64_t)k - r.draws;
        const int64_t u[2] = {(int64_t)r.wins[0] - (nonzero - r.wins[0]), (int64_t)r.wins[1] - (nonzero - r.wins[1])};
        if (amaf && d->select == MR_SELECT_GRAVE) backprop_grave(path, amaf, u, k);
        if (amaf && d->select == MR_SELECT_AMAF) backprop_amaf(path, amaf, u, k);
        for (size_t i = path.size(); i-- > 0;) {
            Stats &s = i == 0 ? t.root_stats : t.edge[t.nodes[path[i - 1].node].first + path[i].idx];
            s.visits += k;
            s.score0 += u[0];
            s.sumsq += nonzero;
            if (i > 0) s.vloss -= k; // each trial's backprop removes one (backprop.rs:722-725)
            else t.root_inflight -= k;
            if (i > 0 && !mast.empty()) {
                // GLOBAL write for the tree-path part of amaf_actions (backprop.rs:833-870): the edge into this
                // node was played by the parent's mover
                const Node &parent = t.nodes[path[i - 1].node];
                const int pl = parent.player;
                mr_action_stat &m = mast[(size_t)pl * na + mr_host::action_id(t.game, t.action[parent.first + path[i].idx])];
                m.visits += k;
                m.score += (int32_t)u[pl];
            }
        }
        if (path.size() > t.max_depth) t.max_depth = (uint32_t)path.size();
    }
};

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct RootTotals {
    std::vector<mr_root_child> children;
    uint16_t best_action = 0xffff;
    mr_search_stats stats;
};

// One search on one ctx.  `children_out` gets the root children in root action-list order.
int run_search(mr_ctx *ctx, const mr_search_desc *desc, const mr_leaf *root, RootTotals &out) {
    Search s;
    s.ctx = ctx;
    s.d = desc;
    s.t.game = (int)desc->game;
    s.na = mr_num_actions((int)desc->game);
    s.S = mr_num_slots((int)desc->game);
    const bool nst = desc->policy == MR_EPS_NST;
    const bool uses_mast = desc->policy == MR_EPS_MAST || desc->policy == MR_DECISIVE_MAST_WIN || desc->policy == MR_DECISIVE_MAST_WINLOSS || nst ||
                           desc->policy == MR_EPS_LGR2_MAST;
    const size_t n_bigram = nst ? (size_t)2 * s.S * s.S : 0;
    s.bigram.assign(n_bigram, mr_action_stat{0, 0});
    std::vector<mr_action_stat> bigram_delta(n_bigram);
    if (uses_mast) s.mast.assign((size_t)2 * s.na, mr_action_stat{0, 0});
    const bool lgr2 = desc->policy == MR_EPS_LGR2 || desc->policy == MR_EPS_LGR2_MAST;
    const bool lgr = desc->policy == MR_EPS_LGR || lgr2; // LGRF-2 keeps the LGR-1 table too (its inner strategy)
    const size_t n_replies2 = lgr2 ? (size_t)2 * s.S * s.S : 0;
    s.replies2.assign(n_replies2, 0);
    std::vector<mr_reply_update> reply2_ins(n_replies2);
    std::vector<uint8_t> reply2_removed(n_replies2);
    std::vector<uint32_t> last_lose(lgr2 ? (size_t)2 * desc->batch_leaves : 0);
    if (lgr) s.replies.assign((size_t)2 * s.na, 0);
    std::vector<mr_reply_update> reply_upd(lgr ? (size_t)2 * s.na : 0);
    std::vector<uint32_t> last_win(lgr ? (size_t)2 * desc->batch_leaves : 0);
    s.t.nodes.reserve(std::min<size_t>((size_t)desc->max_iterations + 16, (size_t)1 << 22)); // at most one new node per descent
    s.t.use_table = (desc->flags & (MR_SEARCH_TRANSPOSITIONS | MR_SEARCH_SYMMETRY)) != 0;
    s.t.symmetry = (desc->flags & MR_SEARCH_SYMMETRY) != 0 && desc->game == MR_OTHELLO;
    s.t.want_amaf_rows = desc->select == MR_SELECT_AMAF;
    s.t.add_node(*root);
    if (s.t.use_table) s.t.table.emplace(key_of(*root), 0u);
    s.cache_ok = !s.t.use_table;
    s.node_epoch.assign(1, 0u);

    const uint32_t B = desc->batch_leaves, k = desc->playouts_per_leaf;
    const bool want_amaf = desc->select == MR_SELECT_GRAVE || desc->select == MR_SELECT_AMAF;
    mr_search_stats st;
    memset(&st, 0, sizeof st);

    // One buffer set per ABI slot in use.  Batch n goes to slot n % depth; the fixed order is gather(n), submit(n),
    // then -- once `depth` batches are in flight -- wait(n - depth + 1), backprop(n - depth + 1).  The descents of a
    // batch see the virtual losses of every batch still in flight but not yet its results: the staleness the
    // reference's tree-parallel workers live with (search/parallel.rs:188-251), made deterministic.
    uint32_t depth = desc->pipeline_depth ? desc->pipeline_depth : ((desc->flags & MR_SEARCH_PIPELINED) ? 2u : 1u);
    if (depth > MR_NUM_SLOTS) depth = MR_NUM_SLOTS;
    struct BatchBuf {
        std::vector<std::vector<PathEntry>> paths;
        std::vector<mr_leaf> leaves;
        std::vector<mr_leaf_result> results;
        std::vector<mr_action_stat> delta, amaf;
        uint32_t nb = 0;
    };
    std::vector<BatchBuf> buf(depth);
    const size_t delta_len = uses_mast ? (size_t)2 * s.S : 0; // the batch's MAST delta arrives in slot space
    for (auto &bb : buf) {
        bb.paths.resize(B);
        bb.leaves.resize(B);
        bb.results.resize(B);
        bb.delta.resize(delta_len);
        bb.amaf.resize(want_amaf ? (size_t)B * 2 * s.S : 0);
    }
    // slot -> dense action id of the MAST table, per player
    std::vector<uint32_t> slot_id(uses_mast ? (size_t)2 * s.S : 0, 0xffffffffu);
    for (size_t e = 0; e < slot_id.size(); ++e) {
        const int pl = e >= (size_t)s.S ? 1 : 0;
        const int code = mr_slot_action((int)desc->game, pl, (int)(e - (size_t)pl * s.S));
        if (code >= 0) slot_id[e] = (uint32_t)mr_host::action_id((int)desc->game, (uint16_t)code);
    }

    auto gather_and_submit = [&](int slot, uint32_t first, uint32_t nb) -> int {
        BatchBuf &bb = buf[slot];
        bb.nb = nb;
        const double t0 = now_s();
        for (uint32_t b = 0; b < nb; ++b) {
            s.select(first + b, bb.paths[b], bb.leaves[b]);
            s.add_path_virtual_loss(bb.paths[b], k);
            if ((lgr || nst) && bb.paths[b].size() > 1) { // playout()'s prev_action: the last edge of the descent (shared.rs:763-778)
                const std::vector<PathEntry> &path = bb.paths[b];
                const Node &parent = s.t.nodes[path[path.size() - 2].node];
                bb.leaves[b].prev_action_plus1 = (uint16_t)(s.t.action[parent.first + path.back().idx] + 1);
                if (lgr2 && path.size() > 2) { // and the edge before it (the first two plies' LGRF-2 contexts)
                    const Node &grand = s.t.nodes[path[path.size() - 3].node];
                    bb.leaves[b].prev2_action_plus1 = (uint16_t)(s.t.action[grand.first + path[path.size() - 2].idx] + 1);
                }
            }
        }
        if (lgr) {
            int rc = mr_set_replies(ctx, (int)desc->game, s.replies.data());
            if (rc != MR_OK) return rc;
        }
        if (lgr2) {
            int rc = mr_set_replies2(ctx, (int)desc->game, s.replies2.data());
            if (rc != MR_OK) return rc;
        }
        if (nst) { // frozen for the batch, like the MAST table
            int rc = mr_set_bigrams(ctx, (int)desc->game, s.bigram.data(), desc->nst_backoff_threshold);
            if (rc != MR_OK) return rc;
        }
        const double t1 = now_s();
        st.seconds_select += t1 - t0;
        mr_batch_desc bd;
        memset(&bd, 0, sizeof bd);
        bd.game = desc->game;
        bd.policy = desc->policy;
        bd.n_leaves = nb;
        bd.playouts_per_leaf = k;
        bd.max_playout_depth = desc->max_playout_depth;
        bd.flags = (uses_mast ? MR_WANT_MAST_DELTA : 0) | (want_amaf ? MR_WANT_AMAF : 0) | MR_TABLES_IN_SLOTS;
        bd.seed = desc->seed;
        bd.leaf_id_base = first;
        bd.epsilon = desc->epsilon;
        const int rc = mr_submit(ctx, slot, &bd, bb.leaves.data(), s.mast.empty() ? nullptr : s.mast.data());
        st.seconds_submit += now_s() - t1;
        return rc;
    };
    auto wait_and_backprop = [&](int slot) -> int {
        BatchBuf &bb = buf[slot];
        const double t0 = now_s();
        int rc = mr_wait(ctx, slot, bb.results.data(), want_amaf ? bb.amaf.data() : nullptr,
                         bb.delta.empty() ? nullptr : bb.delta.data(), nullptr, nullptr, nullptr);
        if (rc != MR_OK) return rc;
        if (lgr) {
            rc = mr_reply_updates(ctx, slot, reply_upd.data(), last_win.data());
            if (rc != MR_OK) return rc;
        }
        if (nst) {
            rc = mr_bigram_delta(ctx, slot, bigram_delta.data());
            if (rc != MR_OK) return rc;
        }
        if (lgr2) {
            rc = mr_reply2_inserts(ctx, slot, reply2_ins.data(), last_lose.data());
            if (rc != MR_OK) return rc;
            rc = s.merge_replies2(slot, bb.paths, bb.nb, k, reply2_ins, last_win, last_lose, reply2_removed);
            if (rc != MR_OK) return rc;
        }
        const double t1 = now_s();
        ++s.epoch; // statistics change: every cached child score is stale
        if (lgr) s.merge_replies(bb.paths, bb.nb, k, reply_upd, last_win);
        for (size_t i = 0; i < n_bigram; ++i) {
            s.bigram[i].visits += bigram_delta[i].visits;
            s.bigram[i].score += bigram_delta[i].score;
        }
        for (uint32_t b = 0; b < bb.nb; ++b) {
            if (nst) s.backprop_bigrams(bb.paths[b], bb.results[b], k);
            s.backprop(bb.paths[b], bb.results[b], k, want_amaf ? bb.amaf.data() + (size_t)b * 2 * s.S : nullptr);
            st.playout_plies += bb.results[b].sum_depth;
        }
        for (size_t e = 0; e < bb.delta.size(); ++e) { // playout-trace part of the GLOBAL write
            if (!bb.delta[e].visits) continue;
            const int pl = e >= (size_t)s.S ? 1 : 0;
            mr_action_stat &m = s.mast[(size_t)pl * s.na + slot_id[e]];
            m.visits += bb.delta[e].visits;
            m.score += bb.delta[e].score;
        }
        st.seconds_gpu += t1 - t0;
        st.seconds_backprop += now_s() - t1;
        st.batches += 1;
        st.playouts += (uint64_t)bb.nb * k;
        return MR_OK;
    };

    mr_set_kernel_timing(ctx, 0); // two event records and an event wait per batch are not free at 20 us per batch
    struct TimingGuard {
        mr_ctx *c;
        ~TimingGuard() { mr_set_kernel_timing(c, 1); }
    } timing_guard{ctx};

    uint32_t done = 0, submitted = 0, applied = 0; // batch counters
    while (done < desc->max_iterations) {
        const uint32_t nb = std::min(B, desc->max_iterations - done);
        int rc = gather_and_submit((int)(submitted % depth), done, nb);
        if (rc != MR_OK) { // leave no batch in flight behind an error
            for (; applied < submitted; ++applied) mr_wait(ctx, (int)(applied % depth), buf[applied % depth].results.data(), nullptr, nullptr, nullptr, nullptr, nullptr);
            return rc;
        }
        done += nb;
        ++submitted;
        if (submitted - applied >= depth) {
            rc = wait_and_backprop((int)(applied % depth));
            ++applied;
            if (rc != MR_OK) {
                for (; applied < submitted; ++applied) mr_wait(ctx, (int)(applied % depth), buf[applied % depth].results.data(), nullptr, nullptr, nullptr, nullptr, nullptr);
                return rc;
            }
        }
    }
    for (; applied < submitted; ++applied) {
        int rc = wait_and_backprop((int)(applied % depth));
        if (rc != MR_OK) {
            for (++applied; applied < submitted; ++applied) mr_wait(ctx, (int)(applied % depth), buf[applied % depth].results.data(), nullptr, nullptr, nullptr, nullptr, nullptr);
            return rc;
        }
    }

    // final move: RobustChild over the root children with the same random-offset arg-max (core.rs:138-198)
    st.nodes = s.t.nodes.size();
    st.max_tree_depth = s.t.max_depth;
    out.stats = st;
    out.children.clear();
    out.best_action = 0xffff;
    const Node &rootn = s.t.nodes[0];
    if (rootn.status != EXPANDED || rootn.n_children == 0) return MR_OK;
    const uint32_t n = rootn.n_children;
    const int player = rootn.player;
    const double q_unvisited =
        desc->q_init == MR_QINIT_INFINITY ? 10000.0 : (s.t.root_stats.visits == 0 ? 10000.0 : expected_score(s.t.root_stats, player));
    const uint32_t r = mulhi32(philox_word0(desc->seed, 0xffffffffu, 0, 0, 2), 16u * n);
    uint32_t i = r / 16u;
    const uint32_t stride = kPrimes[r % 16u];
    bool have = false;
    int64_t best_v = 0;
    double best_q = 0.0;
    uint32_t best_i = i;
    for (uint32_t c = 0; c < n; ++c) {
        const Stats &e = s.t.edge[rootn.first + i];
        const bool created = s.t.child[rootn.first + i] >= 0;
        const int64_t v = created ? (int64_t)e.visits : 0;
        const double q = created ? expected_score(e, player) : q_unvisited;
        if (!have || v > best_v || (v == best_v && q > best_q)) {
            have = true;
            best_v = v;
            best_q = q;
            best_i = i;
        }
        i = (i + stride) % n;
    }
    out.best_action = s.t.action[rootn.first + best_i];
    out.children.resize(n);
    for (uint32_t c = 0; c < n; ++c) {
        const Stats &e = s.t.edge[rootn.first + c];
        out.children[c].action = s.t.action[rootn.first + c];
        out.children[c].pad = 0;
        out.children[c].visits = e.visits;
        out.children[c].score[0] = e.score0;
        out.children[c].score[1] = -e.score0;
    }
    return MR_OK;
}

int check_desc(const mr_search_desc *desc) {
    if (desc->game >= MR_NUM_GAMES || desc->batch_leaves == 0 || desc->playouts_per_leaf == 0 || desc->max_iterations == 0) return MR_ERR_INVALID;
    if (desc->select > MR_SELECT_AMAF) return MR_ERR_INVALID;
    if (desc->rave_schedule > MR_RAVE_MIN_MSE) return MR_ERR_INVALID;
    if (desc->pipeline_depth > MR_NUM_SLOTS) return MR_ERR_INVALID;
    if ((desc->flags & MR_SEARCH_SYMMETRY) && desc->game == MR_OTHELLO) {
        // canonical nodes hold canonical actions: the action-keyed tables (MAST, replies, n-grams, AMAF / GRAVE) are
        // written in the real orientation by the kernels and would need a per-path translation that is not built
        const bool plain_policy = desc->policy == MR_UNIFORM || desc->policy == MR_DECISIVE_WIN || desc->policy == MR_DECISIVE_WINLOSS ||
                                  desc->policy == MR_DECISIVE_WINLOSSDRAW || desc->policy == MR_DECISIVE_ANTI;
        if (!plain_policy || desc->select > MR_SELECT_UCB1_TUNED) return MR_ERR_UNSUPPORTED;
    }
    return MR_OK;
}

// nothing may unwind across the C boundary: allocation failures of the host tree become an error code
template <class F>
int guarded(F &&f) {
    try {
        return f();
    } catch (const std::bad_alloc &) {
        return MR_ERR_NO_MEMORY;
    } catch (...) {
        return MR_ERR_INVALID;
    }
}

} // namespace

extern "C" int mr_generate_actions(int game, const mr_leaf *state, uint16_t *out) {
    if (game < 0 || game >= MR_NUM_GAMES || !state || !out) return -1;
    return mr_host::generate_actions(game, *state, out);
}
extern "C" int mr_apply(int game, mr_leaf *state, uint16_t action) {
    if (game < 0 || game >= MR_NUM_GAMES || !state) return MR_ERR_INVALID;
    mr_host::apply(game, *state, action);
    return MR_OK;
}
extern "C" int mr_terminal_status(int game, const mr_leaf *state) {
    if (game < 0 || game >= MR_NUM_GAMES || !state) return -1;
    return mr_host::terminal_status(game, *state);
}

extern "C" int mr_canonical_representation(int game, const mr_leaf *state, mr_leaf *canonical_out) {
    if (game < 0 || game >= MR_NUM_GAMES || !state || !canonical_out) return -1;
    if (game == MR_OTHELLO) return mr_host::othello_canonical(*state, *canonical_out);
    *canonical_out = *state;
    return 0;
}

extern "C" int mr_search(mr_ctx *ctx, const mr_search_desc *desc, const mr_leaf *root, uint16_t *best_action,
                         mr_root_child *children, uint32_t *n_children, mr_search_stats *stats) {
    if (!ctx || !desc || !root || !best_action) return MR_ERR_INVALID;
    if (int rc = check_desc(desc)) return rc;
    return guarded([&]() -> int {
        RootTotals rt;
        const int rc = run_search(ctx, desc, root, rt);
        if (rc != MR_OK) return rc;
        *best_action = rt.best_action;
        if (n_children) *n_children = (uint32_t)rt.children.size();
        if (children && n_children)
            for (size_t c = 0; c < rt.children.size() && c < MR_MAX_ROOT_CHILDREN; ++c) children[c] = rt.children[c];
        if (stats) *stats = rt.stats;
        return MR_OK;
    });
}

// search/parallel.rs:56-115.  The reference draws one seed per worker from the search's own rng; here searcher i's seed
// is words 0..1 of Philox(desc->seed; ctr = (i, 0, 0, 3)), and the final random_best draw is word 0 of ctr = (0, 0, 1, 3).
extern "C" int mr_search_root_parallel(mr_ctx *const *ctxs, int n_searchers, const mr_search_desc *desc, const mr_leaf *root,
                                       uint16_t *best_action, mr_root_child *children, uint32_t *n_children, mr_search_stats *stats) {
    if (!ctxs || n_searchers <= 0 || !desc || !root || !best_action) return MR_ERR_INVALID;
    for (int i = 0; i < n_searchers; ++i)
        if (!ctxs[i]) return MR_ERR_INVALID;
    if (int rc = check_desc(desc)) return rc;
    return guarded([&]() -> int {
        std::vector<RootTotals> totals((size_t)n_searchers);
        std::vector<mr_search_desc> descs((size_t)n_searchers, *desc);
        std::vector<int> rcs((size_t)n_searchers, MR_OK);
        for (int i = 0; i < n_searchers; ++i) {
            uint32_t w[4];
            philox4(desc->seed, (uint32_t)i, 0, 0, 3, w);
            descs[(size_t)i].seed = ((uint64_t)w[1] << 32) | w[0];
        }
        auto work = [&](int i) {
            rcs[(size_t)i] = guarded([&]() -> int { return run_search(ctxs[i], &descs[(size_t)i], root, totals[(size_t)i]); });
        };
        std::vector<std::thread> threads;
        for (int i = 1; i < n_searchers; ++i) threads.emplace_back(work, i);
        work(0);
        for (auto &th : threads) th.join();
        for (int i = 0; i < n_searchers; ++i)
            if (rcs[(size_t)i] != MR_OK) return rcs[(size_t)i];
        // merged totals per root action (children are in root action-list order in every searcher; a searcher that
        // never expanded its root reports none)
        std::vector<mr_root_child> merged;
        for (const RootTotals &rt : totals) {
            if (rt.children.empty()) continue;
            if (merged.empty()) {
                merged = rt.children;
                continue;
            }
            for (size_t c = 0; c < merged.size(); ++c) {
                merged[c].visits += rt.children[c].visits;
                merged[c].score[0] += rt.children[c].score[0];
                merged[c].score[1] += rt.children[c].score[1];
            }
        }
        *best_action = 0xffff;
        if (n_children) *n_children = (uint32_t)merged.size();
        if (!merged.empty()) {
            // random_best over merged visits (parallel.rs:111-115, util.rs:135-163)
            const uint32_t n = (uint32_t)merged.size();
            const uint32_t r = mulhi32(philox_word0(desc->seed, 0, 0, 1, 3), 16u * n);
            uint32_t i = r / 16u;
            const uint32_t stride = kPrimes[r % 16u];
            uint32_t best_i = i, best_v = 0;
            bool have = false;
            for (uint32_t c = 0; c < n; ++c) {
                if (!have || merged[i].visits > best_v) {
                    have = true;
                    best_v = merged[i].visits;
                    best_i = i;
                }
                i = (i + stride) % n;
            }
            *best_action = merged[best_i].action;
            if (children)
                for (size_t c = 0; c < merged.size() && c < MR_MAX_ROOT_CHILDREN; ++c) children[c] = merged[c];
        }
        if (stats)
            for (int i = 0; i < n_searchers; ++i) stats[i] = totals[(size_t)i].stats;
        return MR_OK;
    });
}

// ---- known-answer hooks (include/mcts_rollout.h "Known-answer hooks") ----
extern "C" int mr_flat_monte_carlo(mr_ctx *ctx, int game, const mr_leaf *root, uint32_t samples_per_move, uint32_t max_rollout_depth,
                                   int use_ucb1, double ucb1_c, uint32_t r, uint64_t seed, uint16_t *best_action, uint32_t *wins_out,
                                   uint16_t *actions_out, uint32_t *n_actions) {
    if (!ctx || game < 0 || game >= MR_NUM_GAMES || !root || !best_action || !wins_out || !n_actions || samples_per_move == 0)
        return MR_ERR_INVALID;
    return guarded([&]() -> int {
        uint16_t acts[MR_MAX_ROOT_CHILDREN];
        const int cnt = mr_host::terminal_status(game, *root) == MR_NOT_TERMINAL ? mr_host::generate_actions(game, *root, acts) : 0;
        if (cnt <= 0) return MR_ERR_INVALID; // flat_mc.rs:103-105 panics on a terminal state
        const int player = root->turn;
        std::vector<mr_leaf> leaves((size_t)cnt, *root);
        for (int i = 0; i < cnt; ++i) {
            mr_host::apply(game, leaves[(size_t)i], acts[i]);
            leaves[(size_t)i].prev_action_plus1 = 0;
            leaves[(size_t)i].prev2_action_plus1 = 0;
        }
        mr_batch_desc bd;
        memset(&bd, 0, sizeof bd);
        bd.game = (uint32_t)game;
        bd.policy = MR_UNIFORM;
        bd.n_leaves = (uint32_t)cnt;
        bd.playouts_per_leaf = samples_per_move;
        bd.max_playout_depth = max_rollout_depth > 1 ? max_rollout_depth - 1 : 1;
        bd.seed = seed;
        std::vector<mr_leaf_result> res((size_t)cnt);
        const int rc = mr_playout_batch(ctx, &bd, leaves.data(), nullptr, res.data(), nullptr, nullptr, nullptr, nullptr, nullptr);
        if (rc != MR_OK) return rc;
        // random_best over the children's values (util.rs:135-163): the first strict maximum in visiting order
        const double n = (double)samples_per_move;
        uint32_t i = (r / 16u) % (uint32_t)cnt, best_i = 0;
        const uint32_t stride = kPrimes[r % 16u];
        bool have = false;
        double best = 0.0;
        for (int k = 0; k < cnt; ++k) {
            const double w = (double)res[i].wins[player];
            const double v = use_ucb1 ? w / n + ucb1_c * (std::log(n) / n) : w;
            if (!have || v > best) {
                have = true;
                best = v;
                best_i = i;
            }
            i = (uint32_t)(((uint64_t)i + stride) % (uint32_t)cnt);
        }
        for (int k = 0; k < cnt; ++k) {
            wins_out[k] = res[(size_t)k].wins[player];
            if (actions_out) actions_out[k] = acts[k];
        }
        *n_actions = (uint32_t)cnt;
        *best_action = acts[best_i];
        return MR_OK;
    });
}

extern "C" double mr_select_score(const mr_search_desc *desc, int player, const mr_edge_stats *current, const mr_edge_stats *child,
                                  uint32_t grave_n, int64_t grave_q) {
    if (!desc || !current || player < 0 || player > 1) return NAN;
    Stats cur;
    cur.visits = current->visits;
    cur.vloss = current->vloss;
    cur.score0 = current->score[0];
    cur.sumsq = current->sumsq[0];
    const double parent_log = std::log(std::max(1.0, (double)cur.visits));
    if (!child) return unvisited_child_score(desc, cur, player, parent_log);
    Stats e;
    e.visits = child->visits;
    e.vloss = child->vloss;
    e.score0 = child->score[0];
    e.sumsq = child->sumsq[0];
    AmafRow row;
    row.visits = child->amaf_visits;
    row.score0 = child->amaf_score[0];
    return created_child_score(desc, e, &row, player, parent_log, grave_n, (double)grave_q);
}

extern "C" int mr_kat_update_amaf(int game, const mr_leaf *root, const uint16_t *created, uint32_t n_created, uint16_t processed,
                                  const mr_action_stat *amaf, const int64_t utilities[2], uint32_t k, mr_edge_stats *children_out,
                                  uint32_t *n_children) {
    if (game < 0 || game >= MR_NUM_GAMES || !root || !created || !amaf || !utilities || !children_out || !n_children) return MR_ERR_INVALID;
    return guarded([&]() -> int {
        mr_search_desc d;
        memset(&d, 0, sizeof d);
        d.game = (uint32_t)game;
        d.select = MR_SELECT_AMAF;
        Search s;
        s.ctx = nullptr;
        s.d = &d;
        s.t.game = game;
        s.t.want_amaf_rows = true;
        s.na = mr_num_actions(game);
        s.S = mr_num_slots(game);
        s.t.add_node(*root);
        s.t.expand(0);
        const Node &rootn = s.t.nodes[0];
        if (rootn.status != EXPANDED) return MR_ERR_INVALID;
        std::vector<PathEntry> path;
        path.push_back({0, 0});
        auto create = [&](uint16_t a) -> int {
            for (uint32_t j = 0; j < rootn.n_children; ++j)
                if (s.t.action[rootn.first + j] == a) {
                    if (s.t.child[rootn.first + j] < 0) {
                        mr_leaf st = rootn.state;
                        mr_host::apply(game, st, a);
                        const int32_t id = (int32_t)s.t.add_node(st);
                        s.t.child[s.t.nodes[0].first + j] = id;
                    }
                    return (int)j;
                }
            return -1;
        };
        for (uint32_t c = 0; c < n_created; ++c)
            if (create(created[c]) < 0) return MR_ERR_INVALID;
        const int pj = create(processed);
        if (pj < 0) return MR_ERR_INVALID;
        const Node &r2 = s.t.nodes[0];
        path.push_back({(uint32_t)s.t.child[r2.first + (uint32_t)pj], (uint32_t)pj});
        // the occurrence table arrives dense ([2][NA], as the ABI's default form); the driver works in slot space
        std::vector<mr_action_stat> in_slots((size_t)2 * s.S, mr_action_stat{0, 0});
        for (int pl = 0; pl < 2; ++pl)
            for (int sl = 0; sl < s.S; ++sl) {
                const int code = mr_slot_action(game, pl, sl);
                if (code >= 0) in_slots[(size_t)pl * s.S + sl] = amaf[(size_t)pl * s.na + mr_host::action_id(game, (uint16_t)code)];
            }
        s.backprop_amaf(path, in_slots.data(), utilities, k);
        *n_children = r2.n_children;
        for (uint32_t j = 0; j < r2.n_children; ++j) {
            mr_edge_stats &o = children_out[j];
            memset(&o, 0, sizeof o);
            o.action = s.t.action[r2.first + j];
            o.created = s.t.child[r2.first + j] >= 0;
            o.amaf_visits = s.t.amaf_row[r2.first + j].visits;
            o.amaf_score[0] = s.t.amaf_row[r2.first + j].score0;
            o.amaf_score[1] = -s.t.amaf_row[r2.first + j].score0;
        }
        return MR_OK;
    });
}
