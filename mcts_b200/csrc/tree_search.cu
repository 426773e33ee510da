// tree_search.cu -- batched leaf-gather MCTS driver over the GPU rollout engine (host code only).
//
// What it mirrors in the reference (thomasmarsh/mcts):
//   descent ............ select_step, search/shared.rs:569-736 (expand :220-289, one virtual loss per edge :665-671)
//   child choice ....... UCB1 / UCB1-Tuned, select/ucb.rs:9-146, unvisited value node.rs:139-159,
//                        arg-max with random offset + prime stride, strict '>', select/mod.rs:343-386
//   leaf fan-out ....... num_rollouts_per_leaf with k-1 extra virtual loss, search_impl.rs:101-113, shared.rs:787-805
//   backprop ........... backprop.rs:704-731 with ChildArray::update / NodeStats::update (node.rs:295-302,673-681)
//                        summed over the k trials of a leaf (exact: utilities are +1 / 0 / -1)
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
// The tree (arena, edge statistics in the parent's child arrays, root_stats) stays on the host; only the
// simulation phase crosses the C ABI (mr_submit / mr_wait).  Tie-break draws of the tree come from the same
// Philox4x32-10 generator as the playouts, on their own stream: ctr = (descent id, depth, 0, 2).
#include <chrono>
#include <cmath>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "host_games.h"
#include "mcts_rollout.h"

namespace {

const uint32_t kPrimes[16] = {14323, 18713, 19463, 30553, 33469, 45343, 50221, 51991,
                              53201, 56923, 64891, 72763, 74471, 81647, 92581, 94693};

uint32_t philox_word0(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
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
    return c0;
}
inline uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

struct Stats { // NodeStats / one ChildArray row
    uint32_t visits = 0, vloss = 0;
    int64_t score[2] = {0, 0}, sumsq[2] = {0, 0};
    uint32_t amaf_visits = 0; // ChildArray amaf rows (node.rs:392-396,711-719): every player's count moves together
    int64_t amaf_score[2] = {0, 0};
};
// node.rs:126-133
inline double expected_score(const Stats &s, int player) {
    if (s.visits == 0) return 0.0;
    const double lv = (double)s.vloss;
    return ((double)s.score[player] - lv) / ((double)s.visits + lv);
}

enum : uint8_t { UNEXPANDED = 0, EXPANDED = 1, TERMINAL = 2 };

struct Node {
    mr_leaf state;
    uint8_t status = UNEXPANDED;
    uint8_t player = 0;
    uint16_t n_children = 0;
    uint32_t first = 0; // index of child 0 in the edge arrays
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
    std::unordered_map<StateKey, uint32_t, StateKeyHash> table; // use_transpositions: state -> node
    std::vector<Node> nodes;
    std::vector<uint16_t> action;
    std::vector<int32_t> child; // -1 = not created yet
    std::vector<Stats> edge;
    Stats root_stats;
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
            child.push_back(-1);
            edge.push_back(Stats());
        }
        n.status = EXPANDED;
    }
};

struct GraveTable { // per state: [2 * NA] ActionStats (node.rs:51-55) with integer scores, as two arrays so that the
                    // per-leaf table merge of backprop_grave is two straight vectorisable loops
    std::vector<uint32_t> visits;
    std::vector<int64_t> score;
    explicit GraveTable(size_t n) : visits(n, 0u), score(n, 0) {}
};

struct Search {
    mr_ctx *ctx;
    const mr_search_desc *d;
    Tree t;
    std::vector<mr_action_stat> mast; // [2][NA], TreeStats.player_actions
    std::vector<mr_action_stat> bigram; // [2][S][S], TreeStats.player_bigram_actions over compact action slots (MR_EPS_NST)
    std::vector<uint16_t> replies2;   // [2][S][S], TreeStats.player_replies2 over action slots: 1 + reply code (MR_EPS_LGR2)
    std::vector<uint16_t> replies;    // [2][NA], TreeStats.player_replies: 1 + reply code by (player, preceding action id)
    int na;
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
        grave.emplace_back((size_t)2 * na);
        return &grave.back();
    }

    // select/rave.rs:60-75
    double rave_beta(uint32_t n, uint32_t amaf_n) const {
        (void)amaf_n;
        if (d->rave_schedule == MR_RAVE_THRESHOLD) {
            const double rave = (double)d->rave_param;
            return std::max(0.0, rave - (double)n) / rave;
        }
        const double k = (double)d->rave_param;
        return std::sqrt(k / (3.0 * (double)n + k));
    }
    // select/rave.rs:157-184 get_ref: the child itself when its incoming edge has total_visits >= threshold (decided in
    // best_child), else the nearest such node on the path, else the root
    uint32_t grave_ref_ancestor(const std::vector<PathEntry> &path) const {
        for (size_t i = path.size(); i-- > 1;) {
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

    // select/mod.rs:343-386 with select/ucb.rs scores
    uint32_t best_child(const std::vector<PathEntry> &path, uint32_t descent) {
        const Node &node = t.nodes[path.back().node];
        const int player = node.player;
        const Stats &cur = current_stats(path);
        const double parent_log = std::log(std::max(1.0, (double)cur.visits));
        const double c = d->exploration_c;
        // node.rs:139-159 value_estimate_unvisited
        double q_unvisited;
        if (d->q_init == MR_QINIT_INFINITY) q_unvisited = 10000.0;
        else q_unvisited = cur.visits == 0 ? 10000.0 : expected_score(cur, player);
        double unvisited_value;
        if (d->select == MR_SELECT_UCB1_TUNED) unvisited_value = q_unvisited + (parent_log * std::min(1.0, 1.0) + c * std::sqrt(parent_log));
        else if (d->select == MR_SELECT_GRAVE || d->select == MR_SELECT_AMAF)
            unvisited_value = q_unvisited; // rave.rs:226-230, amaf.rs:76-79: no exploration term
        else unvisited_value = q_unvisited + c * std::sqrt(parent_log);

        const uint32_t n = node.n_children;
        const uint32_t r = mulhi32(philox_word0(d->seed, descent, (uint32_t)path.size() - 1, 0, 2), 16u * n);
        uint32_t i = r / 16u;
        const uint32_t stride = kPrimes[r % 16u];
        bool have = false;
        double best = 0.0;
        uint32_t best_i = i;
        bool anc_done = false;
        uint32_t anc_ref = 0;
        const GraveTable *anc_tab = nullptr;
        for (uint32_t k = 0; k < n; ++k) {
            const Stats &e = t.edge[node.first + i];
            double score;
            if (t.child[node.first + i] < 0) {
                score = unvisited_value;
            } else {
                const double exploit = expected_score(e, player);
                const double total = (double)(e.visits + e.vloss);
                if (d->select == MR_SELECT_UCB1_TUNED) {
                    const double var = std::max(0.0, (double)e.sumsq[player] / total - exploit * exploit);
                    const double frac = parent_log / total;
                    score = exploit + (frac * std::min(1.0, var) + c * std::sqrt(frac));
                } else if (d->select == MR_SELECT_GRAVE) { // rave.rs:196-224
                    // get_ref: the child itself once its edge has the threshold visits, else the nearest such ancestor --
                    // which is the same node for every child of this node, so it is resolved once (anc_ref / anc_tab)
                    const bool own_ref = e.visits + e.vloss >= d->grave_threshold;
                    if (!own_ref && !anc_done) {
                        anc_ref = grave_ref_ancestor(path);
                        anc_tab = grave_of(t.nodes[anc_ref], false);
                        anc_done = true;
                    }
                    uint32_t amaf_n = 0;
                    double amaf_q = 0.0;
                    if (const GraveTable *tab = own_ref ? grave_of(t.nodes[(uint32_t)t.child[node.first + i]], false) : anc_tab) {
                        const size_t e = (size_t)player * na + mr_host::action_id(t.game, t.action[node.first + i]);
                        amaf_n = tab->visits[e];
                        amaf_q = (double)tab->score[e];
                    }
                    const uint32_t n_tot = e.visits + e.vloss;
                    const double explore = c * std::sqrt(parent_log / (double)n_tot);
                    const double b = rave_beta(n_tot, amaf_n);
                    const double amaf = amaf_n == 0 ? 0.0 : amaf_q / (double)amaf_n;
                    score = (1.0 - b) * exploit + b * amaf + explore;
                } else if (d->select == MR_SELECT_AMAF) { // amaf.rs:58-74
                    const double amaf_n = (double)std::max<uint32_t>(1u, e.amaf_visits);
                    const double amaf = (double)e.amaf_score[player] / amaf_n;
                    const double ucb1 = exploit + c * std::sqrt(parent_log / total);
                    score = d->amaf_alpha * amaf + (1.0 - d->amaf_alpha) * ucb1;
                } else {
                    score = exploit + c * std::sqrt(parent_log / total);
                }
            }
            if (!have || score > best) {
                have = true;
                best = score;
                best_i = i;
            }
            i = (i + stride) % n;
        }
        return best_i;
    }

    // search/shared.rs:569-736
    void select(uint32_t descent, std::vector<PathEntry> &path) {
        path.clear();
        uint32_t cur = 0, incoming = 0;
        for (;;) {
            path.push_back({cur, incoming});
            const uint32_t num_visits = current_stats(path).visits;
            if (t.nodes[cur].status == TERMINAL) return;
            if (num_visits < d->expand_threshold) return;
            if (t.nodes[cur].status == UNEXPANDED) {
                t.expand(cur);
                if (t.nodes[cur].status == TERMINAL) return;
            }
            const uint32_t best = best_child(path, descent);
            incoming = best;
            const uint32_t slot = t.nodes[cur].first + best;
            t.edge[slot].vloss += 1; // one virtual loss per traversed edge
            if (t.child[slot] >= 0) {
                cur = (uint32_t)t.child[slot];
            } else {
                mr_leaf s = t.nodes[cur].state;
                mr_host::apply(t.game, s, t.action[slot]);
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
                t.child[slot] = (int32_t)id;
                cur = id;
                if (d->expand_threshold > 0) {
                    path.push_back({cur, incoming});
                    return;
                }
            }
        }
    }

    // shared.rs:787-805: k-1 extra virtual losses on every edge of the path
    void add_path_virtual_loss(const std::vector<PathEntry> &path, uint32_t extra) {
        for (size_t i = 1; i < path.size(); ++i) t.edge[t.nodes[path[i - 1].node].first + path[i].idx].vloss += extra;
    }

    // update_grave (backprop.rs:619-640) for the k trials of one leaf: every non-root node X of the path gets, per
    // (action, player) occurrence of [playout trace + the path's edges BELOW X], visits += 1 and score += u[player].
    // `amaf` is the kernel's per-leaf occurrence table of the playout traces ([2][NA]).
    void backprop_grave(const std::vector<PathEntry> &path, const mr_action_stat *amaf, const int64_t u[2], uint32_t k) {
        std::vector<std::pair<uint32_t, int>> below; // (action id, player) of the edges below the node being updated
        for (size_t i = path.size(); i-- > 1;) {
            GraveTable &tab = *grave_of(t.nodes[path[i].node], true);
            uint32_t *__restrict__ tv = tab.visits.data();
            int64_t *__restrict__ ts = tab.score.data();
            const size_t cnt = (size_t)2 * na;
            for (size_t e = 0; e < cnt; ++e) tv[e] += amaf[e].visits;
            for (size_t e = 0; e < cnt; ++e) ts[e] += amaf[e].score;
            for (const auto &b : below) {
                const size_t e = (size_t)b.second * na + b.first;
                tv[e] += k;
                ts[e] += u[b.second];
            }
            const Node &parent = t.nodes[path[i - 1].node];
            below.emplace_back((uint32_t)mr_host::action_id(t.game, t.action[parent.first + path[i].idx]), (int)parent.player);
        }
    }

    // update_amaf (backprop.rs:565-617) for the k trials of one leaf: at every non-root path node X with parent P,
    // each occurrence of (action, P's mover) in [playout trace + the path's edges BELOW X] whose action is one of
    // P's created children adds one AMAF visit and the trial's utilities to that child's row.  The playout part
    // arrives pre-summed per (player, action) from the kernel; utilities are +1/-1/0, so the other player's sum
    // is the negation of the mover's.
    void backprop_amaf(const std::vector<PathEntry> &path, const mr_action_stat *amaf, const int64_t u[2], uint32_t k) {
        std::vector<std::pair<uint32_t, int>> below;
        for (size_t i = path.size(); i-- > 1;) {
            const Node &parent = t.nodes[path[i - 1].node];
            const int mover = parent.player;
            for (uint32_t j = 0; j < parent.n_children; ++j) {
                if (t.child[parent.first + j] < 0) continue;
                const uint32_t a = (uint32_t)mr_host::action_id(t.game, t.action[parent.first + j]);
                uint32_t cnt = amaf[(size_t)mover * na + a].visits;
                int64_t sc = amaf[(size_t)mover * na + a].score;
                for (const auto &b : below)
                    if (b.second == mover && b.first == a) {
                        cnt += k;
                        sc += u[mover];
                    }
                Stats &e = t.edge[parent.first + j];
                e.amaf_visits += cnt;
                e.amaf_score[mover] += sc;
                e.amaf_score[1 - mover] -= sc;
            }
            below.emplace_back((uint32_t)mr_host::action_id(t.game, t.action[parent.first + path[i].idx]), mover);
        }
    }

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
        const size_t S = (size_t)mr_num_slots(t.game);
        for (size_t i = 2; i < path.size(); ++i) {
            const Node &grand = t.nodes[path[i - 2].node], &parent = t.nodes[path[i - 1].node];
            const int p = parent.player;
            const uint16_t prev = t.action[grand.first + path[i - 1].idx], act = t.action[parent.first + path[i].idx];
            mr_action_stat &e = bigram[((size_t)p * S + (size_t)mr_action_slot(t.game, 1 - p, prev)) * S + (size_t)mr_action_slot(t.game, p, act)];
            e.visits += k;
            e.score += (int32_t)u[p];
        }
    }

    // backprop.rs:926-966 for one batch.  Per context (p, own previous action, opponent's answer) the table ends up as:
    // the LAST insert (trials p did not lose), unless a later trial that p lost played the same action in that context;
    // with no insert, the old entry unless some trial that p lost played it.  `ins` arrives with the playout part's last
    // inserts; the tree-path windows of leaf b (three consecutive edges) are inserted by its trials up to
    // last_win[b][p] - 1 and removed by its trials up to last_lose[b][p] - 1, always before that trial's playout windows.
    int merge_replies2(mr_ctx *ctx, int slot, const std::vector<std::vector<PathEntry>> &paths, uint32_t nb, uint32_t k,
                       std::vector<mr_reply_update> &ins, const std::vector<uint32_t> &last_win,
                       const std::vector<uint32_t> &last_lose, std::vector<uint8_t> &removed) {
        const size_t S = (size_t)mr_num_slots(t.game);
        auto key_of_window = [&](const std::vector<PathEntry> &path, size_t i, int &p, uint16_t &act) -> size_t {
            const Node &n2 = t.nodes[path[i - 3].node], &n1 = t.nodes[path[i - 2].node], &n0 = t.nodes[path[i - 1].node];
            p = n0.player;
            const uint16_t own = t.action[n2.first + path[i - 2].idx], opp = t.action[n1.first + path[i - 1].idx];
            act = t.action[n0.first + path[i].idx];
            return ((size_t)p * S + (size_t)mr_action_slot(t.game, p, own)) * S + (size_t)mr_action_slot(t.game, 1 - p, opp);
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
            for (int p = 0; p < 2; ++p) {
                s.score[p] += u[p];
                s.sumsq[p] += nonzero;
            }
            if (i > 0) s.vloss -= k; // each trial's backprop removes one (backprop.rs:722-725)
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

extern "C" int mr_search(mr_ctx *ctx, const mr_search_desc *desc, const mr_leaf *root, uint16_t *best_action,
                         mr_root_child *children, uint32_t *n_children, mr_search_stats *stats) {
    if (!ctx || !desc || !root || !best_action) return MR_ERR_INVALID;
    if (desc->game >= MR_NUM_GAMES || desc->batch_leaves == 0 || desc->playouts_per_leaf == 0 || desc->max_iterations == 0)
        return MR_ERR_INVALID;
    if (desc->select > MR_SELECT_AMAF) return MR_ERR_INVALID;
    Search s;
    s.ctx = ctx;
    s.d = desc;
    s.t.game = (int)desc->game;
    s.na = mr_num_actions((int)desc->game);
    const bool nst = desc->policy == MR_EPS_NST;
    const bool uses_mast = desc->policy == MR_EPS_MAST || desc->policy == MR_DECISIVE_MAST_WIN || desc->policy == MR_DECISIVE_MAST_WINLOSS || nst ||
                           desc->policy == MR_EPS_LGR2_MAST;
    const size_t n_bigram = nst ? (size_t)2 * mr_num_slots((int)desc->game) * mr_num_slots((int)desc->game) : 0;
    s.bigram.assign(n_bigram, mr_action_stat{0, 0});
    std::vector<mr_action_stat> bigram_delta(n_bigram);
    if (uses_mast) s.mast.assign((size_t)2 * s.na, mr_action_stat{0, 0});
    const bool lgr2 = desc->policy == MR_EPS_LGR2 || desc->policy == MR_EPS_LGR2_MAST;
    const bool lgr = desc->policy == MR_EPS_LGR || lgr2; // LGRF-2 keeps the LGR-1 table too (its inner strategy)
    const size_t n_replies2 = lgr2 ? (size_t)2 * mr_num_slots((int)desc->game) * mr_num_slots((int)desc->game) : 0;
    s.replies2.assign(n_replies2, 0);
    std::vector<mr_reply_update> reply2_ins(n_replies2);
    std::vector<uint8_t> reply2_removed(n_replies2);
    std::vector<uint32_t> last_lose(lgr2 ? (size_t)2 * desc->batch_leaves : 0);
    if (lgr) s.replies.assign((size_t)2 * s.na, 0);
    std::vector<mr_reply_update> reply_upd(lgr ? (size_t)2 * s.na : 0);
    std::vector<uint32_t> last_win(lgr ? (size_t)2 * desc->batch_leaves : 0);
    s.t.nodes.reserve(desc->max_iterations + 16);
    s.t.use_table = (desc->flags & MR_SEARCH_TRANSPOSITIONS) != 0;
    s.t.add_node(*root);
    if (s.t.use_table) s.t.table.emplace(key_of(*root), 0u);

    const uint32_t B = desc->batch_leaves, k = desc->playouts_per_leaf;
    std::vector<std::vector<PathEntry>> paths(B);
    std::vector<mr_leaf> leaves(B);
    std::vector<mr_leaf_result> results(B);
    std::vector<mr_action_stat> delta(uses_mast ? (size_t)2 * s.na : 0);
    const bool want_amaf = desc->select == MR_SELECT_GRAVE || desc->select == MR_SELECT_AMAF;
    std::vector<mr_action_stat> amaf(want_amaf ? (size_t)B * 2 * s.na : 0);
    mr_search_stats st;
    memset(&st, 0, sizeof st);

    // Two buffer sets, one per ABI slot.  Synchronous mode uses set 0 only.  Pipelined mode
    // (MR_SEARCH_PIPELINED) gathers batch N+1 while batch N is on the GPU: its descents see batch N's virtual
    // losses but not yet its results -- the same staleness the reference's tree-parallel workers live with
    // (search/parallel.rs:188-251), made deterministic by the fixed order gather(N+1), submit(N+1), wait(N), backprop(N).
    const bool pipelined = (desc->flags & MR_SEARCH_PIPELINED) != 0;
    struct BatchBuf {
        std::vector<std::vector<PathEntry>> paths;
        std::vector<mr_leaf> leaves;
        std::vector<mr_leaf_result> results;
        std::vector<mr_action_stat> delta, amaf;
        uint32_t nb = 0;
    } buf[2];
    for (auto &bb : buf) {
        bb.paths.resize(B);
        bb.leaves.resize(B);
        bb.results.resize(B);
        bb.delta.resize(uses_mast ? (size_t)2 * s.na : 0);
        bb.amaf.resize(want_amaf ? (size_t)B * 2 * s.na : 0);
    }
    (void)paths;
    (void)leaves;
    (void)results;
    (void)delta;
    (void)amaf;

    auto gather_and_submit = [&](int slot, uint32_t first, uint32_t nb) -> int {
        BatchBuf &bb = buf[slot];
        bb.nb = nb;
        const double t0 = now_s();
        for (uint32_t b = 0; b < nb; ++b) {
            s.select(first + b, bb.paths[b]);
            if (k > 1) s.add_path_virtual_loss(bb.paths[b], k - 1);
            bb.leaves[b] = s.t.nodes[bb.paths[b].back().node].state;
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
        st.seconds_select += now_s() - t0;
        mr_batch_desc bd;
        memset(&bd, 0, sizeof bd);
        bd.game = desc->game;
        bd.policy = desc->policy;
        bd.n_leaves = nb;
        bd.playouts_per_leaf = k;
        bd.max_playout_depth = desc->max_playout_depth;
        bd.flags = (uses_mast ? MR_WANT_MAST_DELTA : 0) | (want_amaf ? MR_WANT_AMAF : 0);
        bd.seed = desc->seed;
        bd.leaf_id_base = first;
        bd.epsilon = desc->epsilon;
        return mr_submit(ctx, slot, &bd, bb.leaves.data(), s.mast.empty() ? nullptr : s.mast.data());
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
            rc = s.merge_replies2(ctx, slot, bb.paths, bb.nb, k, reply2_ins, last_win, last_lose, reply2_removed);
            if (rc != MR_OK) return rc;
        }
        const double t1 = now_s();
        if (lgr) s.merge_replies(bb.paths, bb.nb, k, reply_upd, last_win);
        for (size_t i = 0; i < n_bigram; ++i) {
            s.bigram[i].visits += bigram_delta[i].visits;
            s.bigram[i].score += bigram_delta[i].score;
        }
        for (uint32_t b = 0; b < bb.nb; ++b) {
            if (nst) s.backprop_bigrams(bb.paths[b], bb.results[b], k);
            s.backprop(bb.paths[b], bb.results[b], k, want_amaf ? bb.amaf.data() + (size_t)b * 2 * s.na : nullptr);
            st.playout_plies += bb.results[b].sum_depth;
        }
        for (size_t i = 0; i < bb.delta.size(); ++i) { // playout-trace part of the GLOBAL write
            s.mast[i].visits += bb.delta[i].visits;
            s.mast[i].score += bb.delta[i].score;
        }
        st.seconds_gpu += t1 - t0;
        st.seconds_backprop += now_s() - t1;
        st.batches += 1;
        st.playouts += (uint64_t)bb.nb * k;
        return MR_OK;
    };

    uint32_t done = 0;
    int in_flight = -1; // slot of the batch on the GPU (pipelined mode)
    while (done < desc->max_iterations) {
        const uint32_t nb = std::min(B, desc->max_iterations - done);
        const int slot = pipelined ? (in_flight == 0 ? 1 : 0) : 0;
        int rc = gather_and_submit(slot, done, nb);
        if (rc != MR_OK) return rc;
        done += nb;
        if (pipelined) {
            if (in_flight >= 0) {
                rc = wait_and_backprop(in_flight);
                if (rc != MR_OK) return rc;
            }
            in_flight = slot;
        } else {
            rc = wait_and_backprop(slot);
            if (rc != MR_OK) return rc;
        }
    }
    if (pipelined && in_flight >= 0) {
        int rc = wait_and_backprop(in_flight);
        if (rc != MR_OK) return rc;
    }

    // final move: RobustChild over the root children with the same random-offset arg-max (core.rs:138-198)
    const Node &rootn = s.t.nodes[0];
    if (rootn.status != EXPANDED || rootn.n_children == 0) {
        *best_action = 0xffff;
        if (n_children) *n_children = 0;
        if (stats) *stats = st;
        return MR_OK;
    }
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
    *best_action = s.t.action[rootn.first + best_i];
    if (children && n_children) {
        *n_children = n;
        for (uint32_t c = 0; c < n && c < MR_MAX_ROOT_CHILDREN; ++c) {
            const Stats &e = s.t.edge[rootn.first + c];
            children[c].action = s.t.action[rootn.first + c];
            children[c].pad = 0;
            children[c].visits = e.visits;
            children[c].score[0] = e.score[0];
            children[c].score[1] = e.score[1];
        }
    }
    st.nodes = s.t.nodes.size();
    st.max_tree_depth = s.t.max_depth;
    if (stats) *stats = st;
    return MR_OK;
}
