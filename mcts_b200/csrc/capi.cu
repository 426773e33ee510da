// capi.cu -- host side of the C ABI declared in include/mcts_rollout.h.
//
// One mr_ctx drives one or more GPUs of this process.  Per GPU: one stream, one task-queue counter per
// slot, grow-on-demand device buffers and pinned host staging.  A batch is cut into tasks
// (leaf, run-of-playouts); the global playout range is split evenly over the ctx's GPUs (no data-path
// collective: every playout is independent given (leaf state, Philox key)); per-leaf results are summed
// on the host.  There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "int_peak.cuh"
#include "kernels.h"

using namespace mr;

namespace {

thread_local std::string g_global_error;

struct GameInfo {
    int num_actions;
    uint32_t depth_bound; // longest possible playout; 0 = unbounded
};
const GameInfo kGames[MR_NUM_GAMES] = {
    {7, 42},    // Connect4: 42 cells
    {4096, 256}, // Breakthrough: pawns only advance: <= 2 * 16 * 7 = 224 plies
    {4096, 0},   // Knightthrough: knights also jump backwards: unbounded
    {65, 128},   // Othello: 60 placements + at most one pass between placements
    {121, 121},  // Hex 11x11
};

template <class T>
struct DevBuf {
    T *d = nullptr;
    T *h = nullptr; // pinned (host-mapped under UVA: the device can read / write it directly)
    size_t cap = 0;
    bool fresh = false; // set when reserve() (re)allocated: the contents are undefined
    cudaError_t reserve(size_t n, bool want_host) {
        if (n <= cap && (!want_host || h)) return cudaSuccess;
        size_t ncap = std::max(n, cap);
        if (n > cap) {
            if (d) cudaFree(d);
            d = nullptr;
            cudaError_t e = cudaMalloc(&d, ncap * sizeof(T));
            if (e != cudaSuccess) {
                release(); // a failed allocation leaves an EMPTY buffer, never a capacity without memory behind it
                return e;
            }
        }
        if (want_host && (n > cap || !h)) {
            if (h) cudaFreeHost(h);
            h = nullptr;
            cudaError_t e = cudaMallocHost(&h, ncap * sizeof(T));
            if (e != cudaSuccess) {
                release();
                return e;
            }
        }
        cap = ncap;
        fresh = true;
        return cudaSuccess;
    }
    void release() {
        if (d) cudaFree(d);
        if (h) cudaFreeHost(h);
        d = nullptr;
        h = nullptr;
        cap = 0;
    }
};

struct SlotDev {
    DevBuf<mr_leaf> leaves;
    DevBuf<mr_leaf_result> results;
    DevBuf<mr_action_stat> amaf;
    DevBuf<mr_action_stat> mast_delta;
    DevBuf<uint16_t> trace;
    DevBuf<mr_playout_record> rec;
    DevBuf<mr_leaf> final_state;
    DevBuf<uint16_t> lgr_in;                // MR_EPS_LGR: frozen reply table
    DevBuf<unsigned long long> lgr_stamps;  // MR_EPS_LGR: last winning occurrence per (player, preceding action)
    DevBuf<uint32_t> lgr_last;              // MR_EPS_LGR: per leaf, last playout each player did not lose
    DevBuf<uint16_t> lgr2_in;               // MR_EPS_LGR2: frozen 2-ply reply table [2][S][S]
    DevBuf<unsigned long long> lgr2_stamps; // MR_EPS_LGR2: last insert per context (pass 1)
    DevBuf<uint32_t> lgr2_last_lose;        // MR_EPS_LGR2: per leaf, last playout each player lost
    DevBuf<unsigned long long> lgr2_target; // MR_EPS_LGR2 pass 2: candidate entry per context
    DevBuf<uint8_t> lgr2_removed;           // MR_EPS_LGR2 pass 2: contexts whose candidate a later losing playout removes
    DevBuf<mr_leaf_result> results2;        // MR_EPS_LGR2 pass 2: scratch results of the replay
    KernelParams kp_saved{};                // the batch's launch parameters, for the pass-2 replay
    DevBuf<uint64_t> nst_planes;            // MR_EPS_NST: rank planes per context [2][S + 1][KINDS][MAST_BITS]
    DevBuf<uint32_t> nst_nbits;             // MR_EPS_NST: [2][S + 1]
    DevBuf<uint16_t> nst_rank_small;        // MR_EPS_NST: [2][S + 1][8] (Connect4)
    DevBuf<unsigned long long> nst_delta;   // MR_EPS_NST: [2][S][S] packed (visits | score << 32)
    unsigned long long *d_counter = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    cudaStream_t stream = nullptr;  // the slot's own stream: batches of different slots are independent
    cudaEvent_t ev_done = nullptr;  // after the slot's last enqueued operation
    // low-latency path (fused epilogue, common.cuh batch_epilogue)
    unsigned int *d_done = nullptr; // ticket counter of the blocks that have finished
    uint32_t *flag_h = nullptr;     // host-mapped: the epilogue publishes the batch's sequence number here
    uint32_t seq = 0;               // sequence number of the last batch submitted on the fast path
    bool fast = false;              // the batch in flight uses the fused epilogue
    bool wait_event = false;        // its completion is ev_done (else: the host flag)
    // whole-buffer-is-zero invariants of the fast path (the epilogue re-zeroes what a batch touched)
    bool results_clean = false, mast_clean = false, amaf_clean = false, counter_clean = false;
    // what this device covers for the batch in flight
    uint64_t g_begin = 0, g_end = 0; // global playout index range
};

struct Device {
    int id = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    SlotDev slot[MR_NUM_SLOTS];
    int mast_game = -1; // game of the MAST snapshot held in ctx->mast_*
    DevBuf<uint16_t> scratch[MR_NUM_SLOTS + 1]; // per slot (+1 for mr_launch_device): launches may overlap across streams
    unsigned long long *d_counter_ext = nullptr; // work counter for mr_launch_device
};

struct SlotState {
    bool busy = false;
    mr_batch_desc desc{};
    float kernel_ms = 0.f;
    bool lgr_ready = false; // a completed MR_EPS_LGR batch whose updates can still be read
    bool nst_ready = false; // a completed MR_EPS_NST batch whose bigram delta can still be read
    bool lgr2_ready = false; // a completed MR_EPS_LGR2 batch that can still be resolved
    uint32_t trace_cap = 0;
};


} // namespace

struct mr_ctx {
    std::vector<Device> dev;
    SlotState slot[MR_NUM_SLOTS];
    std::string error;
    uint64_t launches = 0;
    bool timing = true; // record kernel begin / end events per batch (mr_last_kernel_ms); the search driver turns it off
    std::unordered_map<const void *, int> occupancy; // kernel -> resident blocks per SM (queried once)
    // MAST snapshot as rank bit-planes (copied into KernelParams at launch)
    uint64_t mast_planes[2][MAST_KINDS][MAST_BITS] = {};
    uint32_t mast_nbits[2] = {0, 0};
    uint16_t mast_rank_small[2][8] = {};
    int mast_game = -1;
    // Last Good Reply table (mr_set_replies), [2][num_actions] of its game
    std::vector<uint16_t> replies;
    int replies_game = -1;
    // LGRF-2 table (mr_set_replies2), [2][S][S] of its game
    std::vector<uint16_t> replies2;
    int replies2_game = -1;
    // N-gram Selection Technique (mr_set_bigrams): bigram[2][S][S] of its game and the backoff threshold; the per-context
    // rank tables built from it and the submit's unigram table
    std::vector<mr_action_stat> bigram;
    int bigram_game = -1;
    uint32_t nst_threshold = 5; // Nst::default (simulate.rs:875-881)
    std::vector<uint64_t> nst_planes;
    std::vector<uint32_t> nst_nbits;
    std::vector<uint16_t> nst_rank_small;
    // NCCL (loaded lazily with dlopen; see mr_comm_*)
    void *nccl_lib = nullptr;
    void *nccl_comm = nullptr;
    int comm_rank = 0, comm_world = 1;
};

namespace {

int fail(mr_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->error = buf;
    g_global_error = buf;
    return code;
}

#define CUDA_TRY(ctx, expr)                                                                              \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(ctx, MR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// ---- kernel table: (game, policy, flag variant) -> instantiation (the instantiations live in kernels_<game>.cu) ----
int flag_variant(uint32_t flags) {
    if (flags & (MR_WANT_TRACE | MR_WANT_FINAL_STATE)) return 4;
    return (int)(flags & 3u);
}

kernel_fn pick_kernel(int game, int policy, int v, uint32_t flags) {
    // WinLoss / WinLossDraw pick exactly what Win picks (DESIGN.md "Decisive move")
    if (policy == MR_DECISIVE_WINLOSS || policy == MR_DECISIVE_WINLOSSDRAW) policy = MR_DECISIVE_WIN;
    if (policy == MR_DECISIVE_MAST_WINLOSS) policy = MR_DECISIVE_MAST_WIN;
    switch (game) {
    case MR_CONNECT4: return pick_connect4(policy, v);
    case MR_BREAKTHROUGH: return pick_breakthrough(policy, v);
    case MR_KNIGHTTHROUGH: return pick_knightthrough(policy, v);
    case MR_OTHELLO: return pick_othello(policy, v);
    case MR_HEX11: return pick_hex11(policy, v, flags);
    }
    return nullptr;
}

// ---- MAST snapshot -> order-preserving dense ranks ----
// Mast::select_move compares unigram scores score/visits as f64, unvisited => 1.0 (simulate.rs:806-848).
// score is an integer sum of +-1 utilities, so for visits < 2^26 the f64 comparison equals the exact
// rational comparison score_a * visits_b <=> score_b * visits_a.  Ranks keep order and ties exactly.
struct Rational {
    int64_t num;
    int64_t den;
};
inline bool rat_less(const Rational &a, const Rational &b) { return (__int128)a.num * b.den < (__int128)b.num * a.den; }
inline bool rat_eq(const Rational &a, const Rational &b) { return (__int128)a.num * b.den == (__int128)b.num * a.den; }

// (src, kind) -> dst for the two games that have a MAST kernel; -1 when the move leaves the board
int through_dst(int game, int player, int src, int kind) {
    const int row = src / 8, col = src % 8;
    if (game == MR_BREAKTHROUGH) {
        if (kind > 2) return -1;
        const int dr = player == 0 ? -1 : 1, dc = kind - 1; // W, F, E
        const int r = row + dr, c = col + dc;
        return (r < 0 || r > 7 || c < 0 || c > 7) ? -1 : r * 8 + c;
    }
    static const int off[8][2] = {{-2, -1}, {-2, 1}, {-1, -2}, {-1, 2}, {1, -2}, {1, 2}, {2, -1}, {2, 1}}; // ascending dst
    const int r = row + off[kind][0], c = col + off[kind][1];
    return (r < 0 || r > 7 || c < 0 || c > 7) ? -1 : r * 8 + c;
}

// (player, cell, kind) -> dense action id of the MAST table, -1 when there is no such action.  Breakthrough /
// Knightthrough: cell = source square, kind = move kind.  Othello: cell = square, one kind (PASS is always forced).
// Connect4: cell = column.
int mast_action_id(int game, int player, int cell, int kind) {
    if (game == MR_BREAKTHROUGH || game == MR_KNIGHTTHROUGH) {
        const int dst = through_dst(game, player, cell, kind);
        return dst < 0 ? -1 : cell * 64 + dst;
    }
    if (game == MR_OTHELLO) return kind == 0 ? cell : -1;
    if (game == MR_HEX11) return (kind < 2 && kind * 64 + cell < 121) ? kind * 64 + cell : -1; // kind = 64-bit half of the board
    if (game == MR_CONNECT4) return (kind == 0 && cell < 7) ? cell : -1;
    return -1;
}

// the real actions of (game, player) as (cell, kind, dense id), built once
struct MastAction {
    uint8_t cell, kind;
    uint16_t id;
};
const std::vector<MastAction> &mast_actions(int game, int player) {
    static std::vector<MastAction> cache[MR_NUM_GAMES][2];
    static bool built = false;
    if (!built) {
        for (int g = 0; g < MR_NUM_GAMES; ++g) {
            const int kinds = g == MR_BREAKTHROUGH ? 3 : (g == MR_KNIGHTTHROUGH ? 8 : (g == MR_HEX11 ? 2 : 1));
            for (int pl = 0; pl < 2; ++pl)
                for (int src = 0; src < 64; ++src)
                    for (int q = 0; q < kinds; ++q) {
                        const int id = mast_action_id(g, pl, src, q);
                        if (id >= 0) cache[g][pl].push_back({(uint8_t)src, (uint8_t)q, (uint16_t)id});
                    }
        }
        built = true;
    }
    return cache[game][player];
}

// The unigram score as a sort key.  score / visits in f64 is order- AND equality-preserving for visits < 2^26: two
// different fractions with denominators below 2^26 differ by more than 2^-52 while both lie in [-1, 1] (half an ulp is
// at most 2^-54), and equal fractions round to the same double.  (This is also why the reference's own f64 comparison,
// simulate.rs:806-817, equals the exact rational one.)
int compute_mast_planes(mr_ctx *ctx, int game, const mr_action_stat *mast) {
    const int na = kGames[game].num_actions;
    memset(ctx->mast_planes, 0, sizeof ctx->mast_planes);
    memset(ctx->mast_rank_small, 0, sizeof ctx->mast_rank_small);
    for (int pl = 0; pl < 2; ++pl) {
        const mr_action_stat *tab = mast + (size_t)pl * na;
        const std::vector<MastAction> &acts = mast_actions(game, pl);
        // distinct values over the REAL actions of this player only (ids that no move can produce are ignored)
        double key[513], uniq[514];
        size_t nu = 0;
        uniq[nu++] = 1.0; // unvisited => 1.0
        for (size_t a = 0; a < acts.size(); ++a) {
            const mr_action_stat &e = tab[acts[a].id];
            if (e.visits >= (1u << 26))
                return fail(ctx, MR_ERR_INVALID, "MAST visits %u >= 2^26: exact score comparison not guaranteed", e.visits);
            key[a] = e.visits > 0 ? (double)e.score / (double)e.visits : 1.0;
            if (e.visits > 0) uniq[nu++] = key[a];
        }
        std::sort(uniq, uniq + nu);
        nu = (size_t)(std::unique(uniq, uniq + nu) - uniq);
        uint32_t nbits = 0;
        while ((1u << nbits) < nu) ++nbits;
        if (nbits > (uint32_t)MAST_BITS) return fail(ctx, MR_ERR_INVALID, "too many distinct MAST values (%zu)", nu);
        ctx->mast_nbits[pl] = nbits;
        for (size_t a = 0; a < acts.size(); ++a) {
            const uint32_t rk = (uint32_t)(std::lower_bound(uniq, uniq + nu, key[a]) - uniq);
            const uint64_t bit = 1ull << acts[a].cell;
            for (uint32_t b = 0; b < nbits; ++b)
                if ((rk >> b) & 1u) ctx->mast_planes[pl][acts[a].kind][b] |= bit;
            if (game == MR_CONNECT4) ctx->mast_rank_small[pl][acts[a].cell] = (uint16_t)rk;
        }
    }
    ctx->mast_game = game;
    return MR_OK;
}

// ---- compact per-player action slots (the key space of the bigram tables and of the kernels' shared-memory tables) ----
const int kSlots[MR_NUM_GAMES] = {7, 192, 512, 65, 121};
const int kKinds[MR_NUM_GAMES] = {1, 3, 8, 1, 2};
const int kKnightDelta[8] = {-17, -15, -10, -6, 6, 10, 15, 17};

int action_slot(int game, int player, uint32_t code) {
    if (game == MR_BREAKTHROUGH) {
        const int src = (int)(code >> 8), dst = (int)(code & 0xffu), kind = dst - src - (player ? 7 : -9);
        if (src >= 64 || dst >= 64 || kind < 0 || kind > 2 || through_dst(game, player, src, kind) != dst) return -1;
        return src * 3 + kind;
    }
    if (game == MR_KNIGHTTHROUGH) {
        const int src = (int)(code >> 8), dst = (int)(code & 0xffu);
        if (src >= 64 || dst >= 64) return -1;
        for (int q = 0; q < 8; ++q)
            if (dst - src == kKnightDelta[q] && through_dst(game, player, src, q) == dst) return src * 8 + q;
        return -1;
    }
    return (int)code < kSlots[game] ? (int)code : -1;
}

int slot_action(int game, int player, int slot) {
    if (slot < 0 || slot >= kSlots[game]) return -1;
    if (game == MR_BREAKTHROUGH || game == MR_KNIGHTTHROUGH) {
        const int kinds = kKinds[game], src = slot / kinds, dst = through_dst(game, player, src, slot % kinds);
        return dst < 0 ? -1 : (src << 8) | dst;
    }
    return slot;
}

// (cell, kind) of the rank planes -> slot; the inverse of the kernels' plane indexing
int plane_slot(int game, int cell, int kind) {
    if (game == MR_BREAKTHROUGH) return cell * 3 + kind;
    if (game == MR_KNIGHTTHROUGH) return cell * 8 + kind;
    if (game == MR_HEX11) return kind * 64 + cell;
    return cell;
}

// Nst::select_move (simulate.rs:898-930) scores an action by the bigram average of (preceding action, action) once that
// pair has >= backoff_threshold visits, else by its unigram score.  The score of every action is therefore a function
// of the CONTEXT (mover, preceding action) alone: the host ranks the mover's real actions once per context (exactly, as
// for MAST) and ships rank bit-planes [2][S + 1][KINDS][MAST_BITS]; context 0 = no preceding action = the unigram ranks.
int compute_nst_tables(mr_ctx *ctx, int game, const mr_action_stat *mast) {
    const int na = kGames[game].num_actions, S = kSlots[game], kinds = kKinds[game];
    const bool have_bigram = ctx->bigram_game == game && !ctx->bigram.empty();
    const size_t n_ctx = (size_t)2 * (S + 1);
    ctx->nst_planes.assign(n_ctx * kinds * MAST_BITS, 0);
    ctx->nst_nbits.assign(n_ctx, 0);
    ctx->nst_rank_small.assign(n_ctx * 8, 0);
    struct Act {
        int cell, kind, slot;
        Rational uni;
    };
    for (int pl = 0; pl < 2; ++pl) {
        const mr_action_stat *tab = mast + (size_t)pl * na;
        std::vector<Act> acts;
        for (int cell = 0; cell < 64; ++cell)
            for (int q = 0; q < kinds; ++q) {
                const int id = mast_action_id(game, pl, cell, q);
                if (id < 0) continue;
                const mr_action_stat &e = tab[id];
                if (e.visits >= (1u << 26))
                    return fail(ctx, MR_ERR_INVALID, "MAST visits %u >= 2^26: exact score comparison not guaranteed", e.visits);
                acts.push_back({cell, q, plane_slot(game, cell, q), e.visits > 0 ? Rational{(int64_t)e.score, (int64_t)e.visits} : Rational{1, 1}});
            }
        std::vector<Rational> val(acts.size()), uniq;
        size_t unigram_ctx = 0; // index of the context whose tables equal the unigram ones (context 0 of this player)
        for (int c = 0; c <= S; ++c) {
            const size_t ci = (size_t)pl * (S + 1) + c;
            bool differs = false;
            for (size_t a = 0; a < acts.size(); ++a) {
                val[a] = acts[a].uni;
                if (c > 0 && have_bigram) {
                    const mr_action_stat &b = ctx->bigram[((size_t)pl * S + (c - 1)) * S + acts[a].slot];
                    if (b.visits >= (1u << 26))
                        return fail(ctx, MR_ERR_INVALID, "bigram visits %u >= 2^26: exact score comparison not guaranteed", b.visits);
                    if (b.visits > 0 && b.visits >= ctx->nst_threshold) {
                        val[a] = {(int64_t)b.score, (int64_t)b.visits};
                        differs = true;
                    }
                }
            }
            if (c == 0) unigram_ctx = ci;
            if (c > 0 && !differs) { // no pair of this context has reached the threshold: the unigram tables
                std::copy_n(&ctx->nst_planes[unigram_ctx * kinds * MAST_BITS], (size_t)kinds * MAST_BITS, &ctx->nst_planes[ci * kinds * MAST_BITS]);
                ctx->nst_nbits[ci] = ctx->nst_nbits[unigram_ctx];
                std::copy_n(&ctx->nst_rank_small[unigram_ctx * 8], 8, &ctx->nst_rank_small[ci * 8]);
                continue;
            }
            uniq.assign(val.begin(), val.end());
            uniq.push_back({1, 1});
            std::sort(uniq.begin(), uniq.end(), rat_less);
            uniq.erase(std::unique(uniq.begin(), uniq.end(), rat_eq), uniq.end());
            uint32_t nbits = 0;
            while ((1u << nbits) < uniq.size()) ++nbits;
            if (nbits > (uint32_t)MAST_BITS) return fail(ctx, MR_ERR_INVALID, "too many distinct NST values (%zu)", uniq.size());
            ctx->nst_nbits[ci] = nbits;
            uint64_t *planes = &ctx->nst_planes[ci * kinds * MAST_BITS];
            for (size_t a = 0; a < acts.size(); ++a) {
                const uint32_t rk = (uint32_t)(std::lower_bound(uniq.begin(), uniq.end(), val[a], rat_less) - uniq.begin());
                for (uint32_t b = 0; b < nbits; ++b)
                    if ((rk >> b) & 1u) planes[(size_t)acts[a].kind * MAST_BITS + b] |= 1ull << acts[a].cell;
                if (game == MR_CONNECT4) ctx->nst_rank_small[ci * 8 + acts[a].cell] = (uint16_t)rk;
            }
        }
    }
    return MR_OK;
}

// util.rs:129-132
const uint32_t kPrimes[16] = {14323, 18713, 19463, 30553, 33469, 45343, 50221, 51991,
                              53201, 56923, 64891, 72763, 74471, 81647, 92581, 94693};

// fills MR_STRIDE / MR_STRIDE_INV / MR_RCP32 on the current device
cudaError_t upload_random_best_tables() {
    static uint8_t stride[129][16], inv[129][16];
    static uint32_t rcp[129];
    memset(stride, 0, sizeof stride);
    memset(inv, 0, sizeof inv);
    memset(rcp, 0, sizeof rcp);
    for (uint32_t n = 1; n <= 128; ++n) {
        rcp[n] = n == 1 ? 0u : (uint32_t)((1ull << 32) / n) + 1u;
        for (int k = 0; k < 16; ++k) {
            const uint32_t s = kPrimes[k] % n;
            stride[n][k] = (uint8_t)s;
            uint32_t iv = 0;
            for (uint32_t c = 0; c < n; ++c)
                if ((c * s) % n == 1 % n) {
                    iv = c;
                    break;
                }
            inv[n][k] = (uint8_t)iv;
        }
    }
    // every kernel translation unit holds its own copy of the tables
    const table_upload_fn uploads[] = {upload_tables_connect4, upload_tables_breakthrough, upload_tables_knightthrough,
                                       upload_tables_othello, upload_tables_hex11};
    for (table_upload_fn f : uploads) {
        cudaError_t e = f(stride, inv, rcp);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

int validate(mr_ctx *ctx, const mr_batch_desc *d, bool has_mast, kernel_fn *fn_out, uint32_t *trace_cap_out) {
    if (!d) return fail(ctx, MR_ERR_INVALID, "null batch descriptor");
    if (d->game >= MR_NUM_GAMES) return fail(ctx, MR_ERR_INVALID, "unknown game %u", d->game);
    if (d->n_leaves == 0 || d->playouts_per_leaf == 0)
        return fail(ctx, MR_ERR_INVALID, "empty batch (n_leaves=%u, playouts_per_leaf=%u)", d->n_leaves, d->playouts_per_leaf);
    if (d->flags & ~31u) return fail(ctx, MR_ERR_INVALID, "unknown flags 0x%x", d->flags);
    kernel_fn fn = pick_kernel((int)d->game, (int)d->policy, flag_variant(d->flags), d->flags);
    if (!fn) return fail(ctx, MR_ERR_UNSUPPORTED, "policy %u is not built for game %u", d->policy, d->game);
    if (policy_uses_mast((int)d->policy)) {
        if (!has_mast) return fail(ctx, MR_ERR_INVALID, "a MAST policy needs a MAST snapshot (mast_in / mr_set_mast)");
        if (!(d->epsilon >= 0.0 && d->epsilon <= 1.0)) return fail(ctx, MR_ERR_INVALID, "epsilon %f outside [0,1]", d->epsilon);
    }
    if (policy_is_lgr((int)d->policy)) {
        if (policy_is_lgr2((int)d->policy) && ctx->replies2_game >= 0 && ctx->replies2_game != (int)d->game)
            return fail(ctx, MR_ERR_INVALID, "the 2-ply reply table was set for game %d, the batch is game %u", ctx->replies2_game, d->game);
        if (!(d->epsilon >= 0.0 && d->epsilon <= 1.0)) return fail(ctx, MR_ERR_INVALID, "epsilon %f outside [0,1]", d->epsilon);
        if (ctx->replies_game >= 0 && ctx->replies_game != (int)d->game)
            return fail(ctx, MR_ERR_INVALID, "the reply table was set for game %d, the batch is game %u", ctx->replies_game, d->game);
        if ((uint64_t)d->n_leaves * d->playouts_per_leaf >= (1ull << 37)) return fail(ctx, MR_ERR_INVALID, "batch too large for MR_EPS_LGR");
    }
    if (d->policy == MR_EPS_NST && ctx->bigram_game >= 0 && ctx->bigram_game != (int)d->game)
        return fail(ctx, MR_ERR_INVALID, "the bigram table was set for game %d, the batch is game %u", ctx->bigram_game, d->game);
    if ((d->flags & MR_WANT_TRACE) && d->max_trace == 0) return fail(ctx, MR_ERR_INVALID, "MR_WANT_TRACE needs max_trace > 0");
    uint32_t cap = 0;
    // (the Hex fill-then-search kernels derive AMAF and the MAST delta from board masks; every per-ply kernel uses the slot trace)
    const bool hex_fast = d->game == MR_HEX11 && (d->policy == MR_UNIFORM || d->policy == MR_EPS_MAST);
    const bool needs_scratch = ((d->flags & (MR_WANT_MAST_DELTA | MR_WANT_AMAF)) && !hex_fast) || policy_is_lgr((int)d->policy) ||
                               d->policy == MR_EPS_NST;
    if (needs_scratch) {
        const uint32_t bound = kGames[d->game].depth_bound;
        cap = d->max_playout_depth ? d->max_playout_depth : bound;
        if (bound && cap > bound) cap = bound;
        if (cap == 0)
            return fail(ctx, MR_ERR_INVALID, "AMAF / MAST-delta / LGR / NST need a bounded playout: set max_playout_depth for game %u", d->game);
        if (cap > 4096) return fail(ctx, MR_ERR_INVALID, "max_playout_depth %u too large for AMAF / MAST-delta / LGR (limit 4096)", cap);
    }
    if ((uint64_t)d->n_leaves * d->playouts_per_leaf > (1ull << 40)) return fail(ctx, MR_ERR_INVALID, "batch too large");
    *fn_out = fn;
    *trace_cap_out = cap;
    return MR_OK;
}

// entries per player of the AMAF / MAST-delta tables of a batch: dense action ids, or compact slots
inline int table_len(const mr_batch_desc *d) { return (d->flags & MR_TABLES_IN_SLOTS) ? kSlots[d->game] : kGames[d->game].num_actions; }

struct LaunchPlan {
    int blocks_per_sm;
};

int plan_launch(mr_ctx *ctx, const Device &dv, kernel_fn fn, const mr_batch_desc *d, int n_devices, LaunchPlan *lp) {
    (void)d;
    (void)n_devices;
    (void)dv;
    auto it = ctx->occupancy.find((const void *)fn);
    if (it == ctx->occupancy.end()) {
        int bps = 0;
        CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, fn, THREADS_PER_BLOCK, 0));
        if (bps < 1) return fail(ctx, MR_ERR_CUDA, "kernel does not fit on an SM");
        it = ctx->occupancy.emplace((const void *)fn, bps).first;
    }
    lp->blocks_per_sm = it->second;
    return MR_OK;
}

void fill_params(const mr_ctx *ctx, KernelParams &kp, const mr_batch_desc *d, const LaunchPlan &lp, uint32_t trace_cap) {
    memset(&kp, 0, sizeof kp);
    if (policy_uses_mast((int)d->policy)) {
        memcpy(kp.mast_planes, ctx->mast_planes, sizeof kp.mast_planes);
        kp.mast_nbits[0] = ctx->mast_nbits[0];
        kp.mast_nbits[1] = ctx->mast_nbits[1];
        memcpy(kp.mast_rank_small, ctx->mast_rank_small, sizeof kp.mast_rank_small);
    }
    kp.n_leaves = d->n_leaves;
    kp.k = d->playouts_per_leaf;
    (void)lp;
    kp.max_depth = d->max_playout_depth ? d->max_playout_depth : 0xffffffffu;
    kp.seed_lo = (uint32_t)d->seed;
    kp.seed_hi = (uint32_t)(d->seed >> 32);
    for (uint32_t r = 0; r < 10; ++r) {
        kp.philox_keys[r][0] = kp.seed_lo + r * 0x9E3779B9u;
        kp.philox_keys[r][1] = kp.seed_hi + r * 0xBB67AE85u;
    }
    kp.leaf_id_base = d->leaf_id_base;
    // explore iff x1 < floor(eps * 2^32)
    double t = d->epsilon * 4294967296.0;
    uint64_t thr = t <= 0.0 ? 0 : (t >= 4294967296.0 ? (1ull << 32) : (uint64_t)t);
    kp.eps_enable = thr > 0;
    kp.eps_thr_m1 = thr > 0 ? (uint32_t)(thr - 1) : 0;
    kp.max_trace = d->max_trace;
    kp.trace_cap = trace_cap;
    // the per-lane playout log (rollout_kernel.cuh) is a ring with room for a whole playout plus slack, so that a lane's
    // log holds several playouts when it is drained; the LGR / NST kernels index the scratch by ply
    kp.trace_rows = trace_cap;
    if (trace_cap && !policy_is_lgr((int)d->policy) && d->policy != MR_EPS_NST) {
        uint32_t rows = 512;
        while (rows < trace_cap + trace_cap / 2u + 2u) rows <<= 1;
        kp.trace_rows = rows;
    }
    kp.grab_div = 2;
    kp.tables_in_slots = (d->flags & MR_TABLES_IN_SLOTS) ? 1u : 0u;
    kp.amaf_stride = 2u * (uint32_t)table_len(d);
}

int launch_on(mr_ctx *ctx, Device &dv, kernel_fn fn, KernelParams &kp, const LaunchPlan &lp, cudaStream_t stream,
              unsigned long long *d_counter, int scratch_slot, bool reset_counter = true) {
    const uint64_t work = kp.g_end - kp.g_begin;
    if (work == 0) return MR_OK;
    uint32_t blocks = (uint32_t)dv.num_sms * lp.blocks_per_sm;
    const uint64_t need_blocks = (work + 32ull * WARPS_PER_BLOCK - 1) / (32ull * WARPS_PER_BLOCK);
    if (blocks > need_blocks) blocks = (uint32_t)need_blocks;
    kp.total_warps = blocks * WARPS_PER_BLOCK;
    // Guided self-scheduling (profiles/grab_sweep_r01.txt): a warp grabs remaining / (2 * warps) playouts, between 128 and
    // 512 (16 per lane) when the batch is large enough for >= 4 full grabs per resident warp, smaller powers of two
    // otherwise -- large grabs while there is plenty of work, small ones at the end so that the warps finish together.
    {
        uint32_t chunk = 512;
        while (chunk > 32 && (uint64_t)chunk * kp.total_warps * 4 > work) chunk >>= 1;
        kp.grab_max = chunk;
        kp.grab_min = chunk >= 128 ? chunk / 4 : 32;
    }
    static const char *env_min = getenv("MR_GRAB_MIN"), *env_max = getenv("MR_GRAB_MAX"), *env_div = getenv("MR_GRAB_DIV"); // tuning knobs
    if (env_min) kp.grab_min = (uint32_t)atoi(env_min) & ~31u;
    if (env_max) kp.grab_max = (uint32_t)atoi(env_max) & ~31u;
    if (env_div) kp.grab_div = (uint32_t)atoi(env_div);
    if (kp.grab_min < 32) kp.grab_min = 32;
    if (kp.grab_max < kp.grab_min) kp.grab_max = kp.grab_min;
    if (kp.trace_cap) {
        const size_t need = (size_t)blocks * WARPS_PER_BLOCK * kp.trace_rows * 32u;
        CUDA_TRY(ctx, dv.scratch[scratch_slot].reserve(need, false));
        kp.trace_scratch = dv.scratch[scratch_slot].d;
    }
    kp.work_counter = d_counter;
    if (reset_counter) CUDA_TRY(ctx, cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), stream));
    fn<<<blocks, THREADS_PER_BLOCK, 0, stream>>>(kp);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches += 1;
    return MR_OK;
}

} // namespace

extern "C" {

int mr_abi_version(void) { return MR_ABI_VERSION; }

int mr_num_actions(int game) {
    if (game < 0 || game >= MR_NUM_GAMES) return 0;
    return kGames[game].num_actions;
}

const char *mr_last_error(const mr_ctx *ctx) { return ctx ? ctx->error.c_str() : g_global_error.c_str(); }

int mr_init(const int *device_ids, int n_devices, mr_ctx **out) {
    if (!out) return fail(nullptr, MR_ERR_INVALID, "mr_init: out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, MR_ERR_NO_DEVICE, "no CUDA device: %s (this engine has no CPU fallback)",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (n_devices <= 0) n_devices = 1;
    if (n_devices > count && !device_ids)
        return fail(nullptr, MR_ERR_NO_DEVICE, "%d devices requested, %d present", n_devices, count);
    mr_ctx *ctx = new mr_ctx();
    ctx->dev.resize((size_t)n_devices);
    for (int i = 0; i < n_devices; ++i) {
        Device &dv = ctx->dev[(size_t)i];
        dv.id = device_ids ? device_ids[i] : i;
        cudaDeviceProp prop;
        if (dv.id < 0 || dv.id >= count || cudaGetDeviceProperties(&prop, dv.id) != cudaSuccess) {
            int rc = fail(nullptr, MR_ERR_NO_DEVICE, "device %d not usable", dv.id);
            delete ctx;
            return rc;
        }
        if (prop.major != 10) {
            int rc = fail(nullptr, MR_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", dv.id,
                          prop.major, prop.minor);
            delete ctx;
            return rc;
        }
        dv.num_sms = prop.multiProcessorCount;
        cudaSetDevice(dv.id);
        if (cudaStreamCreateWithFlags(&dv.stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaMalloc(&dv.d_counter_ext, sizeof(unsigned long long)) != cudaSuccess || upload_random_best_tables() != cudaSuccess) {
            int rc = fail(nullptr, MR_ERR_CUDA, "stream/counter creation failed on device %d", dv.id);
            mr_destroy(ctx);
            return rc;
        }
        for (int s = 0; s < MR_NUM_SLOTS; ++s) {
            SlotDev &sd = dv.slot[s];
            if (cudaMalloc(&sd.d_counter, sizeof(unsigned long long)) != cudaSuccess || cudaEventCreate(&sd.ev_begin) != cudaSuccess ||
                cudaEventCreate(&sd.ev_end) != cudaSuccess || cudaStreamCreateWithFlags(&sd.stream, cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&sd.ev_done, cudaEventDisableTiming) != cudaSuccess ||
                cudaMalloc(&sd.d_done, sizeof(unsigned int)) != cudaSuccess || cudaMallocHost(&sd.flag_h, 64) != cudaSuccess ||
                cudaMemset(sd.d_done, 0, sizeof(unsigned int)) != cudaSuccess || cudaMemset(sd.d_counter, 0, sizeof(unsigned long long)) != cudaSuccess) {
                int rc = fail(nullptr, MR_ERR_CUDA, "slot setup failed on device %d", dv.id);
                mr_destroy(ctx);
                return rc;
            }
            *sd.flag_h = 0;
            sd.counter_clean = true;
        }
    }
    *out = ctx;
    return MR_OK;
}

void mr_destroy(mr_ctx *ctx) {
    if (!ctx) return;
    for (Device &dv : ctx->dev) {
        cudaSetDevice(dv.id);
        if (dv.stream) cudaStreamSynchronize(dv.stream);
        for (int s = 0; s < MR_NUM_SLOTS; ++s) {
            SlotDev &sd = dv.slot[s];
            if (sd.stream) {
                cudaStreamSynchronize(sd.stream);
                cudaStreamDestroy(sd.stream);
            }
            if (sd.ev_done) cudaEventDestroy(sd.ev_done);
            if (sd.d_done) cudaFree(sd.d_done);
            if (sd.flag_h) cudaFreeHost(sd.flag_h);
            sd.leaves.release();
            sd.results.release();
            sd.amaf.release();
            sd.mast_delta.release();
            sd.trace.release();
            sd.rec.release();
            sd.final_state.release();
            sd.lgr_in.release();
            sd.lgr_stamps.release();
            sd.lgr_last.release();
            sd.lgr2_in.release();
            sd.lgr2_stamps.release();
            sd.lgr2_last_lose.release();
            sd.lgr2_target.release();
            sd.lgr2_removed.release();
            sd.results2.release();
            sd.nst_planes.release();
            sd.nst_nbits.release();
            sd.nst_rank_small.release();
            sd.nst_delta.release();
            if (sd.d_counter) cudaFree(sd.d_counter);
            if (sd.ev_begin) cudaEventDestroy(sd.ev_begin);
            if (sd.ev_end) cudaEventDestroy(sd.ev_end);
        }
        for (auto &sc : dv.scratch) sc.release();
        if (dv.d_counter_ext) cudaFree(dv.d_counter_ext);
        if (dv.stream) cudaStreamDestroy(dv.stream);
    }
    if (ctx->nccl_comm && ctx->nccl_lib) {
        typedef int (*destroy_fn)(void *);
        destroy_fn f = (destroy_fn)dlsym(ctx->nccl_lib, "ncclCommDestroy");
        if (f) f(ctx->nccl_comm);
    }
    delete ctx;
}

int mr_num_devices(const mr_ctx *ctx) { return ctx ? (int)ctx->dev.size() : 0; }

int mr_set_mast(mr_ctx *ctx, int game, const mr_action_stat *mast_in) {
    if (!ctx) return fail(nullptr, MR_ERR_INVALID, "null ctx");
    if (game < 0 || game >= MR_NUM_GAMES || !mast_in) return fail(ctx, MR_ERR_INVALID, "mr_set_mast: bad arguments");
    return compute_mast_planes(ctx, game, mast_in);
}

int mr_set_replies(mr_ctx *ctx, int game, const uint16_t *replies) {
    if (!ctx) return fail(nullptr, MR_ERR_INVALID, "null ctx");
    if (game < 0 || game >= MR_NUM_GAMES) return fail(ctx, MR_ERR_INVALID, "mr_set_replies: unknown game %d", game);
    if (!replies) {
        ctx->replies.clear();
        ctx->replies_game = -1;
        return MR_OK;
    }
    const size_t cnt = (size_t)2 * kGames[game].num_actions;
    ctx->replies.assign(replies, replies + cnt);
    ctx->replies_game = game;
    return MR_OK;
}

int mr_submit(mr_ctx *ctx, int slot, const mr_batch_desc *desc, const mr_leaf *leaves, const mr_action_stat *mast_in) {
    if (!ctx) return fail(nullptr, MR_ERR_INVALID, "null ctx");
    if (slot < 0 || slot >= MR_NUM_SLOTS) return fail(ctx, MR_ERR_INVALID, "slot %d out of range", slot);
    SlotState &ss = ctx->slot[slot];
    if (ss.busy) return fail(ctx, MR_ERR_BUSY, "slot %d already has a batch in flight", slot);
    if (!leaves) return fail(ctx, MR_ERR_INVALID, "null leaves");
    kernel_fn fn = nullptr;
    uint32_t trace_cap = 0;
    int rc = validate(ctx, desc, mast_in != nullptr, &fn, &trace_cap);
    if (rc != MR_OK) return rc;
    if (policy_uses_mast((int)desc->policy)) {
        rc = compute_mast_planes(ctx, (int)desc->game, mast_in);
        if (rc != MR_OK) return rc;
    }
    if (desc->policy == MR_EPS_NST) {
        rc = compute_nst_tables(ctx, (int)desc->game, mast_in);
        if (rc != MR_OK) return rc;
    }
    const int nd = (int)ctx->dev.size();
    LaunchPlan lp;
    CUDA_TRY(ctx, cudaSetDevice(ctx->dev[0].id));
    rc = plan_launch(ctx, ctx->dev[0], fn, desc, nd, &lp);
    if (rc != MR_OK) return rc;
    const int tl = table_len(desc); // entries per player of the AMAF / MAST-delta tables
    const int na = kGames[desc->game].num_actions;
    const uint32_t n = desc->n_leaves, k = desc->playouts_per_leaf;
    const uint64_t total = (uint64_t)n * k;
    // Low-latency path: no parity outputs, no LGR / NST side tables -> one kernel launch whose fused epilogue delivers
    // the results (and the MAST delta) into host-mapped memory.  Small batches also READ their leaves from the
    // host-mapped staging buffer, so nothing but the launch is enqueued.
    const bool fast = !(desc->flags & (MR_WANT_TRACE | MR_WANT_FINAL_STATE)) && !policy_is_lgr((int)desc->policy) && desc->policy != MR_EPS_NST;
    static const uint64_t zero_copy_max = getenv("MR_ZERO_COPY_MAX") ? strtoull(getenv("MR_ZERO_COPY_MAX"), nullptr, 10) : (1ull << 20);
    const bool zero_copy = fast && total <= zero_copy_max;

    for (int di = 0; di < nd; ++di) {
        Device &dv = ctx->dev[(size_t)di];
        SlotDev &sd = dv.slot[slot];
        cudaStream_t stream = sd.stream;
        CUDA_TRY(ctx, cudaSetDevice(dv.id));
        sd.g_begin = total * (uint64_t)di / (uint64_t)nd;
        sd.g_end = total * (uint64_t)(di + 1) / (uint64_t)nd;
        sd.fast = fast;
        sd.wait_event = !fast;
        if (sd.g_end == sd.g_begin) continue; // nothing for this device: mr_wait skips it

        CUDA_TRY(ctx, sd.leaves.reserve(n, true));
        CUDA_TRY(ctx, sd.results.reserve(n, true));
        if (sd.results.fresh) sd.results_clean = false;
        sd.results.fresh = false;
        memcpy(sd.leaves.h, leaves, sizeof(mr_leaf) * n);
        if (!zero_copy) CUDA_TRY(ctx, cudaMemcpyAsync(sd.leaves.d, sd.leaves.h, sizeof(mr_leaf) * n, cudaMemcpyHostToDevice, stream));
        if (fast) {
            if (!sd.results_clean) CUDA_TRY(ctx, cudaMemsetAsync(sd.results.d, 0, sizeof(mr_leaf_result) * sd.results.cap, stream));
            sd.results_clean = true; // the epilogue re-zeroes the entries the batch touched
        } else {
            CUDA_TRY(ctx, cudaMemsetAsync(sd.results.d, 0, sizeof(mr_leaf_result) * n, stream));
            sd.results_clean = false;
        }
        KernelParams kp;
        fill_params(ctx, kp, desc, lp, trace_cap);
        kp.leaves = zero_copy ? sd.leaves.h : sd.leaves.d;
        kp.results = sd.results.d;
        kp.g_begin = sd.g_begin;
        kp.g_end = sd.g_end;
        if (desc->flags & MR_WANT_AMAF) {
            CUDA_TRY(ctx, sd.amaf.reserve((size_t)n * 2 * tl, true));
            if (sd.amaf.fresh) sd.amaf_clean = false;
            sd.amaf.fresh = false;
            if (fast) { // zeroed again behind the copy-out below, off the next batch's critical path
                if (!sd.amaf_clean) CUDA_TRY(ctx, cudaMemsetAsync(sd.amaf.d, 0, sizeof(mr_action_stat) * sd.amaf.cap, stream));
                sd.amaf_clean = true;
            } else {
                CUDA_TRY(ctx, cudaMemsetAsync(sd.amaf.d, 0, sizeof(mr_action_stat) * n * 2 * tl, stream));
                sd.amaf_clean = false;
            }
            kp.amaf = sd.amaf.d;
        }
        if (desc->flags & MR_WANT_MAST_DELTA) {
            CUDA_TRY(ctx, sd.mast_delta.reserve((size_t)2 * na, true)); // (always the dense size: the buffer serves both forms)
            if (sd.mast_delta.fresh) sd.mast_clean = false;
            sd.mast_delta.fresh = false;
            if (fast) {
                if (!sd.mast_clean) CUDA_TRY(ctx, cudaMemsetAsync(sd.mast_delta.d, 0, sizeof(mr_action_stat) * sd.mast_delta.cap, stream));
                sd.mast_clean = true;
            } else {
                CUDA_TRY(ctx, cudaMemsetAsync(sd.mast_delta.d, 0, sizeof(mr_action_stat) * 2 * tl, stream));
                sd.mast_clean = false;
            }
            kp.mast_delta = sd.mast_delta.d;
        }
        if (desc->flags & MR_WANT_TRACE) {
            CUDA_TRY(ctx, sd.trace.reserve((size_t)total * desc->max_trace, true));
            CUDA_TRY(ctx, sd.rec.reserve((size_t)total, true));
            CUDA_TRY(ctx, cudaMemsetAsync(sd.trace.d, 0, sizeof(uint16_t) * total * desc->max_trace, stream));
            kp.trace = sd.trace.d;
            kp.rec = sd.rec.d;
        }
        if (desc->flags & MR_WANT_FINAL_STATE) {
            CUDA_TRY(ctx, sd.final_state.reserve((size_t)total, true));
            kp.final_state = sd.final_state.d;
        }
        if (policy_is_lgr((int)desc->policy)) {
            const size_t cnt = (size_t)2 * na;
            CUDA_TRY(ctx, sd.lgr_stamps.reserve(cnt, true));
            CUDA_TRY(ctx, sd.lgr_last.reserve((size_t)2 * n, true));
            CUDA_TRY(ctx, cudaMemsetAsync(sd.lgr_stamps.d, 0, sizeof(unsigned long long) * cnt, stream));
            CUDA_TRY(ctx, cudaMemsetAsync(sd.lgr_last.d, 0, sizeof(uint32_t) * 2 * n, stream));
            kp.lgr_stamps = sd.lgr_stamps.d;
            kp.lgr_last = sd.lgr_last.d;
            if (!ctx->replies.empty()) {
                CUDA_TRY(ctx, sd.lgr_in.reserve(cnt, true));
                memcpy(sd.lgr_in.h, ctx->replies.data(), sizeof(uint16_t) * cnt);
                CUDA_TRY(ctx, cudaMemcpyAsync(sd.lgr_in.d, sd.lgr_in.h, sizeof(uint16_t) * cnt, cudaMemcpyHostToDevice, stream));
                kp.lgr_in = sd.lgr_in.d;
            }
        }
        if (policy_is_lgr2((int)desc->policy)) {
            const size_t S = (size_t)kSlots[desc->game], cnt2 = 2 * S * S;
            CUDA_TRY(ctx, sd.lgr2_stamps.reserve(cnt2, true));
            CUDA_TRY(ctx, sd.lgr2_last_lose.reserve((size_t)2 * n, true));
            CUDA_TRY(ctx, cudaMemsetAsync(sd.lgr2_stamps.d, 0, sizeof(unsigned long long) * cnt2, stream));
            CUDA_TRY(ctx, cudaMemsetAsync(sd.lgr2_last_lose.d, 0, sizeof(uint32_t) * 2 * n, stream));
            kp.lgr2_stamps = sd.lgr2_stamps.d;
            kp.lgr2_last_lose = sd.lgr2_last_lose.d;
            if (!ctx->replies2.empty()) {
                CUDA_TRY(ctx, sd.lgr2_in.reserve(cnt2, true));
                memcpy(sd.lgr2_in.h, ctx->replies2.data(), sizeof(uint16_t) * cnt2);
                CUDA_TRY(ctx, cudaMemcpyAsync(sd.lgr2_in.d, sd.lgr2_in.h, sizeof(uint16_t) * cnt2, cudaMemcpyHostToDevice, stream));
                kp.lgr2_in = sd.lgr2_in.d;
            }
        }
        if (desc->policy == MR_EPS_NST) {
            const size_t S = (size_t)kSlots[desc->game];
            CUDA_TRY(ctx, sd.nst_planes.reserve(ctx->nst_planes.size(), true));
            CUDA_TRY(ctx, sd.nst_nbits.reserve(ctx->nst_nbits.size(), true));
            CUDA_TRY(ctx, sd.nst_rank_small.reserve(ctx->nst_rank_small.size(), true));
            CUDA_TRY(ctx, sd.nst_delta.reserve(2 * S * S, true));
            memcpy(sd.nst_planes.h, ctx->nst_planes.data(), sizeof(uint64_t) * ctx->nst_planes.size());
            memcpy(sd.nst_nbits.h, ctx->nst_nbits.data(), sizeof(uint32_t) * ctx->nst_nbits.size());
            memcpy(sd.nst_rank_small.h, ctx->nst_rank_small.data(), sizeof(uint16_t) * ctx->nst_rank_small.size());
            CUDA_TRY(ctx, cudaMemcpyAsync(sd.nst_planes.d, sd.nst_planes.h, sizeof(uint64_t) * ctx->nst_planes.size(), cudaMemcpyHostToDevice, stream));
            CUDA_TRY(ctx, cudaMemcpyAsync(sd.nst_nbits.d, sd.nst_nbits.h, sizeof(uint32_t) * ctx->nst_nbits.size(), cudaMemcpyHostToDevice, stream));
            CUDA_TRY(ctx, cudaMemcpyAsync(sd.nst_rank_small.d, sd.nst_rank_small.h, sizeof(uint16_t) * ctx->nst_rank_small.size(), cudaMemcpyHostToDevice, stream));
            CUDA_TRY(ctx, cudaMemsetAsync(sd.nst_delta.d, 0, sizeof(unsigned long long) * 2 * S * S, stream));
            kp.nst_planes = sd.nst_planes.d;
            kp.nst_nbits = sd.nst_nbits.d;
            kp.nst_rank_small = sd.nst_rank_small.d;
            kp.nst_delta = sd.nst_delta.d;
        }
        if (fast) {
            kp.done_counter = sd.d_done;
            kp.host_results = sd.results.h;
            kp.host_mast_delta = (desc->flags & MR_WANT_MAST_DELTA) ? sd.mast_delta.h : nullptr;
            kp.mast_delta_len = (desc->flags & MR_WANT_MAST_DELTA) ? 2u * (uint32_t)tl : 0u;
            kp.done_seq = ++sd.seq;
            kp.host_flag = sd.flag_h;
        }
        if (ctx->timing) CUDA_TRY(ctx, cudaEventRecord(sd.ev_begin, stream));
        rc = launch_on(ctx, dv, fn, kp, lp, stream, sd.d_counter, slot, !(fast && sd.counter_clean));
        if (rc != MR_OK) return rc;
        sd.counter_clean = fast; // the epilogue resets the work counter; the other kernels leave it at the batch size
        if (ctx->timing) CUDA_TRY(ctx, cudaEventRecord(sd.ev_end, stream));
        sd.kp_saved = kp;

        if (fast) {
            if (desc->flags & MR_WANT_AMAF) { // the one large output: copied out (and zeroed again) behind the kernel
                CUDA_TRY(ctx, cudaMemcpyAsync(sd.amaf.h, sd.amaf.d, sizeof(mr_action_stat) * n * 2 * tl, cudaMemcpyDeviceToHost, stream));
                CUDA_TRY(ctx, cudaMemsetAsync(sd.amaf.d, 0, sizeof(mr_action_stat) * n * 2 * tl, stream));
                CUDA_TRY(ctx, cudaEventRecord(sd.ev_done, stream));
                sd.wait_event = true;
            }
            continue;
        }
        // results come back on the same stream
        CUDA_TRY(ctx, cudaMemcpyAsync(sd.results.h, sd.results.d, sizeof(mr_leaf_result) * n, cudaMemcpyDeviceToHost, stream));
        if (desc->flags & MR_WANT_AMAF)
            CUDA_TRY(ctx, cudaMemcpyAsync(sd.amaf.h, sd.amaf.d, sizeof(mr_action_stat) * n * 2 * tl, cudaMemcpyDeviceToHost, stream));
        if (desc->flags & MR_WANT_MAST_DELTA)
            CUDA_TRY(ctx, cudaMemcpyAsync(sd.mast_delta.h, sd.mast_delta.d, sizeof(mr_action_stat) * 2 * tl, cudaMemcpyDeviceToHost, stream));
        if (policy_is_lgr2((int)desc->policy)) {
            const size_t S = (size_t)kSlots[desc->game];
            CUDA_TRY(ctx, cudaMemcpyAsync(sd.lgr2_stamps.h, sd.lgr2_stamps.d, sizeof(unsigned long long) * 2 * S * S, cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(ctx, cudaMemcpyAsync(sd.lgr2_last_lose.h, sd.lgr2_last_lose.d, sizeof(uint32_t) * 2 * n, cudaMemcpyDeviceToHost, stream));
        }
        if (policy_is_lgr((int)desc->policy)) {
            CUDA_TRY(ctx, cudaMemcpyAsync(sd.lgr_stamps.h, sd.lgr_stamps.d, sizeof(unsigned long long) * 2 * na, cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(ctx, cudaMemcpyAsync(sd.lgr_last.h, sd.lgr_last.d, sizeof(uint32_t) * 2 * n, cudaMemcpyDeviceToHost, stream));
        }
        if (desc->policy == MR_EPS_NST) {
            const size_t S = (size_t)kSlots[desc->game];
            CUDA_TRY(ctx, cudaMemcpyAsync(sd.nst_delta.h, sd.nst_delta.d, sizeof(unsigned long long) * 2 * S * S, cudaMemcpyDeviceToHost, stream));
        }
        const uint64_t gb = sd.g_begin, ge = sd.g_end;
        if (desc->flags & MR_WANT_TRACE) {
            CUDA_TRY(ctx, cudaMemcpyAsync(sd.trace.h + gb * desc->max_trace, sd.trace.d + gb * desc->max_trace,
                                          sizeof(uint16_t) * (ge - gb) * desc->max_trace, cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(ctx, cudaMemcpyAsync(sd.rec.h + gb, sd.rec.d + gb, sizeof(mr_playout_record) * (ge - gb), cudaMemcpyDeviceToHost, stream));
        }
        if (desc->flags & MR_WANT_FINAL_STATE)
            CUDA_TRY(ctx, cudaMemcpyAsync(sd.final_state.h + gb, sd.final_state.d + gb, sizeof(mr_leaf) * (ge - gb), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(ctx, cudaEventRecord(sd.ev_done, stream));
    }
    ss.busy = true;
    ss.lgr_ready = false;
    ss.nst_ready = false;
    ss.lgr2_ready = false;
    ss.trace_cap = trace_cap;
    ss.desc = *desc;
    return MR_OK;
}

// Waits for one device's share of the slot's batch.  Fast path: spin on the host-mapped flag the fused epilogue raises
// (a PCIe write away from the kernel's last instruction); otherwise poll the slot's completion event.  Either way no
// other slot's work is waited for.
static int wait_slot_device(mr_ctx *ctx, Device &dv, SlotDev &sd) {
    CUDA_TRY(ctx, cudaSetDevice(dv.id));
    if (sd.wait_event) {
        for (;;) {
            cudaError_t e = cudaEventQuery(sd.ev_done);
            if (e == cudaSuccess) break;
            if (e != cudaErrorNotReady) return fail(ctx, MR_ERR_CUDA, "batch failed: %s", cudaGetErrorString(e));
            __builtin_ia32_pause();
        }
        return MR_OK;
    }
    const uint32_t want = sd.seq;
    for (uint64_t spins = 0;; ++spins) {
        if (__atomic_load_n(sd.flag_h, __ATOMIC_ACQUIRE) == want) return MR_OK;
        __builtin_ia32_pause();
        if ((spins & 0xfffffu) == 0xfffffu) { // every ~1M spins: a faulted kernel never raises the flag
            cudaError_t e = cudaStreamQuery(sd.stream);
            if (e != cudaSuccess && e != cudaErrorNotReady) return fail(ctx, MR_ERR_CUDA, "batch failed: %s", cudaGetErrorString(e));
            if (e == cudaSuccess && __atomic_load_n(sd.flag_h, __ATOMIC_ACQUIRE) != want)
                return fail(ctx, MR_ERR_CUDA, "batch finished without publishing its results (flag %u, expected %u)", *sd.flag_h, want);
        }
    }
}

int mr_wait(mr_ctx *ctx, int slot, mr_leaf_result *out, mr_action_stat *amaf_out, mr_action_stat *mast_delta_out,
            uint16_t *trace_out, mr_playout_record *rec_out, mr_leaf *final_out) {
    if (!ctx) return fail(nullptr, MR_ERR_INVALID, "null ctx");
    if (slot < 0 || slot >= MR_NUM_SLOTS) return fail(ctx, MR_ERR_INVALID, "slot %d out of range", slot);
    SlotState &ss = ctx->slot[slot];
    if (!ss.busy) return fail(ctx, MR_ERR_INVALID, "slot %d has no batch in flight", slot);
    if (!out) return fail(ctx, MR_ERR_INVALID, "null result buffer");
    const mr_batch_desc &d = ss.desc;
    const int na = table_len(&d); // entries per player of the two tables (dense ids or slots)
    const uint32_t n = d.n_leaves;
    ss.busy = false; // whatever happens below, the slot is released
    float worst_ms = 0.f;
    for (Device &dv : ctx->dev) {
        SlotDev &sd = dv.slot[slot];
        if (sd.g_end == sd.g_begin) continue;
        int rc = wait_slot_device(ctx, dv, sd);
        if (rc != MR_OK) return rc;
        if (ctx->timing) {
            float ms = 0.f;
            CUDA_TRY(ctx, cudaEventSynchronize(sd.ev_end)); // (the fused epilogue raises its flag before the event completes)
            CUDA_TRY(ctx, cudaEventElapsedTime(&ms, sd.ev_begin, sd.ev_end));
            worst_ms = std::max(worst_ms, ms);
        }
    }
    ss.kernel_ms = worst_ms;
    ss.lgr_ready = policy_is_lgr((int)d.policy);
    ss.lgr2_ready = policy_is_lgr2((int)d.policy);
    ss.nst_ready = d.policy == MR_EPS_NST;
    const bool single = ctx->dev.size() == 1; // one device: its tables ARE the totals, a copy instead of zero-and-add
    if (!single) {
        memset(out, 0, sizeof(mr_leaf_result) * n);
        if (amaf_out && (d.flags & MR_WANT_AMAF)) memset(amaf_out, 0, sizeof(mr_action_stat) * (size_t)n * 2 * na);
        if (mast_delta_out && (d.flags & MR_WANT_MAST_DELTA)) memset(mast_delta_out, 0, sizeof(mr_action_stat) * 2 * na);
    }
    for (Device &dv : ctx->dev) {
        SlotDev &sd = dv.slot[slot];
        if (sd.g_end == sd.g_begin) continue;
        if (single) {
            memcpy(out, sd.results.h, sizeof(mr_leaf_result) * n);
            if (amaf_out && (d.flags & MR_WANT_AMAF)) memcpy(amaf_out, sd.amaf.h, sizeof(mr_action_stat) * (size_t)n * 2 * na);
            if (mast_delta_out && (d.flags & MR_WANT_MAST_DELTA)) memcpy(mast_delta_out, sd.mast_delta.h, sizeof(mr_action_stat) * 2 * na);
        } else {
            for (uint32_t i = 0; i < n; ++i) {
                out[i].wins[0] += sd.results.h[i].wins[0];
                out[i].wins[1] += sd.results.h[i].wins[1];
                out[i].draws += sd.results.h[i].draws;
                out[i].cutoffs += sd.results.h[i].cutoffs;
                out[i].sum_depth += sd.results.h[i].sum_depth;
            }
            if (amaf_out && (d.flags & MR_WANT_AMAF)) {
                const size_t cnt = (size_t)n * 2 * na;
                for (size_t i = 0; i < cnt; ++i) {
                    amaf_out[i].visits += sd.amaf.h[i].visits;
                    amaf_out[i].score += sd.amaf.h[i].score;
                }
            }
            if (mast_delta_out && (d.flags & MR_WANT_MAST_DELTA)) {
                for (int i = 0; i < 2 * na; ++i) {
                    mast_delta_out[i].visits += sd.mast_delta.h[i].visits;
                    mast_delta_out[i].score += sd.mast_delta.h[i].score;
                }
            }
        }
        const uint64_t gb = sd.g_begin, ge = sd.g_end;
        if (ge > gb) {
            if (d.flags & MR_WANT_TRACE) {
                if (trace_out)
                    memcpy(trace_out + gb * d.max_trace, sd.trace.h + gb * d.max_trace, sizeof(uint16_t) * (ge - gb) * d.max_trace);
                if (rec_out) memcpy(rec_out + gb, sd.rec.h + gb, sizeof(mr_playout_record) * (ge - gb));
            }
            if ((d.flags & MR_WANT_FINAL_STATE) && final_out) memcpy(final_out + gb, sd.final_state.h + gb, sizeof(mr_leaf) * (ge - gb));
        }
    }
    return MR_OK;
}

int mr_reply_updates(mr_ctx *ctx, int slot, mr_reply_update *updates_out, uint32_t *last_win_out) {
    if (!ctx) return fail(nullptr, MR_ERR_INVALID, "null ctx");
    if (slot < 0 || slot >= MR_NUM_SLOTS) return fail(ctx, MR_ERR_INVALID, "slot %d out of range", slot);
    SlotState &ss = ctx->slot[slot];
    if (ss.busy || !ss.lgr_ready)
        return fail(ctx, MR_ERR_INVALID, "slot %d holds no completed MR_EPS_LGR / MR_EPS_LGR2 batch (call mr_wait first)", slot);
    const mr_batch_desc &d = ss.desc;
    const int game = (int)d.game, na = kGames[game].num_actions;
    const bool through = game == MR_BREAKTHROUGH || game == MR_KNIGHTTHROUGH;
    if (updates_out) {
        for (int i = 0; i < 2 * na; ++i) {
            unsigned long long best = 0;
            for (Device &dv : ctx->dev) {
                SlotDev &sd = dv.slot[slot];
                if (sd.g_end > sd.g_begin) best = std::max(best, sd.lgr_stamps.h[i]);
            }
            mr_reply_update &u = updates_out[i];
            memset(&u, 0, sizeof u);
            if (best) {
                const uint32_t id = (uint32_t)(best & 0x1fffull);
                u.order = best >> 13;
                u.action_plus1 = (uint16_t)((through ? ((id >> 6) << 8) | (id & 63u) : id) + 1u);
            }
        }
    }
    if (last_win_out) {
        for (uint32_t i = 0; i < 2 * d.n_leaves; ++i) {
            uint32_t best = 0;
            for (Device &dv : ctx->dev) {
                SlotDev &sd = dv.slot[slot];
                if (sd.g_end > sd.g_begin) best = std::max(best, sd.lgr_last.h[i]);
            }
            last_win_out[i] = best;
        }
    }
    return MR_OK;
}

int mr_set_replies2(mr_ctx *ctx, int game, const uint16_t *replies2) {
    if (!ctx) return fail(nullptr, MR_ERR_INVALID, "null ctx");
    if (game < 0 || game >= MR_NUM_GAMES) return fail(ctx, MR_ERR_INVALID, "mr_set_replies2: unknown game %d", game);
    if (!replies2) {
        ctx->replies2.clear();
        ctx->replies2_game = -1;
        return MR_OK;
    }
    const size_t S = (size_t)kSlots[game];
    ctx->replies2.assign(replies2, replies2 + 2 * S * S);
    ctx->replies2_game = game;
    return MR_OK;
}

int mr_reply2_inserts(mr_ctx *ctx, int slot, mr_reply_update *inserts_out, uint32_t *last_lose_out) {
    if (!ctx) return fail(nullptr, MR_ERR_INVALID, "null ctx");
    if (slot < 0 || slot >= MR_NUM_SLOTS) return fail(ctx, MR_ERR_INVALID, "slot %d out of range", slot);
    SlotState &ss = ctx->slot[slot];
    if (ss.busy || !ss.lgr2_ready)
        return fail(ctx, MR_ERR_INVALID, "slot %d holds no completed MR_EPS_LGR2 batch (call mr_wait first)", slot);
    const mr_batch_desc &d = ss.desc;
    const int game = (int)d.game;
    const size_t S = (size_t)kSlots[game];
    if (inserts_out) {
        for (size_t i = 0; i < 2 * S * S; ++i) {
            unsigned long long best = 0;
            for (Device &dv : ctx->dev) {
                SlotDev &sd = dv.slot[slot];
                if (sd.g_end > sd.g_begin) best = std::max(best, sd.lgr2_stamps.h[i]);
            }
            mr_reply_update &u = inserts_out[i];
            memset(&u, 0, sizeof u);
            if (best) {
                const int pl = i >= S * S ? 1 : 0;
                u.order = best >> 13;
                u.action_plus1 = (uint16_t)(slot_action(game, pl, (int)(best & 0x1fffull)) + 1);
            }
        }
    }
    if (last_lose_out) {
        for (uint32_t i = 0; i < 2 * d.n_leaves; ++i) {
            uint32_t best = 0;
            for (Device &dv : ctx->dev) {
                SlotDev &sd = dv.slot[slot];
                if (sd.g_end > sd.g_begin) best = std::max(best, sd.lgr2_last_lose.h[i]);
            }
            last_lose_out[i] = best;
        }
    }
    return MR_OK;
}

int mr_reply2_resolve(mr_ctx *ctx, int slot, const mr_reply_update *merged, uint8_t *removed_out) {
    if (!ctx) return fail(nullptr, MR_ERR_INVALID, "null ctx");
    if (slot < 0 || slot >= MR_NUM_SLOTS) return fail(ctx, MR_ERR_INVALID, "slot %d out of range", slot);
    SlotState &ss = ctx->slot[slot];
    if (ss.busy || !ss.lgr2_ready)
        return fail(ctx, MR_ERR_INVALID, "slot %d holds no completed MR_EPS_LGR2 batch (call mr_wait first)", slot);
    if (!merged || !removed_out) return fail(ctx, MR_ERR_INVALID, "mr_reply2_resolve: null buffer");
    const mr_batch_desc &d = ss.desc;
    const int game = (int)d.game;
    const size_t S = (size_t)kSlots[game], cnt = 2 * S * S;
    // candidate entry per context: the caller's last insert (order > 0), or the entry its table holds now (order == 0,
    // removable by any losing playout of the batch)
    std::vector<unsigned long long> target(cnt, 0ull);
    for (size_t i = 0; i < cnt; ++i) {
        if (!merged[i].action_plus1) continue;
        const int pl = i >= S * S ? 1 : 0;
        const int a = action_slot(game, pl, (uint32_t)merged[i].action_plus1 - 1u);
        if (a < 0) return fail(ctx, MR_ERR_INVALID, "mr_reply2_resolve: merged[%zu] names no action of player %d", i, pl);
        target[i] = ((unsigned long long)(merged[i].order >> 13) << 16) | (unsigned long long)(a + 1);
    }
    kernel_fn fn = pick_kernel(game, (int)d.policy, 0, 0);
    if (!fn) return fail(ctx, MR_ERR_UNSUPPORTED, "MR_EPS_LGR2 is not built for game %d", game);
    memset(removed_out, 0, cnt);
    for (Device &dv : ctx->dev) {
        SlotDev &sd = dv.slot[slot];
        if (sd.g_end == sd.g_begin) continue;
        CUDA_TRY(ctx, cudaSetDevice(dv.id));
        LaunchPlan lp;
        int rc = plan_launch(ctx, dv, fn, &d, 1, &lp);
        if (rc != MR_OK) return rc;
        CUDA_TRY(ctx, sd.lgr2_target.reserve(cnt, true));
        CUDA_TRY(ctx, sd.lgr2_removed.reserve(cnt, true));
        CUDA_TRY(ctx, sd.results2.reserve(d.n_leaves, false));
        memcpy(sd.lgr2_target.h, target.data(), sizeof(unsigned long long) * cnt);
        CUDA_TRY(ctx, cudaMemcpyAsync(sd.lgr2_target.d, sd.lgr2_target.h, sizeof(unsigned long long) * cnt, cudaMemcpyHostToDevice, sd.stream));
        CUDA_TRY(ctx, cudaMemsetAsync(sd.lgr2_removed.d, 0, cnt, sd.stream));
        CUDA_TRY(ctx, cudaMemsetAsync(sd.results2.d, 0, sizeof(mr_leaf_result) * d.n_leaves, sd.stream));
        // the same playouts again (same leaves, tables, seed and range), with every other output switched off
        KernelParams kp = sd.kp_saved;
        kp.results = sd.results2.d;
        kp.amaf = nullptr;
        kp.mast_delta = nullptr;
        kp.trace = nullptr;
        kp.rec = nullptr;
        kp.final_state = nullptr;
        kp.lgr_stamps = nullptr;
        kp.lgr_last = nullptr;
        kp.lgr2_stamps = nullptr;
        kp.lgr2_last_lose = nullptr;
        kp.lgr2_target = sd.lgr2_target.d;
        kp.lgr2_removed = sd.lgr2_removed.d;
        rc = launch_on(ctx, dv, fn, kp, lp, sd.stream, sd.d_counter, slot);
        sd.counter_clean = false;
        if (rc != MR_OK) return rc;
        CUDA_TRY(ctx, cudaMemcpyAsync(sd.lgr2_removed.h, sd.lgr2_removed.d, cnt, cudaMemcpyDeviceToHost, sd.stream));
    }
    for (Device &dv : ctx->dev) {
        SlotDev &sd = dv.slot[slot];
        if (sd.g_end == sd.g_begin) continue;
        CUDA_TRY(ctx, cudaSetDevice(dv.id));
        CUDA_TRY(ctx, cudaStreamSynchronize(sd.stream));
        for (size_t i = 0; i < cnt; ++i) removed_out[i] |= sd.lgr2_removed.h[i];
    }
    return MR_OK;
}

int mr_num_slots(int game) { return (game < 0 || game >= MR_NUM_GAMES) ? 0 : kSlots[game]; }

int mr_action_slot(int game, int player, uint16_t code) {
    if (game < 0 || game >= MR_NUM_GAMES || player < 0 || player > 1) return -1;
    return action_slot(game, player, code);
}

int mr_slot_action(int game, int player, int slot) {
    if (game < 0 || game >= MR_NUM_GAMES || player < 0 || player > 1) return -1;
    return slot_action(game, player, slot);
}

int mr_set_bigrams(mr_ctx *ctx, int game, const mr_action_stat *bigram, uint32_t backoff_threshold) {
    if (!ctx) return fail(nullptr, MR_ERR_INVALID, "null ctx");
    if (game < 0 || game >= MR_NUM_GAMES) return fail(ctx, MR_ERR_INVALID, "mr_set_bigrams: unknown game %d", game);
    ctx->nst_threshold = backoff_threshold;
    if (!bigram) {
        ctx->bigram.clear();
        ctx->bigram_game = -1;
        return MR_OK;
    }
    const size_t S = (size_t)kSlots[game];
    ctx->bigram.assign(bigram, bigram + 2 * S * S);
    ctx->bigram_game = game;
    return MR_OK;
}

int mr_bigram_delta(mr_ctx *ctx, int slot, mr_action_stat *bigram_delta_out) {
    if (!ctx) return fail(nullptr, MR_ERR_INVALID, "null ctx");
    if (slot < 0 || slot >= MR_NUM_SLOTS) return fail(ctx, MR_ERR_INVALID, "slot %d out of range", slot);
    SlotState &ss = ctx->slot[slot];
    if (ss.busy || !ss.nst_ready)
        return fail(ctx, MR_ERR_INVALID, "slot %d holds no completed MR_EPS_NST batch (call mr_wait first)", slot);
    if (!bigram_delta_out) return fail(ctx, MR_ERR_INVALID, "null bigram delta buffer");
    const size_t S = (size_t)kSlots[ss.desc.game], cnt = 2 * S * S;
    memset(bigram_delta_out, 0, sizeof(mr_action_stat) * cnt);
    for (Device &dv : ctx->dev) {
        SlotDev &sd = dv.slot[slot];
        if (sd.g_end == sd.g_begin) continue;
        for (size_t i = 0; i < cnt; ++i) {
            const unsigned long long v = sd.nst_delta.h[i];
            bigram_delta_out[i].visits += (uint32_t)v;
            bigram_delta_out[i].score += (int32_t)(uint32_t)(v >> 32);
        }
    }
    return MR_OK;
}

int mr_playout_batch(mr_ctx *ctx, const mr_batch_desc *desc, const mr_leaf *leaves, const mr_action_stat *mast_in,
                     mr_leaf_result *out, mr_action_stat *amaf_out, mr_action_stat *mast_delta_out, uint16_t *trace_out,
                     mr_playout_record *rec_out, mr_leaf *final_out) {
    int rc = mr_submit(ctx, 0, desc, leaves, mast_in);
    if (rc != MR_OK) return rc;
    return mr_wait(ctx, 0, out, amaf_out, mast_delta_out, trace_out, rec_out, final_out);
}

int mr_launch_device(mr_ctx *ctx, int device_index, const mr_batch_desc *desc, const void *d_leaves, void *d_results,
                     void *d_amaf, void *d_mast_delta, void *stream) {
    if (!ctx) return fail(nullptr, MR_ERR_INVALID, "null ctx");
    if (device_index < 0 || device_index >= (int)ctx->dev.size()) return fail(ctx, MR_ERR_INVALID, "device index %d out of range", device_index);
    if (!d_leaves || !d_results) return fail(ctx, MR_ERR_INVALID, "null device buffer");
    Device &dv = ctx->dev[(size_t)device_index];
    if (!desc) return fail(ctx, MR_ERR_INVALID, "null batch descriptor");
    if (desc->flags & (MR_WANT_TRACE | MR_WANT_FINAL_STATE))
        return fail(ctx, MR_ERR_INVALID, "trace / final-state outputs are not available on the device-resident path");
    if (policy_is_lgr((int)desc->policy)) return fail(ctx, MR_ERR_UNSUPPORTED, "MR_EPS_LGR / MR_EPS_LGR2 are not available on the device-resident path");
    if (desc->policy == MR_EPS_NST) return fail(ctx, MR_ERR_UNSUPPORTED, "MR_EPS_NST is not available on the device-resident path");
    kernel_fn fn = nullptr;
    uint32_t trace_cap = 0;
    int rc = validate(ctx, desc, ctx->mast_game == (int)desc->game, &fn, &trace_cap);
    if (rc != MR_OK) return rc;
    if ((desc->flags & MR_WANT_AMAF) && !d_amaf) return fail(ctx, MR_ERR_INVALID, "MR_WANT_AMAF without d_amaf");
    if ((desc->flags & MR_WANT_MAST_DELTA) && !d_mast_delta) return fail(ctx, MR_ERR_INVALID, "MR_WANT_MAST_DELTA without d_mast_delta");
    CUDA_TRY(ctx, cudaSetDevice(dv.id));
    LaunchPlan lp;
    rc = plan_launch(ctx, dv, fn, desc, 1, &lp);
    if (rc != MR_OK) return rc;
    KernelParams kp;
    fill_params(ctx, kp, desc, lp, trace_cap);
    kp.leaves = (const mr_leaf *)d_leaves;
    kp.results = (mr_leaf_result *)d_results;
    kp.amaf = (mr_action_stat *)d_amaf;
    kp.mast_delta = (mr_action_stat *)d_mast_delta;
    kp.g_begin = 0;
    kp.g_end = (uint64_t)desc->n_leaves * desc->playouts_per_leaf;
    cudaStream_t st = stream ? (cudaStream_t)stream : dv.stream;
    return launch_on(ctx, dv, fn, kp, lp, st, dv.d_counter_ext, MR_NUM_SLOTS);
}

int mr_last_kernel_ms(mr_ctx *ctx, int slot, float *ms) {
    if (!ctx || !ms || slot < 0 || slot >= MR_NUM_SLOTS) return fail(ctx, MR_ERR_INVALID, "mr_last_kernel_ms: bad arguments");
    *ms = ctx->slot[slot].kernel_ms;
    return MR_OK;
}

uint64_t mr_kernel_launches(const mr_ctx *ctx) { return ctx ? ctx->launches : 0; }

int mr_set_kernel_timing(mr_ctx *ctx, int enabled) {
    if (!ctx) return fail(nullptr, MR_ERR_INVALID, "null ctx");
    ctx->timing = enabled != 0;
    return MR_OK;
}

int mr_allreduce_root(mr_ctx *ctx, int32_t *const *bufs, int n_bufs, size_t n) {
    if (!ctx || !bufs || n_bufs <= 0) return fail(ctx, MR_ERR_INVALID, "mr_allreduce_root: bad arguments");
    std::vector<int64_t> sum(n, 0);
    for (int b = 0; b < n_bufs; ++b) {
        if (!bufs[b]) return fail(ctx, MR_ERR_INVALID, "mr_allreduce_root: null buffer %d", b);
        for (size_t i = 0; i < n; ++i) sum[i] += bufs[b][i];
    }
    for (int b = 0; b < n_bufs; ++b)
        for (size_t i = 0; i < n; ++i) bufs[b][i] = (int32_t)sum[i];
    return MR_OK;
}

// ---- NCCL, loaded at run time so the library carries no link-time dependency on it ----
namespace {
typedef struct {
    char internal[MR_COMM_ID_BYTES];
} nccl_uid;
typedef int (*nccl_get_uid_fn)(nccl_uid *);
typedef int (*nccl_init_rank_fn)(void **, int, nccl_uid, int);
typedef int (*nccl_allreduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef const char *(*nccl_errstr_fn)(int);

void *open_nccl() {
    // prefer the copy already mapped into the process (torch's), then the system one
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        void *h = dlopen(nm, RTLD_NOW | RTLD_NOLOAD);
        if (h) return h;
    }
    for (const char *nm : names) {
        void *h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) return h;
    }
    return nullptr;
}
} // namespace

int mr_comm_unique_id(uint8_t id[MR_COMM_ID_BYTES]) {
    void *lib = open_nccl();
    if (!lib) return fail(nullptr, MR_ERR_COMM, "libnccl.so.2 not found");
    nccl_get_uid_fn f = (nccl_get_uid_fn)dlsym(lib, "ncclGetUniqueId");
    if (!f) return fail(nullptr, MR_ERR_COMM, "ncclGetUniqueId missing");
    nccl_uid u;
    int rc = f(&u);
    if (rc != 0) return fail(nullptr, MR_ERR_COMM, "ncclGetUniqueId failed: %d", rc);
    memcpy(id, u.internal, MR_COMM_ID_BYTES);
    return MR_OK;
}

int mr_comm_init(mr_ctx *ctx, const uint8_t id[MR_COMM_ID_BYTES], int rank, int world_size) {
    if (!ctx || !id || rank < 0 || rank >= world_size) return fail(ctx, MR_ERR_INVALID, "mr_comm_init: bad arguments");
    if (ctx->dev.size() != 1) return fail(ctx, MR_ERR_INVALID, "mr_comm_init: one device per rank");
    ctx->nccl_lib = open_nccl();
    if (!ctx->nccl_lib) return fail(ctx, MR_ERR_COMM, "libnccl.so.2 not found");
    nccl_init_rank_fn f = (nccl_init_rank_fn)dlsym(ctx->nccl_lib, "ncclCommInitRank");
    if (!f) return fail(ctx, MR_ERR_COMM, "ncclCommInitRank missing");
    nccl_uid u;
    memcpy(u.internal, id, MR_COMM_ID_BYTES);
    CUDA_TRY(ctx, cudaSetDevice(ctx->dev[0].id));
    int rc = f(&ctx->nccl_comm, world_size, u, rank);
    if (rc != 0) return fail(ctx, MR_ERR_COMM, "ncclCommInitRank failed: %d", rc);
    ctx->comm_rank = rank;
    ctx->comm_world = world_size;
    return MR_OK;
}

int mr_comm_allreduce_i32(mr_ctx *ctx, void *d_buf, size_t n, void *stream) {
    if (!ctx || !d_buf) return fail(ctx, MR_ERR_INVALID, "mr_comm_allreduce_i32: bad arguments");
    if (!ctx->nccl_comm) return fail(ctx, MR_ERR_COMM, "communicator not initialised (mr_comm_init)");
    nccl_allreduce_fn f = (nccl_allreduce_fn)dlsym(ctx->nccl_lib, "ncclAllReduce");
    if (!f) return fail(ctx, MR_ERR_COMM, "ncclAllReduce missing");
    CUDA_TRY(ctx, cudaSetDevice(ctx->dev[0].id));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->dev[0].stream;
    const int nccl_int32 = 2, nccl_sum = 0;
    int rc = f(d_buf, d_buf, n, nccl_int32, nccl_sum, ctx->nccl_comm, st);
    if (rc != 0) return fail(ctx, MR_ERR_COMM, "ncclAllReduce failed: %d", rc);
    return MR_OK;
}

int mr_measure_int_peak(mr_ctx *ctx, int device_index, int kind, double *ops_per_second, double *sm_clock_mhz) {
    if (!ctx || !ops_per_second) return fail(ctx, MR_ERR_INVALID, "mr_measure_int_peak: bad arguments");
    if (device_index < 0 || device_index >= (int)ctx->dev.size()) return fail(ctx, MR_ERR_INVALID, "device index out of range");
    if (kind < 0 || kind >= mr::INT_PEAK_KINDS) return fail(ctx, MR_ERR_INVALID, "unknown kind %d", kind);
    Device &dv = ctx->dev[(size_t)device_index];
    CUDA_TRY(ctx, cudaSetDevice(dv.id));
    uint32_t *d_sink = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&d_sink, sizeof(uint32_t) * 1024));
    cudaEvent_t e0, e1;
    CUDA_TRY(ctx, cudaEventCreate(&e0));
    CUDA_TRY(ctx, cudaEventCreate(&e1));
    const int blocks = dv.num_sms * 2, threads = 1024;
    const int iters = 4096;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        CUDA_TRY(ctx, cudaEventRecord(e0, dv.stream));
        mr::launch_int_peak(kind, blocks, threads, iters, d_sink, dv.stream);
        CUDA_TRY(ctx, cudaGetLastError());
        CUDA_TRY(ctx, cudaEventRecord(e1, dv.stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(dv.stream));
        float ms = 0.f;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, e0, e1));
        const double ops = (double)blocks * threads * iters * mr::INT_PEAK_OPS_PER_ITER;
        if (rep > 0) best = std::max(best, ops / (ms * 1e-3));
    }
    ctx->launches += 0; // micro-benchmark launches are not rollout launches
    *ops_per_second = best;
    if (sm_clock_mhz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dv.id);
        *sm_clock_mhz = khz / 1000.0;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_sink);
    return MR_OK;
}

} // extern "C"
