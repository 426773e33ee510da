// rollout_kernel.cuh -- the persistent-warp playout kernel (sm_100a).
//
// Work decomposition
//   work      = global playout indices g = leaf * k + playout; a warp grabs contiguous ranges from a global
//               atomic counter with guided self-scheduling and cuts them at leaf boundaries into
//   task      = (leaf, run of consecutive playout ids); a task never straddles leaves
//   lane      = one playout at a time; lanes that finish pull the next playout id of the task at the
//               next WINDOW boundary (a window = one Philox4x32 block: 4 plies, or 2 plies for the
//               epsilon-MAST policy which needs two words per ply), so all live lanes of a warp are always
//               at the same ply phase: the Philox block is computed warp-synchronously and the mover's
//               colour is a warp-uniform, compile-time-alternating quantity.
//   reduction = per-lane counters -> warp REDUX -> one atomic set per task on results[leaf];
//               MAST / AMAF occurrence tables are accumulated in per-warp SHARED-MEMORY tables (packed
//               visits:16 | score:16) and flushed to global memory with one atomic per touched entry.
//
// Draw discipline (shared with the CPU oracle, DESIGN.md "Draw discipline"), Philox4x32-10 keyed by the seed:
//   uniform / decisive : ctr = (leaf_id, playout_id, ply >> 2, 0);  x0 = out[ply & 3]
//   epsilon-MAST / LGR : ctr = (leaf_id, playout_id, ply >> 1, 1);  x0 = out[2*(ply & 1)], x1 = out[2*(ply & 1) + 1]
#pragma once

#include <type_traits>

#include "games_device.cuh"

namespace mr {

__device__ __forceinline__ int end_code_to_outcome(int code) {
    const int e = code & ST_END_MASK;
    return e == ST_WIN_P0 ? MR_WINNER_P0 : (e == ST_WIN_P1 ? MR_WINNER_P1 : MR_DRAW);
}

constexpr uint32_t STAT_FLUSH_PLIES = 32000; // packed 16-bit fields: flush before any entry can overflow

// A/B switch (the default is the measured winner)
#ifndef MR_LOG_FLUSH
#define MR_LOG_FLUSH 1 // per-lane playout log drained in bulk, instead of a flattened flush at every window boundary
#endif
constexpr uint32_t LOG_TRAILER = 0x8000u; // log entry: table index (player * NSLOTS + slot) or LOG_TRAILER | (utility of player 0 + 1)

// per-warp packed table -> global (visits, score) table, one atomic per non-zero field
template <class G>
__device__ __forceinline__ void flush_stat_table(uint32_t *tab, mr_action_stat *out, uint32_t lane, uint32_t in_slots) {
    __syncwarp();
    for (uint32_t e = lane; e < 2u * G::NSLOTS; e += 32) {
        const uint32_t v = tab[e];
        if (v) {
            tab[e] = 0;
            const uint32_t pl = e >= (uint32_t)G::NSLOTS ? 1u : 0u;
            const uint32_t slot = e - pl * G::NSLOTS;
            mr_action_stat *o = in_slots ? out + e : out + (size_t)pl * G::NA + G::slot_action_id(pl, slot);
            const uint32_t visits = v & 0xffffu;
            const int score = (int)v >> 16; // v = visits + score * 65536 with visits < 65536: the arithmetic shift floors to score
            atomicAdd(&o->visits, visits);
            if (score) atomicAdd(&o->score, score);
        }
    }
    __syncwarp();
}

template <class G, int POLICY, int FLAGS>
__global__ void __launch_bounds__(THREADS_PER_BLOCK, G::min_blocks(POLICY)) rollout_kernel(const KernelParams p) {
    constexpr bool WANT_AMAF = (FLAGS & MR_WANT_AMAF) != 0;
    constexpr bool WANT_MAST = (FLAGS & MR_WANT_MAST_DELTA) != 0;
    constexpr bool WANT_TRACE = (FLAGS & MR_WANT_TRACE) != 0;
    constexpr bool WANT_FINAL = (FLAGS & MR_WANT_FINAL_STATE) != 0;
    constexpr bool MASK_AMAF = WANT_AMAF && G::MASK_AMAF;
    constexpr bool IS_LGR2 = policy_is_lgr2(POLICY); // LGRF-2 runs the LGR-1 machinery too (its inner strategy, both tables written)
    constexpr bool IS_LGR = policy_is_lgr(POLICY);
    constexpr bool IS_NST = POLICY == MR_EPS_NST;
    constexpr bool TRACK_PREV = IS_LGR || IS_NST; // the lane remembers the slot of the preceding action
    constexpr bool PAIR_DRAWS = policy_uses_mast(POLICY) || IS_LGR; // two Philox words per ply
    constexpr bool NEED_SCRATCH = WANT_MAST || (WANT_AMAF && !G::MASK_AMAF) || IS_LGR || IS_NST; // per-lane slot trace
    constexpr int WINDOW = PAIR_DRAWS ? 2 : 4;
    // MAST delta / AMAF occurrence tables without the reply / bigram side tables: every lane appends its moves (and, when a
    // playout ends, a trailer with the outcome) to its own ring log in the scratch buffer; the logs are drained when a task
    // ends or a lane could run out of room -- each lane walks its own log, which holds many playouts by then, so the lanes'
    // trip counts are balanced and no prefix sum / owner search is needed (the LGR / NST tables need (playout id, ply) per
    // entry and keep the flattened flush at every window boundary)
    constexpr bool LOG_FLUSH = MR_LOG_FLUSH && NEED_SCRATCH && !IS_LGR && !IS_NST;

    __shared__ __align__(16) mr_leaf s_leaf[WARPS_PER_BLOCK];
    __shared__ typename G::Shared s_game[WARPS_PER_BLOCK];
    __shared__ uint32_t s_mast[WARPS_PER_BLOCK][WANT_MAST ? 2 * G::NSLOTS : 1];
    __shared__ uint32_t s_amaf[WARPS_PER_BLOCK][WANT_AMAF ? 2 * G::NSLOTS : 1];
    // LGR: per (player, slot of the preceding action) the largest stamp ((g + 1) << 26 | (4096 + ply) << 13 | reply slot);
    // one table per BLOCK (64-bit entries: per-warp copies would not fit next to the MAST / AMAF tables)
    __shared__ unsigned long long s_lgr[IS_LGR ? 2 * G::NSLOTS : 1];

    const uint32_t lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t gwarp = blockIdx.x * WARPS_PER_BLOCK + warp;
    typename G::Shared &sh = s_game[warp];
    uint16_t *scratch = NEED_SCRATCH ? p.trace_scratch + (size_t)gwarp * p.trace_rows * 32u : nullptr;
    uint32_t *mast_tab = s_mast[warp];
    uint32_t *amaf_tab = s_amaf[warp];
    unsigned long long *lgr_tab = s_lgr;
    uint32_t mast_plies = 0, amaf_plies = 0; // warp-uniform: plies accumulated since the last flush

    if (WANT_MAST)
        for (uint32_t e = lane; e < 2u * G::NSLOTS; e += 32) mast_tab[e] = 0;
    if (WANT_AMAF)
        for (uint32_t e = lane; e < 2u * G::NSLOTS; e += 32) amaf_tab[e] = 0;
    if (IS_LGR) {
        for (uint32_t e = threadIdx.x; e < 2u * G::NSLOTS; e += THREADS_PER_BLOCK) lgr_tab[e] = 0;
        __syncthreads();
    }
    G::init_shared(sh, lane); // per-warp lookup tables of the game (Othello: ray masks)
    __syncwarp();

    unsigned long long grab_lo = 0, grab_hi = 0; // the warp's current grab, relative to g_begin
    const unsigned long long work_total = p.g_end - p.g_begin;
    // per-lane playout log (LOG_FLUSH): a ring of p.trace_rows entries per lane, lane-interleaved in the warp's scratch
    uint32_t log_pos = 0, log_tail = 0; // entries [log_tail, log_pos) are not yet in the shared tables
    const uint32_t log_mask = p.trace_rows - 1u;
    const bool log_on = LOG_FLUSH && p.trace_rows != 0; // (the parity variant is compiled with every output and run with any subset)
    for (;;) {
        // ---- next task: the part of the current grab that lies inside one leaf ----
        if (grab_lo >= grab_hi) {
            unsigned long long lo = 0;
            uint32_t want = 0;
            if (lane == 0) {
                const unsigned long long seen = *reinterpret_cast<volatile unsigned long long *>(p.work_counter);
                const unsigned long long remaining = seen < work_total ? work_total - seen : 0ull;
                unsigned long long w = remaining / ((unsigned long long)p.grab_div * p.total_warps);
                w = w > p.grab_max ? p.grab_max : w;
                w = w < p.grab_min ? p.grab_min : w;
                want = (uint32_t)w & ~31u;
                lo = atomicAdd(p.work_counter, (unsigned long long)want);
            }
            lo = __shfl_sync(FULL_MASK, lo, 0);
            want = __shfl_sync(FULL_MASK, want, 0);
            if (lo >= work_total) break;
            grab_lo = lo;
            grab_hi = lo + want < work_total ? lo + want : work_total;
        }
        const unsigned long long g0 = p.g_begin + grab_lo;
        const uint32_t li = (uint32_t)(g0 / p.k);
        const uint32_t pid_begin = (uint32_t)(g0 - (unsigned long long)li * p.k);
        const unsigned long long room = (unsigned long long)p.k - pid_begin; // playouts left in this leaf
        const unsigned long long take = grab_hi - grab_lo < room ? grab_hi - grab_lo : room;
        const uint32_t pid_end = pid_begin + (uint32_t)take;
        grab_lo += take;
        const uint32_t leaf_id = p.leaf_id_base + li;

        // ---- stage the leaf record in shared memory (one coalesced 40-byte read per warp) ----
        __syncwarp();
        if (lane < sizeof(mr_leaf) / 4)
            reinterpret_cast<uint32_t *>(&s_leaf[warp])[lane] = reinterpret_cast<const uint32_t *>(p.leaves + li)[lane];
        __syncwarp();
        const int leaf_status = G::prepare(sh, s_leaf[warp], lane);
        __syncwarp();
        const uint32_t leaf_turn = s_leaf[warp].turn;
        // LGR: the action that led to the leaf (playout()'s prev_action), as 1 + its slot in the numbering of the
        // player who made it (the opponent of the leaf's mover); 0 = none
        uint32_t leaf_prev_slot1 = 0, leaf_prev2_slot1 = 0;
        if constexpr (TRACK_PREV) {
            const uint32_t pa = s_leaf[warp].prev_action_plus1;
            if (pa) leaf_prev_slot1 = G::code_to_slot(1u - (leaf_turn & 1u), pa - 1u) + 1u;
            if constexpr (IS_LGR2) { // the action before that, made by the leaf's mover (table updates of plies 0 and 1 only)
                const uint32_t pa2 = s_leaf[warp].prev2_action_plus1;
                if (pa && pa2) leaf_prev2_slot1 = G::code_to_slot(leaf_turn & 1u, pa2 - 1u) + 1u;
            }
        }

        uint32_t acc_w0 = 0, acc_w1 = 0, acc_draw = 0, acc_cut = 0, acc_depth = 0;
        uint32_t acc_lw0 = 0, acc_lw1 = 0; // LGR: 1 + last playout id each player did not lose
        uint32_t acc_ll0 = 0, acc_ll1 = 0; // LGRF-2: 1 + last playout id each player lost

        if (leaf_status != MR_NOT_TERMINAL) {
            // a terminal leaf: every playout is the zero-ply trial with the leaf's own status
            // (simulate.rs:97-102); nothing is drawn
            const uint32_t cnt = pid_end - pid_begin;
            if (lane == 0) {
                if (leaf_status == MR_WINNER_P0) acc_w0 = cnt;
                else if (leaf_status == MR_WINNER_P1) acc_w1 = cnt;
                else acc_draw = cnt;
                if (leaf_status != MR_WINNER_P1) acc_lw0 = pid_end;
                else acc_ll0 = pid_end;
                if (leaf_status != MR_WINNER_P0) acc_lw1 = pid_end;
                else acc_ll1 = pid_end;
            }
            if (WANT_TRACE || WANT_FINAL) {
                for (uint32_t pid = pid_begin + lane; pid < pid_end; pid += 32) {
                    const size_t g = (size_t)li * p.k + pid;
                    if (WANT_TRACE && p.rec) {
                        mr_playout_record r;
                        r.depth = 0;
                        r.outcome = (uint8_t)leaf_status;
                        r.end_type = MR_END_TERMINAL;
                        r.pad[0] = r.pad[1] = 0;
                        p.rec[g] = r;
                    }
                    if (WANT_FINAL && p.final_state) {
                        mr_leaf fs = s_leaf[warp];
                        fs.prev_action_plus1 = 0; // call context, not game state
                        fs.prev2_action_plus1 = 0;
                        p.final_state[g] = fs;
                    }
                }
            }
        } else {
            G st;
            uint32_t pid = pid_begin + lane;
            uint32_t next_pid = min(pid_begin + 32u, pid_end);
            bool idle = pid >= pid_end;
            bool alive = !idle;
            uint32_t depth = 0, win = 0;
            int end_code = 0;
            bool cutoff = false;
            uint32_t prev_slot1 = leaf_prev_slot1; // LGR: 1 + slot of the preceding action
            uint32_t own_slot1 = 0;                // LGRF-2: 1 + slot of the mover's own previous action in THIS playout
            st.reset(sh);
            // the same value in every lane, but produced by a warp reduction: the compiler can then keep it (and the
            // mover's colour, MAST table index, rank-bit loop and plane addresses derived from it) in uniform registers
            st.turn_u = __reduce_or_sync(FULL_MASK, leaf_turn & 1u);

            // Drain the lanes' logs into the per-warp shared tables: every (action, player) occurrence adds visits += 1,
            // score += utilities[player] (backprop.rs:619-640 GRAVE/AMAF, :857-870 GLOBAL/MAST).  A lane walks its own log
            // BACKWARDS, so the trailer of a playout (its outcome) is met before its moves; the moves of a playout that is
            // still running (the last min(depth, trace_cap) entries of a live lane) stay in the ring.
            auto drain_logs = [&]() {
                const uint32_t keep = alive ? min(depth, p.trace_cap) : 0u;
                const uint32_t n = log_pos - log_tail - keep;
                const uint32_t nmax = __reduce_max_sync(FULL_MASK, n);
                const uint32_t top = log_pos - keep; // one past the newest entry to drain
                const uint32_t off_mask = p.trace_rows * 32u - 1u;
                uint32_t off = (((top - 1u) & log_mask) << 5) | lane; // element offset of the entry under the cursor
                uint32_t inc0 = 1u, inc1 = 1u; // the packed (visits +1, score +-1) increments of the playout under the cursor
                for (uint32_t base = 0; base < nmax; base += 256u) {
                    const uint32_t lim = min(nmax, base + 256u);
                    // (the packed 16-bit fields must not overflow: an entry receives at most 32 * 256 additions per chunk)
                    if (WANT_MAST && p.mast_delta) {
                        if (mast_plies + 32u * 256u > STAT_FLUSH_PLIES) {
                            flush_stat_table<G>(mast_tab, p.mast_delta, lane, p.tables_in_slots);
                            mast_plies = 0;
                        }
                        mast_plies += 32u * (lim - base);
                    }
                    if (WANT_AMAF && !G::MASK_AMAF && p.amaf) {
                        if (amaf_plies + 32u * 256u > STAT_FLUSH_PLIES) {
                            flush_stat_table<G>(amaf_tab, p.amaf + (size_t)li * p.amaf_stride, lane, p.tables_in_slots);
                            amaf_plies = 0;
                        }
                        amaf_plies += 32u * (lim - base);
                    }
#pragma unroll 4
                    for (uint32_t i = base; i < lim; ++i) {
                        if (i < n) {
                            const uint32_t e = scratch[off];
                            off = (off - 32u) & off_mask;
                            if (e & LOG_TRAILER) {
                                const uint32_t u = (e & 3u) - 1u; // utility of player 0 (two's complement)
                                inc0 = 1u + (u << 16);
                                inc1 = 1u - (u << 16);
                            } else {
                                const uint32_t inc = e >= (uint32_t)G::NSLOTS ? inc1 : inc0;
                                if (WANT_MAST && p.mast_delta) atomicAdd(&mast_tab[e], inc);
                                if (WANT_AMAF && !G::MASK_AMAF && p.amaf) atomicAdd(&amaf_tab[e], inc);
                            }
                        }
                    }
                }
                if (alive) {
                    log_tail = top;
                } else { // nothing of this lane is left in the ring: start over at row 0 (small, L2-resident footprint)
                    log_pos = 0;
                    log_tail = 0;
                }
            };

            for (;;) {
                uint32_t r[4];
                philox4x32_10(leaf_id, pid, win, PAIR_DRAWS ? 1u : 0u, p.philox_keys, r);
                // the depth cutoff can only fire in a window that reaches max_depth: one test per window, not per ply
                const bool near_cutoff = depth + (uint32_t)WINDOW > p.max_depth || p.max_depth < (uint32_t)WINDOW;
                // one ply.  PAR is the ply's parity inside the window: a compile-time 0/1 when the window is
                // unrolled (the mover's colour = leaf turn ^ PAR is then fixed per call site), or 2 = "run-time,
                // read st.rt_parity" for games whose window runs as a loop
                auto ply = [&](auto PAR, const uint32_t x0, const uint32_t x1) {
                    constexpr int par = decltype(PAR)::value;
                    if (alive) {
                        if (near_cutoff && depth >= p.max_depth) { // simulate.rs:104-108 TurnLimit
                            alive = false;
                            cutoff = true;
                            end_code = ST_DRAW;
                        } else {
                            uint32_t action = 0, slot = 0;
                            if constexpr (IS_LGR) {
                                // Lgr::select_move (simulate.rs:1013-1032) behind EpsilonGreedy (:538-570): unless this
                                // ply explores, look up the mover's recorded reply to the preceding action
                                uint32_t reply = 0;
                                if (prev_slot1 && p.lgr_in && !(p.eps_enable && x1 <= p.eps_thr_m1)) {
                                    const uint32_t pl = (leaf_turn ^ (par == 2 ? st.rt_parity : (uint32_t)par)) & 1u;
                                    reply = p.lgr_in[pl * (uint32_t)G::NA + G::slot_action_id(1u - pl, prev_slot1 - 1u)];
                                }
                                st.lgr_reply = reply;
                                if constexpr (IS_LGR2) {
                                    // Lgr2::select_move (simulate.rs:1113-1141): the context exists from the playout's
                                    // third ply on (own_prev_action starts empty, simulate.rs:98)
                                    uint32_t reply2 = 0;
                                    if (depth >= 2u && p.lgr2_in && !(p.eps_enable && x1 <= p.eps_thr_m1)) {
                                        const uint32_t pl = (leaf_turn ^ (par == 2 ? st.rt_parity : (uint32_t)par)) & 1u;
                                        reply2 = p.lgr2_in[((size_t)pl * G::NSLOTS + own_slot1 - 1u) * G::NSLOTS + prev_slot1 - 1u];
                                    }
                                    st.lgr_reply2 = reply2;
                                }
                            }
                            if constexpr (IS_NST) {
                                // Nst::select_move (simulate.rs:898-930): the ranks of this ply's context
                                const uint32_t pl = (leaf_turn ^ (par == 2 ? st.rt_parity : (uint32_t)par)) & 1u;
                                const uint32_t ctx = pl * ((uint32_t)G::NSLOTS + 1u) + prev_slot1;
                                st.nst_planes = reinterpret_cast<const uint64_t(*)[MAST_BITS]>(p.nst_planes + (size_t)ctx * G::KINDS * MAST_BITS);
                                st.nst_nbits = p.nst_nbits[ctx];
                                st.nst_ranks = p.nst_rank_small + (size_t)ctx * 8u;
                            }
                            const int code = st.template step<POLICY, par>(sh, x0, x1, p, action, slot);
                            if (code & ST_MOVED) {
                                if (IS_LGR2) own_slot1 = prev_slot1; // the next mover's own previous action is the one before this
                                if (TRACK_PREV) prev_slot1 = slot + 1u;
                                if (WANT_TRACE && p.trace) {
                                    if (depth < p.max_trace)
                                        p.trace[((size_t)li * p.k + pid) * p.max_trace + depth] = (uint16_t)action;
                                }
                                if constexpr (LOG_FLUSH) {
                                    if (depth < p.trace_cap) {
                                        const uint32_t pl = (st.turn_u ^ (par == 2 ? st.rt_parity : (uint32_t)par)) & 1u;
                                        scratch[(log_pos & log_mask) * 32u + lane] = (uint16_t)(pl * (uint32_t)G::NSLOTS + slot);
                                        ++log_pos;
                                    }
                                } else if (NEED_SCRATCH) {
                                    if (depth < p.trace_cap) scratch[depth * 32u + lane] = (uint16_t)slot;
                                }
                                ++depth;
                            }
                            if (code & ST_END_MASK) {
                                alive = false;
                                end_code = code & ST_END_MASK;
                            }
                        }
                    }
                };
                if constexpr (G::rolled(POLICY)) {
#pragma unroll 1
                    for (uint32_t j = 0; j < (uint32_t)WINDOW; ++j) {
                        st.rt_parity = j & 1u;
                        const uint32_t x0 = r[0], x1 = r[1];
                        if constexpr (PAIR_DRAWS) {
                            r[0] = r[2];
                            r[1] = r[3];
                        } else {
                            r[0] = r[1];
                            r[1] = r[2];
                            r[2] = r[3];
                        }
                        ply(std::integral_constant<int, 2>{}, x0, x1);
                    }
                } else if constexpr (PAIR_DRAWS) {
                    ply(std::integral_constant<int, 0>{}, r[0], r[1]);
                    ply(std::integral_constant<int, 1>{}, r[2], r[3]);
                } else {
                    ply(std::integral_constant<int, 0>{}, r[0], 0u);
                    ply(std::integral_constant<int, 1>{}, r[1], 0u);
                    ply(std::integral_constant<int, 0>{}, r[2], 0u);
                    ply(std::integral_constant<int, 1>{}, r[3], 0u);
                }
                ++win;

                // ---- window boundary: retire finished playouts, hand out new playout ids ----
                const bool fin = !alive && !idle;
                const uint32_t fin_mask = __ballot_sync(FULL_MASK, fin);
                if (fin_mask) {
                    if (fin) {
                        const int e = end_code;
                        acc_w0 += (e == ST_WIN_P0);
                        acc_w1 += (e == ST_WIN_P1);
                        acc_draw += (e == ST_DRAW || e == ST_NO_ACTIONS);
                        acc_cut += cutoff;
                        acc_depth += depth;
                        if (IS_LGR) {
                            if (e != ST_WIN_P1) acc_lw0 = max(acc_lw0, pid + 1u);
                            else acc_ll0 = max(acc_ll0, pid + 1u);
                            if (e != ST_WIN_P0) acc_lw1 = max(acc_lw1, pid + 1u);
                            else acc_ll1 = max(acc_ll1, pid + 1u);
                        }
                        const size_t g = (size_t)li * p.k + pid;
                        if (WANT_TRACE && p.rec) {
                            mr_playout_record rr;
                            rr.depth = depth;
                            rr.outcome = (uint8_t)end_code_to_outcome(e);
                            rr.end_type = (uint8_t)(cutoff ? MR_END_TURN_LIMIT
                                                           : (e == ST_NO_ACTIONS ? MR_END_NO_ACTIONS : MR_END_TERMINAL));
                            rr.pad[0] = rr.pad[1] = 0;
                            p.rec[g] = rr;
                        }
                        if (WANT_FINAL && p.final_state) st.store_final(sh, &p.final_state[g], depth, e);
                        if (log_on) { // the playout's trailer: utility of player 0, biased by one
                            const uint32_t code = e == ST_WIN_P0 ? 2u : (e == ST_WIN_P1 ? 0u : 1u);
                            scratch[(log_pos & log_mask) * 32u + lane] = (uint16_t)(LOG_TRAILER | code);
                            ++log_pos;
                        }
                    }
                    if (NEED_SCRATCH && !LOG_FLUSH) {
                        // Statistics from the finished playouts' traces, FLATTENED over the warp: entry q of the
                        // concatenated traces belongs to the finished lane f with excl[f] <= q < incl[f].
                        // Every (action, player) occurrence: visits += 1, score += utilities[player]
                        // (backprop.rs:619-640 GRAVE/AMAF, :857-870 GLOBAL/MAST).
                        __syncwarp();
                        const uint32_t my_d = fin ? min(depth, p.trace_cap) : 0u;
                        uint32_t incl = my_d;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const uint32_t v = __shfl_up_sync(FULL_MASK, incl, o);
                            if (lane >= (uint32_t)o) incl += v;
                        }
                        const uint32_t excl = incl - my_d;
                        const uint32_t total = __shfl_sync(FULL_MASK, incl, 31);
                        for (uint32_t base = 0; base < total; base += 32) {
                            // (the packed 16-bit fields must not overflow: flush before the running count of additions
                            // since the last flush could pass the limit -- checked per 32 entries, so a single retire
                            // round of up to 32 * trace_cap entries is safe for any trace_cap)
                            if (WANT_MAST && p.mast_delta) {
                                if (mast_plies + 32u > STAT_FLUSH_PLIES) {
                                    flush_stat_table<G>(mast_tab, p.mast_delta, lane, p.tables_in_slots);
                                    mast_plies = 0;
                                }
                                mast_plies += 32u;
                            }
                            if (WANT_AMAF && !G::MASK_AMAF && p.amaf) {
                                if (amaf_plies + 32u > STAT_FLUSH_PLIES) {
                                    flush_stat_table<G>(amaf_tab, p.amaf + (size_t)li * p.amaf_stride, lane, p.tables_in_slots);
                                    amaf_plies = 0;
                                }
                                amaf_plies += 32u;
                            }
                            const uint32_t q = base + lane;
                            uint32_t f = 0; // number of lanes whose inclusive sum is <= q == owner lane of entry q
#pragma unroll
                            for (int step = 16; step >= 1; step >>= 1) {
                                const uint32_t v = __shfl_sync(FULL_MASK, incl, (f + step - 1) & 31u);
                                if (v <= q) f += step;
                            }
                            f &= 31u;
                            const uint32_t ex_f = __shfl_sync(FULL_MASK, excl, f);
                            const int e_f = __shfl_sync(FULL_MASK, end_code, f);
                            const uint32_t pid_f = IS_LGR ? __shfl_sync(FULL_MASK, pid, f) : 0u;
                            if (q < total) {
                                const uint32_t t = q - ex_f;
                                const uint32_t slot = scratch[t * 32u + f];
                                const uint32_t pl = (leaf_turn + t) & 1u;
                                const int u_p0 = e_f == ST_WIN_P0 ? 1 : (e_f == ST_WIN_P1 ? -1 : 0);
                                const int u = pl == 0 ? u_p0 : -u_p0;
                                const uint32_t inc = 1u + ((uint32_t)u << 16);
                                if (WANT_MAST && p.mast_delta) atomicAdd(&mast_tab[pl * G::NSLOTS + slot], inc);
                                if (WANT_AMAF && !G::MASK_AMAF && p.amaf) atomicAdd(&amaf_tab[pl * G::NSLOTS + slot], inc);
                                if constexpr (IS_NST) {
                                    // backprop.rs:885-899: every consecutive (preceding action, action) pair of the playout
                                    const uint32_t prev1 = t > 0 ? (uint32_t)scratch[(t - 1u) * 32u + f] + 1u : leaf_prev_slot1;
                                    if (p.nst_delta && prev1)
                                        atomicAdd(&p.nst_delta[((size_t)pl * G::NSLOTS + prev1 - 1u) * G::NSLOTS + slot],
                                                  1ull + ((unsigned long long)(uint32_t)u << 32));
                                }
                                if constexpr (IS_LGR) {
                                    // backprop.rs:908-923: a reply is recorded when its player's utility is the trial's
                                    // maximum (u >= 0 here); last write wins == the largest (playout index, ply) stamp
                                    const uint32_t prev1 = t > 0 ? (uint32_t)scratch[(t - 1u) * 32u + f] + 1u : leaf_prev_slot1;
                                    if (p.lgr_stamps && u >= 0 && prev1) {
                                        const unsigned long long g = (unsigned long long)li * p.k + pid_f;
                                        const unsigned long long stamp = ((g + 1ull) << 26) | ((unsigned long long)(4096u + t) << 13) | slot;
                                        atomicMax(&lgr_tab[pl * G::NSLOTS + prev1 - 1u], stamp);
                                    }
                                    if constexpr (IS_LGR2) {
                                        // backprop.rs:926-966 over the windows (own previous, opponent's answer, this
                                        // action) that end in a playout action
                                        const uint32_t own1 = t > 1 ? (uint32_t)scratch[(t - 2u) * 32u + f] + 1u
                                                                    : (t == 1 ? leaf_prev_slot1 : leaf_prev2_slot1);
                                        if (own1 && prev1) {
                                            const size_t key = ((size_t)pl * G::NSLOTS + own1 - 1u) * G::NSLOTS + prev1 - 1u;
                                            const unsigned long long g = (unsigned long long)li * p.k + pid_f;
                                            if (p.lgr2_target) { // pass 2: does a LATER losing playout remove the candidate?
                                                const unsigned long long tgt = p.lgr2_target[key];
                                                if (u < 0 && tgt && (uint32_t)(tgt & 0xffffull) == slot + 1u && g + 1ull > (tgt >> 16))
                                                    p.lgr2_removed[key] = 1;
                                            } else if (p.lgr2_stamps && u >= 0) { // pass 1: the last insert per context
                                                atomicMax(&p.lgr2_stamps[key], ((g + 1ull) << 26) | ((unsigned long long)(4096u + t) << 13) | slot);
                                            }
                                        }
                                    }
                                }
                            }
                        }
                        __syncwarp();
                    }
                    if constexpr (MASK_AMAF) {
                        // Hex: occurrences = final stones minus leaf stones, one finished lane at a time, no atomics
                        if (p.amaf) {
                            __syncwarp();
                            for (uint32_t m = fin_mask; m; m &= m - 1) {
                                const int f = __ffs(m) - 1;
                                const uint32_t d_f = __shfl_sync(FULL_MASK, depth, f);
                                const int e_f = __shfl_sync(FULL_MASK, end_code, f);
                                const int u_p0 = e_f == ST_WIN_P0 ? 1 : (e_f == ST_WIN_P1 ? -1 : 0);
                                if (amaf_plies + d_f > STAT_FLUSH_PLIES) {
                                    flush_stat_table<G>(amaf_tab, p.amaf + (size_t)li * p.amaf_stride, lane, p.tables_in_slots);
                                    amaf_plies = 0;
                                }
                                amaf_plies += d_f;
                                st.amaf_from_masks(sh, amaf_tab, f, lane, u_p0);
                            }
                            __syncwarp();
                        }
                    }
                    const uint32_t rank = __popc(fin_mask & lanemask_lt());
                    const uint32_t nfin = __popc(fin_mask);
                    const uint32_t avail = pid_end - next_pid; // ids left in the task (next_pid <= pid_end)
                    const bool got = fin && rank < avail;
                    if (log_on) {
                        // a lane starts a playout only with room for all of it (trace_cap moves and the trailer)
                        if (__any_sync(FULL_MASK, got && log_pos - log_tail + p.trace_cap + 1u > p.trace_rows)) drain_logs();
                    }
                    if (fin) {
                        if (got) {
                            pid = next_pid + rank;
                            st.reset(sh);
                            alive = true;
                            depth = 0;
                            win = 0;
                            end_code = 0;
                            cutoff = false;
                            prev_slot1 = leaf_prev_slot1; // the new playout starts from the leaf's context again
                            own_slot1 = 0;
                        } else {
                            idle = true;
                        }
                    }
                    next_pid += min(nfin, avail);
                    if (__all_sync(FULL_MASK, idle)) break;
                }
            }
            if (log_on) drain_logs(); // every lane is idle: the logs are emptied and restart at row 0
            // the AMAF table is per leaf: it leaves shared memory with the task
            if (WANT_AMAF && p.amaf) {
                flush_stat_table<G>(amaf_tab, p.amaf + (size_t)li * p.amaf_stride, lane, p.tables_in_slots);
                amaf_plies = 0;
            }
        }

        // ---- one atomic set per task ----
        acc_w0 = __reduce_add_sync(FULL_MASK, acc_w0);
        acc_w1 = __reduce_add_sync(FULL_MASK, acc_w1);
        acc_draw = __reduce_add_sync(FULL_MASK, acc_draw);
        acc_cut = __reduce_add_sync(FULL_MASK, acc_cut);
        acc_depth = __reduce_add_sync(FULL_MASK, acc_depth);
        if (IS_LGR && p.lgr_last) {
            acc_lw0 = __reduce_max_sync(FULL_MASK, acc_lw0);
            acc_lw1 = __reduce_max_sync(FULL_MASK, acc_lw1);
            if (lane == 0) {
                if (acc_lw0) atomicMax(&p.lgr_last[2u * li], acc_lw0);
                if (acc_lw1) atomicMax(&p.lgr_last[2u * li + 1u], acc_lw1);
            }
        }
        if (IS_LGR2 && p.lgr2_last_lose) {
            acc_ll0 = __reduce_max_sync(FULL_MASK, acc_ll0);
            acc_ll1 = __reduce_max_sync(FULL_MASK, acc_ll1);
            if (lane == 0) {
                if (acc_ll0) atomicMax(&p.lgr2_last_lose[2u * li], acc_ll0);
                if (acc_ll1) atomicMax(&p.lgr2_last_lose[2u * li + 1u], acc_ll1);
            }
        }
        if (lane == 0) {
            mr_leaf_result *res = p.results + li;
            if (acc_w0) atomicAdd(&res->wins[0], acc_w0);
            if (acc_w1) atomicAdd(&res->wins[1], acc_w1);
            if (acc_draw) atomicAdd(&res->draws, acc_draw);
            if (acc_cut) atomicAdd(&res->cutoffs, acc_cut);
            if (acc_depth) atomicAdd(reinterpret_cast<unsigned long long *>(&res->sum_depth), (unsigned long long)acc_depth);
        }
    }
    // the MAST delta is batch-wide: it leaves shared memory when the warp runs out of tasks
    if (WANT_MAST && p.mast_delta) flush_stat_table<G>(mast_tab, p.mast_delta, lane, p.tables_in_slots);
    if (IS_LGR && p.lgr_stamps) {
        // slot space -> dense action ids: key = (player, id of the preceding action, made by the other player)
        __syncthreads(); // every warp of the block has run out of work
        for (uint32_t e = threadIdx.x; e < 2u * G::NSLOTS; e += THREADS_PER_BLOCK) {
            const unsigned long long v = lgr_tab[e];
            if (v) {
                const uint32_t pl = e >= (uint32_t)G::NSLOTS ? 1u : 0u;
                const uint32_t prev_slot = e - pl * G::NSLOTS;
                const unsigned long long out = (v & ~0x1fffull) | G::slot_action_id(pl, (uint32_t)(v & 0x1fffull));
                atomicMax(&p.lgr_stamps[(size_t)pl * G::NA + G::slot_action_id(1u - pl, prev_slot)], out);
            }
        }
    }
    batch_epilogue(p);
}

} // namespace mr
