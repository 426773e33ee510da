//! mcts-gpu -- the simulation phase of thomasmarsh/mcts on a B200, behind the crate's own plug-in surface.
//!
//! * [`GpuRollout<G>`] implements `SimulateStrategy<G>` (mcts/src/strategies/mcts/simulate.rs:48-155): same `playout`
//!   signature and `Trial` result, so it can stand wherever `Uniform`, `EpsilonGreedy<Mast>` or `DecisiveMove<..>` stand
//!   (`Strategy::Simulate`, config.rs:274-288), and behind `DynSimulate` (mcts-tune/src/config_ir.rs:475-555) through the
//!   blanket `ErasedSimulateStrategy` impl -- which needs nothing beyond `SimulateStrategy + 'static`.
//! * [`BatchedRollout<G>`] + [`run_batched`] are the throughput form: the sibling of `choose_action_tree_parallel`
//!   (search/parallel.rs:118-258) that gathers a batch of leaves under virtual loss, submits it as ONE kernel launch and
//!   applies the aggregated results (node.rs:673-681 summed over the k trials of a leaf).
//!
//! NOT COMPILED in this repository's image (no cargo / rustc there).  `ffi.rs` is generated from the C header and
//! checked by tests/test_rust_shim.py; everything this file calls is exercised through the same C ABI by the Python
//! parity suite.  The C++ driver in csrc/tree_search.cu is the measured stand-in for `run_batched`.
pub mod ffi;

use ffi::*;
use mcts::game::{Game, PlayerIndex, TerminalStatus};
use mcts::strategies::mcts::config::{BackpropFlags, GLOBAL, LGR, LGR2, NST};
use mcts::strategies::mcts::simulate::{EndType, SimulateStrategy, Status, Trial};
use mcts::strategies::mcts::TreeStats;
use rand::rngs::SmallRng;
use rand::Rng;
use std::ffi::CStr;
use std::marker::PhantomData;
use std::ptr;
use std::sync::{Arc, Mutex};

// ---------------------------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------------------------

/// Error of the rollout engine: the `mr_status` code and `mr_last_error()`'s text.  The reference's strategies have no
/// error channel (they panic on impossible states); `GpuRollout::playout` therefore panics with this message, the
/// batched driver returns it.
#[derive(Debug, Clone)]
pub struct GpuError {
    pub code: i32,
    pub message: String,
}

/// One `mr_ctx`: streams, staging buffers and slots for the GPUs it drives.  Calls on one context are not thread-safe
/// (include/mcts_rollout.h "Threading"), hence the mutex: a cloned strategy shares its context, exactly as the
/// reference's workers share `&TreeStats`.
pub struct Ctx {
    raw: *mut MrCtx,
}
unsafe impl Send for Ctx {}

impl Ctx {
    /// `device_ids` empty = device 0.
    pub fn new(device_ids: &[i32]) -> Result<Ctx, GpuError> {
        let mut raw: *mut MrCtx = ptr::null_mut();
        let rc = unsafe {
            if device_ids.is_empty() {
                mr_init(ptr::null(), 1, &mut raw)
            } else {
                mr_init(device_ids.as_ptr(), device_ids.len() as i32, &mut raw)
            }
        };
        if rc != MR_OK as i32 {
            return Err(GpuError { code: rc, message: last_error(ptr::null()) });
        }
        Ok(Ctx { raw })
    }
    fn check(&self, rc: i32) -> Result<(), GpuError> {
        if rc == MR_OK as i32 {
            Ok(())
        } else {
            Err(GpuError { code: rc, message: last_error(self.raw) })
        }
    }
}
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { mr_destroy(self.raw) }
    }
}
fn last_error(ctx: *const MrCtx) -> String {
    unsafe { CStr::from_ptr(mr_last_error(ctx)).to_string_lossy().into_owned() }
}

// ---------------------------------------------------------------------------------------------------------------------
// per-game encoding: what replaces moving `G::S` into `playout`
// ---------------------------------------------------------------------------------------------------------------------

/// A game whose playouts the engine runs.  `encode` keeps exactly the fields a playout reads (include/mcts_rollout.h
/// `mr_leaf`); Zobrist hashes are tree-side state and are rebuilt by `decode`.
pub trait GpuGame: Game
where
    Self::A: Clone,
{
    const GAME_ID: u32;
    fn encode(state: &Self::S) -> MrLeaf;
    /// The state a final `mr_leaf` describes (parity / single-trial path only).
    fn decode(leaf: &MrLeaf) -> Self::S;
    /// Compact action code of traces and reply tables: column / (src << 8) | dst / square (64 = PASS) / cell.
    fn action_code(a: &Self::A) -> u16;
    fn action_from_code(code: u16) -> Self::A;
    fn player_from_index(i: usize) -> Self::P;
}

fn blank_leaf() -> MrLeaf {
    MrLeaf { board: [0; 4], turn: 0, flag: 0, prev_action_plus1: 0, prev2_action_plus1: 0, pad: [0; 2] }
}

mod games {
    use super::*;
    use game_breakthrough as bt;
    use game_connect4 as c4;
    use game_hex_gen as hex;
    use game_knightthrough as kt;
    use game_othello as oth;

    // games/connect4/src/lib.rs:211-221 (State<6,7>: black, white, turn, winner), :81 Move(col)
    impl GpuGame for c4::Standard {
        const GAME_ID: u32 = MR_CONNECT4;
        fn encode(s: &Self::S) -> MrLeaf {
            MrLeaf { board: [s.black().bits(), s.white().bits(), 0, 0], turn: s.turn().to_index() as u8, flag: s.has_winner() as u8, ..blank_leaf() }
        }
        fn decode(l: &MrLeaf) -> Self::S {
            c4::State::from_parts(c4::BitBoard::from_bits(l.board[0]), c4::BitBoard::from_bits(l.board[1]), Self::player_from_index(l.turn as usize), l.flag != 0)
        }
        fn action_code(a: &Self::A) -> u16 {
            a.col() as u16
        }
        fn action_from_code(c: u16) -> Self::A {
            c4::Move(c as u8)
        }
        fn player_from_index(i: usize) -> Self::P {
            if i == 0 { c4::Player::Black } else { c4::Player::White }
        }
    }

    // games/breakthrough/src/lib.rs:56-62 (State<8,8>), :42 Move(src, dst)
    impl GpuGame for bt::Breakthrough<8, 8> {
        const GAME_ID: u32 = MR_BREAKTHROUGH;
        fn encode(s: &Self::S) -> MrLeaf {
            MrLeaf { board: [s.black().bits(), s.white().bits(), 0, 0], turn: s.turn().to_index() as u8, flag: s.has_winner() as u8, ..blank_leaf() }
        }
        fn decode(l: &MrLeaf) -> Self::S {
            bt::State::new(bt::BitBoard::from_bits(l.board[0]), bt::BitBoard::from_bits(l.board[1]), Self::player_from_index(l.turn as usize), l.flag != 0)
        }
        fn action_code(a: &Self::A) -> u16 {
            ((a.src() as u16) << 8) | a.dst() as u16
        }
        fn action_from_code(c: u16) -> Self::A {
            bt::Move((c >> 8) as u8, (c & 0xff) as u8)
        }
        fn player_from_index(i: usize) -> Self::P {
            if i == 0 { bt::Player::Black } else { bt::Player::White }
        }
    }

    // games/knightthrough/src/lib.rs:51-57: the same state shape and move type
    impl GpuGame for kt::Knightthrough<8, 8> {
        const GAME_ID: u32 = MR_KNIGHTTHROUGH;
        fn encode(s: &Self::S) -> MrLeaf {
            MrLeaf { board: [s.black().bits(), s.white().bits(), 0, 0], turn: s.turn().to_index() as u8, flag: s.has_winner() as u8, ..blank_leaf() }
        }
        fn decode(l: &MrLeaf) -> Self::S {
            kt::State::new(kt::BitBoard::from_bits(l.board[0]), kt::BitBoard::from_bits(l.board[1]), Self::player_from_index(l.turn as usize), l.flag != 0)
        }
        fn action_code(a: &Self::A) -> u16 {
            ((a.src() as u16) << 8) | a.dst() as u16
        }
        fn action_from_code(c: u16) -> Self::A {
            kt::Move((c >> 8) as u8, (c & 0xff) as u8)
        }
        fn player_from_index(i: usize) -> Self::P {
            if i == 0 { kt::Player::Black } else { kt::Player::White }
        }
    }

    // games/othello/src/lib.rs:358-370 (pub black, white, turn, last_pass, hashes), :81 Move(square), PASS = Move(64)
    impl GpuGame for oth::Othello {
        const GAME_ID: u32 = MR_OTHELLO;
        fn encode(s: &Self::S) -> MrLeaf {
            MrLeaf { board: [s.black.bits(), s.white.bits(), 0, 0], turn: s.turn.to_index() as u8, flag: s.last_pass as u8, ..blank_leaf() }
        }
        fn decode(l: &MrLeaf) -> Self::S {
            // the eight incremental hashes are a function of the boards, the turn and last_pass; rebuilt as State::default()
            // builds them (lib.rs:372-388), plus the turn / pass constants apply() toggles (:530-540)
            let mut hashes = [0u64; 8];
            oth::xor_piece_range(&mut hashes, l.board[0], oth::Player::Black);
            oth::xor_piece_range(&mut hashes, l.board[1], oth::Player::White);
            if l.turn != 0 {
                oth::xor_const(&mut hashes, oth::ZOBRIST_TURN); // the index apply() uses for "White to move"
            }
            if l.flag != 0 {
                oth::xor_const(&mut hashes, oth::ZOBRIST_LAST_PASS);
            }
            oth::State { black: oth::BB::from_bits(l.board[0]), white: oth::BB::from_bits(l.board[1]), turn: Self::player_from_index(l.turn as usize), last_pass: l.flag != 0, hashes }
        }
        fn action_code(a: &Self::A) -> u16 {
            a.0 as u16
        }
        fn action_from_code(c: u16) -> Self::A {
            oth::Move(c as u8)
        }
        fn player_from_index(i: usize) -> Self::P {
            if i == 0 { oth::Player::Black } else { oth::Player::White }
        }
    }

    // games/hex-gen/src/lib.rs:95-99 (Position<11,2>: pub turn, pub occupied[2] of Board<[u64;2]>), S = HashedPosition<11,2>
    // (:153-156, pub position / hash)
    impl GpuGame for hex::Hex<11, 2> {
        const GAME_ID: u32 = MR_HEX11;
        fn encode(s: &Self::S) -> MrLeaf {
            let mut board = [0u64; 4];
            for p in 0..2 {
                for (w, word) in s.position.occupied[p].words().enumerate() {
                    board[2 * p + w] = word;
                }
            }
            MrLeaf { board, turn: s.position.turn.to_index() as u8, ..blank_leaf() }
        }
        fn decode(l: &MrLeaf) -> Self::S {
            let mut position = hex::Position::<11, 2>::new();
            position.turn = Self::player_from_index(l.turn as usize);
            for p in 0..2 {
                for cell in 0..121usize {
                    if (l.board[2 * p + cell / 64] >> (cell % 64)) & 1 == 1 {
                        position.occupied[p].set_index(cell);
                    }
                }
            }
            // HashedPosition::compute_hash is private; the same FNV-1a over the occupied words and the turn (:171-183)
            let mut hash: u64 = 0x49534419cd9c812a;
            for b in position.occupied {
                for w in b.words() {
                    hash ^= w;
                    hash = hash.wrapping_mul(0x100000001b3);
                }
            }
            hash ^= position.turn.to_index() as u64;
            hex::HashedPosition { position, hash }
        }
        fn action_code(a: &Self::A) -> u16 {
            a.0 as u16
        }
        fn action_from_code(c: u16) -> Self::A {
            hex::Move(c as u8)
        }
        fn player_from_index(i: usize) -> Self::P {
            if i == 0 { hex::Player::P0 } else { hex::Player::P1 }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// policies
// ---------------------------------------------------------------------------------------------------------------------

/// The reference's `Simulate` compositions (strategy.rs) as engine policies.
#[derive(Clone, Copy, Debug, PartialEq)]
pub enum Policy {
    /// `Uniform` (simulate.rs:157-216)
    Uniform,
    /// `EpsilonGreedy<G, Mast>` (simulate.rs:485-570, 819-848); bare `Mast` is epsilon = 0
    EpsilonGreedyMast { epsilon: f64 },
    /// `DecisiveMove<G, Uniform>` (simulate.rs:573-760); `mode` = MR_DECISIVE_WIN / WINLOSS / WINLOSSDRAW / ANTI
    DecisiveMove { mode: u32 },
    /// `DecisiveMove<G, EpsilonGreedy<G, Mast>>` (the `ucb1_tuned_dm_mast` and `rave` families); win_loss = mode WinLoss
    DecisiveMoveMast { epsilon: f64, win_loss: bool },
    /// `EpsilonGreedy<G, Lgr<G, Uniform>>` (Ucb1Lgr)
    EpsilonGreedyLgr { epsilon: f64 },
    /// `EpsilonGreedy<G, Nst>` (Ucb1Nst)
    EpsilonGreedyNst { epsilon: f64, backoff_threshold: u32 },
    /// `EpsilonGreedy<G, Lgr2<G, Lgr<G, Uniform | Mast>>>` (Ucb1Lgr2, Ucb1Lgr2Mast)
    EpsilonGreedyLgr2 { epsilon: f64, mast: bool },
}

impl Policy {
    pub fn id(&self) -> u32 {
        match *self {
            Policy::Uniform => MR_UNIFORM,
            Policy::EpsilonGreedyMast { .. } => MR_EPS_MAST,
            Policy::DecisiveMove { mode } => mode,
            Policy::DecisiveMoveMast { win_loss, .. } => if win_loss { MR_DECISIVE_MAST_WINLOSS } else { MR_DECISIVE_MAST_WIN },
            Policy::EpsilonGreedyLgr { .. } => MR_EPS_LGR,
            Policy::EpsilonGreedyNst { .. } => MR_EPS_NST,
            Policy::EpsilonGreedyLgr2 { mast, .. } => if mast { MR_EPS_LGR2_MAST } else { MR_EPS_LGR2 },
        }
    }
    pub fn epsilon(&self) -> f64 {
        match *self {
            Policy::EpsilonGreedyMast { epsilon }
            | Policy::DecisiveMoveMast { epsilon, .. }
            | Policy::EpsilonGreedyLgr { epsilon }
            | Policy::EpsilonGreedyNst { epsilon, .. }
            | Policy::EpsilonGreedyLgr2 { epsilon, .. } => epsilon,
            _ => 0.0,
        }
    }
    pub fn uses_mast(&self) -> bool {
        matches!(self, Policy::EpsilonGreedyMast { .. } | Policy::DecisiveMoveMast { .. } | Policy::EpsilonGreedyNst { .. } | Policy::EpsilonGreedyLgr2 { mast: true, .. })
    }
    /// The flags of the strategy this policy replaces (Mast :826-828, Nst :894-896, Lgr :984-986, Lgr2 :1101-1103).
    pub fn backprop_flags(&self) -> BackpropFlags {
        match *self {
            Policy::Uniform | Policy::DecisiveMove { .. } => BackpropFlags(0),
            Policy::EpsilonGreedyMast { .. } | Policy::DecisiveMoveMast { .. } => BackpropFlags(GLOBAL),
            Policy::EpsilonGreedyLgr { .. } => BackpropFlags(LGR),
            Policy::EpsilonGreedyNst { .. } => BackpropFlags(GLOBAL | NST),
            Policy::EpsilonGreedyLgr2 { mast, .. } => BackpropFlags(LGR2 | LGR | if mast { GLOBAL } else { 0 }),
        }
    }
}

/// `TreeStats.player_actions` (search/shared.rs:88) as the dense `[2][mr_num_actions]` table `mr_submit` takes.  The
/// reference's scores are f64 sums of +1 / 0 / -1 utilities, i.e. integers.
pub fn mast_snapshot<G: GpuGame>(stats: &TreeStats<G>) -> Vec<MrActionStat>
where
    G::A: Clone,
{
    let na = unsafe { mr_num_actions(G::GAME_ID as i32) } as usize;
    let through = G::GAME_ID == MR_BREAKTHROUGH || G::GAME_ID == MR_KNIGHTTHROUGH;
    let mut table = vec![MrActionStat { visits: 0, score: 0 }; 2 * na];
    for (p, lock) in stats.player_actions.iter().enumerate().take(2) {
        for (action, st) in lock.read().unwrap().iter() {
            let code = G::action_code(action) as usize;
            let id = if through { (code >> 8) * 64 + (code & 0xff) } else { code };
            table[p * na + id] = MrActionStat { visits: st.num_visits, score: st.score as i32 };
        }
    }
    table
}

// ---------------------------------------------------------------------------------------------------------------------
// GpuRollout<G>: the SimulateStrategy drop-in
// ---------------------------------------------------------------------------------------------------------------------

pub struct GpuRollout<G: GpuGame>
where
    G::A: Clone,
{
    ctx: Arc<Mutex<Ctx>>,
    pub policy: Policy,
    _g: PhantomData<G>,
}

impl<G: GpuGame> Clone for GpuRollout<G>
where
    G::A: Clone,
{
    fn clone(&self) -> Self {
        GpuRollout { ctx: self.ctx.clone(), policy: self.policy, _g: PhantomData }
    }
}

/// `SimulateStrategy: Default` (the trait bound at simulate.rs:48): device 0, uniform rollouts.  Panics without a
/// usable sm_100 device -- there is no CPU fallback to fall back to.
impl<G: GpuGame> Default for GpuRollout<G>
where
    G::A: Clone,
{
    fn default() -> Self {
        GpuRollout::new(&[], Policy::Uniform).unwrap_or_else(|e| panic!("mcts-gpu: {} ({})", e.message, e.code))
    }
}

impl<G: GpuGame> GpuRollout<G>
where
    G::A: Clone,
{
    pub fn new(device_ids: &[i32], policy: Policy) -> Result<Self, GpuError> {
        Ok(GpuRollout { ctx: Arc::new(Mutex::new(Ctx::new(device_ids)?)), policy, _g: PhantomData })
    }
}

unsafe impl<G: GpuGame> Sync for GpuRollout<G> where G::A: Clone {}
unsafe impl<G: GpuGame> Send for GpuRollout<G> where G::A: Clone {}

impl<G: GpuGame> SimulateStrategy<G> for GpuRollout<G>
where
    G::A: Clone,
{
    /// One trial with the reference's argument meaning (simulate.rs:84-143).  `rng` supplies the key of the
    /// counter-based stream (its next u64 = the Philox seed; DESIGN.md "Draw discipline"), `prev_action` travels in the
    /// leaf record (read by the LGR / LGRF-2 / NST policies only, as in the reference).  One kernel launch for one
    /// playout is latency-bound by construction: searches that want throughput use [`run_batched`].
    fn playout(&mut self, state: G::S, max_playout_depth: usize, stats: &TreeStats<G>, prev_action: Option<G::A>, rng: &mut SmallRng) -> Trial<G> {
        const MAX_TRACE: usize = 512; // longest playout of the five games is Knightthrough's, bounded by max_playout_depth
        let mut leaf = G::encode(&state);
        if let Some(a) = prev_action.as_ref() {
            leaf.prev_action_plus1 = G::action_code(a) + 1;
        }
        let desc = MrBatchDesc {
            game: G::GAME_ID,
            policy: self.policy.id(),
            n_leaves: 1,
            playouts_per_leaf: 1,
            // usize::MAX is the reference's "no limit" (config.rs:440); the engine's is 0
            max_playout_depth: if max_playout_depth >= u32::MAX as usize { 0 } else { max_playout_depth as u32 },
            flags: MR_WANT_TRACE | MR_WANT_FINAL_STATE,
            seed: rng.gen::<u64>(),
            leaf_id_base: 0,
            max_trace: MAX_TRACE as u32,
            epsilon: self.policy.epsilon(),
        };
        let mast = if self.policy.uses_mast() { Some(mast_snapshot::<G>(stats)) } else { None };
        let mut result = MrLeafResult { wins: [0; 2], draws: 0, cutoffs: 0, sum_depth: 0 };
        let mut trace = vec![0u16; MAX_TRACE];
        let mut rec = MrPlayoutRecord { depth: 0, outcome: 0, end_type: 0, pad: [0; 2] };
        let mut fin = blank_leaf();
        {
            let ctx = self.ctx.lock().unwrap();
            let rc = unsafe {
                mr_playout_batch(
                    ctx.raw, &desc, &leaf, mast.as_ref().map_or(ptr::null(), |m| m.as_ptr()), &mut result, ptr::null_mut(), ptr::null_mut(),
                    trace.as_mut_ptr(), &mut rec, &mut fin,
                )
            };
            if let Err(e) = ctx.check(rc) {
                panic!("mcts-gpu: playout failed: {} ({})", e.message, e.code); // the trait has no error channel
            }
        }
        let depth = rec.depth as usize;
        let first = leaf.turn as usize;
        let actions: Vec<(G::A, usize)> = trace[..depth.min(MAX_TRACE)].iter().enumerate().map(|(t, &c)| (G::action_from_code(c), (first + t) & 1)).collect();
        // EndType / terminal exactly as the loop at simulate.rs:97-116 leaves them
        let (end_type, terminal) = match rec.end_type as u32 {
            MR_END_TURN_LIMIT => (EndType::TurnLimit, TerminalStatus::NotTerminal),
            MR_END_NO_ACTIONS => (EndType::NaturalEnd, TerminalStatus::NotTerminal),
            _ => (
                EndType::NaturalEnd,
                match rec.outcome as u32 {
                    MR_WINNER_P0 => TerminalStatus::Winner(G::player_from_index(0)),
                    MR_WINNER_P1 => TerminalStatus::Winner(G::player_from_index(1)),
                    _ => TerminalStatus::Draw,
                },
            ),
        };
        Trial { actions, state: G::decode(&fin), status: Status { end_type: Some(end_type) }, depth, terminal, cutoff_utilities: None }
    }

    fn backprop_flags(&self) -> BackpropFlags {
        self.policy.backprop_flags()
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// the batched driver
// ---------------------------------------------------------------------------------------------------------------------

/// What the k trials of one leaf add up to: the arguments of k calls of `ChildArray::update` / `NodeStats::update`
/// (node.rs:295-302, 673-681) and of `backprop_step`'s depth accounting (shared.rs:814-818).
#[derive(Clone, Copy, Debug)]
pub struct LeafAggregate {
    pub trials: u32,
    pub wins: [u32; 2],
    pub draws: u32,
    pub cutoffs: u32,
    pub sum_depth: u64,
}
impl LeafAggregate {
    /// Sum of `utilities[p]` over the trials (winner +1, others -1, draw 0: game.rs:128-141).
    pub fn utility_sum(&self, p: usize) -> f64 {
        let decisive = self.trials - self.draws;
        self.wins[p] as f64 - (decisive - self.wins[p]) as f64
    }
    /// Sum of `utilities[p]^2` over the trials.
    pub fn squared_sum(&self) -> f64 {
        (self.trials - self.draws) as f64
    }
}

/// One batch in flight on one slot.
pub struct PendingBatch {
    slot: i32,
    n_leaves: usize,
    k: u32,
    flags: u32,
}

/// Results of one batch.  `mast_delta` / `amaf` are in SLOT space (`MR_TABLES_IN_SLOTS`): `[2][S]` and `[n_leaves][2][S]`
/// with S = `mr_num_slots`; `mr_slot_action(game, player, slot)` gives the action code of an entry.
pub struct BatchResults {
    pub leaves: Vec<LeafAggregate>,
    pub mast_delta: Vec<MrActionStat>,
    pub amaf: Vec<MrActionStat>,
}

pub struct BatchedRollout<G: GpuGame>
where
    G::A: Clone,
{
    ctx: Ctx,
    pub policy: Policy,
    pub seed: u64,
    next_leaf_id: u32,
    _g: PhantomData<G>,
}

impl<G: GpuGame> BatchedRollout<G>
where
    G::A: Clone,
{
    pub fn new(device_ids: &[i32], policy: Policy, seed: u64) -> Result<Self, GpuError> {
        let ctx = Ctx::new(device_ids)?;
        unsafe { mr_set_kernel_timing(ctx.raw, 0) }; // latency matters here, kernel timing does not
        Ok(BatchedRollout { ctx, policy, seed, next_leaf_id: 0, _g: PhantomData })
    }

    /// Replaces one `simulate_many` fan-out (search/core.rs:218-258) per leaf: `leaves[i]` = (state, the last edge of the
    /// descent = playout()'s prev_action, the edge before it) gets `k` playouts.  Returns at once; up to MR_NUM_SLOTS
    /// batches may be in flight, each on its own slot.
    pub fn submit(&mut self, slot: usize, leaves: &[(G::S, Option<G::A>, Option<G::A>)], k: u32, max_playout_depth: usize, stats: &TreeStats<G>, want_amaf: bool) -> Result<PendingBatch, GpuError> {
        assert!(slot < MR_NUM_SLOTS && !leaves.is_empty());
        let recs: Vec<MrLeaf> = leaves
            .iter()
            .map(|(s, prev, prev2)| {
                let mut l = G::encode(s);
                l.prev_action_plus1 = prev.as_ref().map_or(0, |a| G::action_code(a) + 1);
                l.prev2_action_plus1 = prev2.as_ref().map_or(0, |a| G::action_code(a) + 1);
                l
            })
            .collect();
        let mast = if self.policy.uses_mast() { Some(mast_snapshot::<G>(stats)) } else { None };
        let flags = MR_TABLES_IN_SLOTS | if mast.is_some() { MR_WANT_MAST_DELTA } else { 0 } | if want_amaf { MR_WANT_AMAF } else { 0 };
        let desc = MrBatchDesc {
            game: G::GAME_ID,
            policy: self.policy.id(),
            n_leaves: recs.len() as u32,
            playouts_per_leaf: k,
            max_playout_depth: if max_playout_depth >= u32::MAX as usize { 0 } else { max_playout_depth as u32 },
            flags,
            seed: self.seed,
            leaf_id_base: self.next_leaf_id, // distinct Philox streams for every leaf of the search
            max_trace: 0,
            epsilon: self.policy.epsilon(),
        };
        self.next_leaf_id = self.next_leaf_id.wrapping_add(recs.len() as u32);
        let rc = unsafe { mr_submit(self.ctx.raw, slot as i32, &desc, recs.as_ptr(), mast.as_ref().map_or(ptr::null(), |m| m.as_ptr())) };
        self.ctx.check(rc)?; // mr_submit has copied leaves and table into its staging buffers
        Ok(PendingBatch { slot: slot as i32, n_leaves: recs.len(), k, flags })
    }

    /// Blocks until THAT slot's batch is done (later slots keep running).
    pub fn wait(&mut self, batch: PendingBatch) -> Result<BatchResults, GpuError> {
        let s = unsafe { mr_num_slots(G::GAME_ID as i32) } as usize;
        let mut raw = vec![MrLeafResult { wins: [0; 2], draws: 0, cutoffs: 0, sum_depth: 0 }; batch.n_leaves];
        let mut delta = vec![MrActionStat { visits: 0, score: 0 }; if batch.flags & MR_WANT_MAST_DELTA != 0 { 2 * s } else { 0 }];
        let mut amaf = vec![MrActionStat { visits: 0, score: 0 }; if batch.flags & MR_WANT_AMAF != 0 { batch.n_leaves * 2 * s } else { 0 }];
        let rc = unsafe {
            mr_wait(
                self.ctx.raw, batch.slot, raw.as_mut_ptr(),
                if amaf.is_empty() { ptr::null_mut() } else { amaf.as_mut_ptr() },
                if delta.is_empty() { ptr::null_mut() } else { delta.as_mut_ptr() },
                ptr::null_mut(), ptr::null_mut(), ptr::null_mut(),
            )
        };
        self.ctx.check(rc)?;
        let leaves = raw.iter().map(|r| LeafAggregate { trials: batch.k, wins: r.wins, draws: r.draws, cutoffs: r.cutoffs, sum_depth: r.sum_depth }).collect();
        Ok(BatchResults { leaves, mast_delta: delta, amaf })
    }
}

/// The tree side of the batched loop, implemented INSIDE the mcts crate (its pieces are `pub(crate)`), next to
/// `choose_action_tree_parallel` (search/parallel.rs:118-258):
///
/// ```ignore
/// impl<G: GpuGame, S: Strategy<G>> LeafSource<G> for TreeSearch<G, S> {
///     type Path = Stack;                                             // search/stack.rs
///     fn gather(&mut self, k: u32) -> (Stack, G::S, Option<G::A>, Option<G::A>) {
///         let ctx = self.select_step(..);                            // shared.rs:569-736: one virtual loss per edge
///         self.add_path_virtual_loss(&ctx.stack, k as usize - 1);    // shared.rs:787-805
///         (ctx.stack, ctx.state, ctx.prev_action, ctx.prev2_action)
///     }
///     fn apply(&mut self, path: Stack, r: &LeafAggregate, amaf: &[MrActionStat]) {
///         for (parent, idx) in path.edges() {                        // backprop.rs:704-731 summed over the k trials
///             let row = self.index.get(parent).children();
///             for p in 0..2 { row.add(idx, p, r.trials, r.utility_sum(p), r.squared_sum()); }
///             row.remove_virtual_loss(idx, r.trials);
///         }
///         self.root_stats.add(r.trials, ..);
///         self.stats.accum_depth.fetch_add(r.sum_depth as usize + r.trials as usize * (path.len() - 1), Relaxed); // shared.rs:814-818
///         // GRAVE: for every non-root path node, grave[node.hash][p][a] += amaf[p][slot(a)]      (backprop.rs:619-640)
///     }
///     fn add_mast_delta(&mut self, delta: &[MrActionStat]) { /* player_actions[p][a] += delta[p][slot(a)], backprop.rs:857-870 */ }
/// }
/// ```
pub trait LeafSource<G: GpuGame>
where
    G::A: Clone,
{
    type Path;
    /// One descent that keeps its virtual losses: (path, leaf state, last edge, the edge before it).
    fn gather(&mut self, k: u32) -> (Self::Path, G::S, Option<G::A>, Option<G::A>);
    fn apply(&mut self, path: Self::Path, result: &LeafAggregate, amaf: &[MrActionStat]);
    fn add_mast_delta(&mut self, delta: &[MrActionStat]);
    fn stats(&self) -> &TreeStats<G>;
}

/// `iterations` leaf selections in batches of `batch`, `k` playouts per leaf, up to `depth` (<= MR_NUM_SLOTS) batches in
/// flight: batch n is gathered and submitted before the results of batches n - depth + 1 .. n - 1 are applied, so the
/// host's select / backprop work overlaps the GPU.  This is csrc/tree_search.cu's loop (the measured stand-in).
pub fn run_batched<G, T>(tree: &mut T, gpu: &mut BatchedRollout<G>, iterations: usize, batch: usize, k: u32, depth: usize, max_playout_depth: usize, want_amaf: bool) -> Result<(), GpuError>
where
    G: GpuGame,
    G::A: Clone,
    T: LeafSource<G>,
{
    let depth = depth.clamp(1, MR_NUM_SLOTS);
    let s = unsafe { mr_num_slots(G::GAME_ID as i32) } as usize;
    let mut in_flight: std::collections::VecDeque<(PendingBatch, Vec<T::Path>)> = Default::default();
    let (mut done, mut submitted) = (0usize, 0usize);
    let mut drain = |tree: &mut T, gpu: &mut BatchedRollout<G>, q: &mut std::collections::VecDeque<(PendingBatch, Vec<T::Path>)>| -> Result<(), GpuError> {
        let (pending, paths) = q.pop_front().unwrap();
        let res = gpu.wait(pending)?;
        for (i, (path, r)) in paths.into_iter().zip(res.leaves.iter()).enumerate() {
            let amaf = if res.amaf.is_empty() { &res.amaf[..] } else { &res.amaf[i * 2 * s..(i + 1) * 2 * s] };
            tree.apply(path, r, amaf);
        }
        if !res.mast_delta.is_empty() {
            tree.add_mast_delta(&res.mast_delta);
        }
        Ok(())
    };
    while done < iterations {
        let nb = batch.min(iterations - done);
        let mut paths = Vec::with_capacity(nb);
        let mut leaves = Vec::with_capacity(nb);
        for _ in 0..nb {
            let (path, state, prev, prev2) = tree.gather(k);
            paths.push(path);
            leaves.push((state, prev, prev2));
        }
        let pending = gpu.submit(submitted % depth, &leaves, k, max_playout_depth, tree.stats(), want_amaf)?;
        in_flight.push_back((pending, paths));
        submitted += 1;
        done += nb;
        if in_flight.len() >= depth {
            drain(tree, gpu, &mut in_flight)?;
        }
    }
    while !in_flight.is_empty() {
        drain(tree, gpu, &mut in_flight)?;
    }
    Ok(())
}
