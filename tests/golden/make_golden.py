#!/usr/bin/env python
"""Writes tests/golden/playouts_v1.npz: frozen playout vectors of the draw discipline.

The reference itself cannot run here (Rust workspace, no cargo in the image) and its SmallRng stream is not pinned by
any of its tests, so these vectors are produced by the C oracle AFTER it has passed every golden vector the reference's
own tests hold for the rules (tests/test_oracle_golden.py).  They freeze the (seed, leaf, playout, ply) -> move mapping:
a later change of the oracle, of Philox, of the action order or of a policy's tie-break shows up as a diff against
this file on the CPU, and the CUDA path is compared with the same file on the GPU.

    python tests/golden/make_golden.py        # rewrites the fixture
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import oracle_lib as o  # noqa: E402

OUT = Path(__file__).resolve().parent / "playouts_v1.npz"
MAX_TRACE = {o.CONNECT4: 42, o.BREAKTHROUGH: 200, o.KNIGHTTHROUGH: 200, o.OTHELLO: 128, o.HEX11: 121}
POLICIES = {"uniform": o.UNIFORM, "eps_mast": o.EPS_MAST, "decisive_win": o.DECISIVE_WIN, "anti_decisive": o.DECISIVE_ANTI,
            "eps_lgr": o.EPS_LGR, "eps_nst": o.EPS_NST, "eps_lgr2": o.EPS_LGR2}
K, SEED = 12, 20260120


def cases():
    for game in range(5):
        leaves = np.concatenate([o.initial_state(game), o.random_positions(game, 5, 9, 3), o.random_positions(game, 6, 24, 3)])
        depth = 200 if game in (o.BREAKTHROUGH, o.KNIGHTTHROUGH) else 0
        base = o.playout_batch(game, leaves, 24, seed=1, max_depth=depth, want_mast_delta=True, want_bigram_delta=True,
                               want_trace=True, max_trace=4)
        # contexts for the table policies: the first move of another leaf's playout as "the action that led here"
        for i in range(len(leaves)):
            for j in range(len(leaves)):
                if j != i and leaves[j]["turn"] != leaves[i]["turn"] and base["records"][j * 24]["depth"] > 0:
                    leaves["prev"][i] = int(base["trace"][j * 24][0]) + 1
                    break
        first = o.playout_batch(game, leaves, K, policy=o.EPS_LGR2, eps=0.2, seed=2, max_depth=depth, want_final_tables=True)
        for name, policy in POLICIES.items():
            yield game, name, policy, leaves, depth, base, first


def run(game, policy, leaves, depth, base, first):
    kw = dict(policy=policy, eps=0.1, seed=SEED, max_depth=depth, want_trace=True, max_trace=MAX_TRACE[game], want_final=True,
              want_amaf=True, leaf_id_base=17)
    if policy in (o.EPS_MAST, o.EPS_NST):
        kw.update(mast=base["mast_delta"], want_mast_delta=True)
    if policy == o.EPS_NST:
        kw.update(bigram=base["bigram_delta"], nst_threshold=2, want_bigram_delta=True)
    if policy in (o.EPS_LGR, o.EPS_LGR2):
        kw.update(replies=first["replies_final"], want_reply_updates=True)
    if policy == o.EPS_LGR2:
        kw.update(replies2=first["replies2_final"], want_final_tables=True)
    return o.playout_batch(game, leaves, K, **kw)


def main():
    out = {}
    for game, name, policy, leaves, depth, base, first in cases():
        r = run(game, policy, leaves, depth, base, first)
        key = f"{o.GAME_NAMES[game]}.{name}"
        out[f"{key}.leaves"] = leaves
        out[f"{key}.trace"] = r["trace"]
        out[f"{key}.records"] = r["records"]
        out[f"{key}.results"] = r["results"]
        out[f"{key}.final"] = r["final"]
        out[f"{key}.amaf"] = r["amaf"]
        for opt in ("mast_delta", "bigram_delta", "reply_order", "reply_action", "replies2_final"):
            if r.get(opt) is not None:
                out[f"{key}.{opt}"] = r[opt]
    np.savez_compressed(OUT, **out)
    print(f"{OUT}: {len(out)} arrays, {OUT.stat().st_size} bytes")


if __name__ == "__main__":
    main()
