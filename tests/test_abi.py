"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol the
header declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_library_builds_and_exports_every_declared_symbol():
    from mcts_b200 import build

    build.build()
    import mcts_b200

    lib = mcts_b200.load()
    header = (ROOT / "include" / "mcts_rollout.h").read_text()
    declared = set(re.findall(r"\b(mr_[a-z0-9_]+)\s*\(", header))
    declared -= {"mr_ctx"}
    assert declared == set(mcts_b200.EXPORTS), declared ^ set(mcts_b200.EXPORTS)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.mr_abi_version() == 2
    assert [lib.mr_num_actions(g) for g in range(5)] == [7, 4096, 4096, 65, 121]
    assert lib.mr_num_actions(9) == 0


def test_struct_layouts_match_header():
    import mcts_b200
    from mcts_b200._lib import BatchDesc

    assert mcts_b200.LEAF_DTYPE.itemsize == 40
    assert mcts_b200.RESULT_DTYPE.itemsize == 24
    assert mcts_b200.STAT_DTYPE.itemsize == 8
    assert mcts_b200.RECORD_DTYPE.itemsize == 8
    assert C.sizeof(BatchDesc) == 48
    assert BatchDesc.seed.offset == 24 and BatchDesc.epsilon.offset == 40


def test_no_cpu_fallback_without_a_device():
    import torch

    import mcts_b200

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path cannot be exercised")
    with pytest.raises(mcts_b200.RolloutError) as e:
        mcts_b200.GpuRollout(mcts_b200.Game.CONNECT4)
    assert "MR_ERR_NO_DEVICE" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under mcts_b200/ or include/ may import, link or name it."""
    bad = []
    for p in list((ROOT / "mcts_b200").rglob("*.py")) + list((ROOT / "mcts_b200" / "csrc").glob("*")) + list((ROOT / "include").glob("*")):
        if p.is_file() and p.suffix in {".py", ".cu", ".cuh", ".h", ".cpp"}:
            txt = p.read_text()
            if re.search(r"oracle_lib|liboracle|oracle/|orc_", txt):
                bad.append(str(p))
    assert not bad, bad


def test_initial_states_match_oracle():
    import mcts_b200 as mb
    import oracle_lib as o

    for g in range(5):
        assert mb.initial_state(mb.Game(g)).tobytes() == o.initial_state(g).tobytes()


def test_action_slots_match_oracle():
    """mr_num_slots / mr_action_slot / mr_slot_action (host functions, no GPU): the compact per-player action index of
    the bigram tables, against the oracle's, and slot -> action -> slot round trips."""
    import mcts_b200 as mb
    import oracle_lib as o
    from mcts_b200.rollout import action_slot, num_slots, slot_action

    for g in range(5):
        ns = num_slots(mb.Game(g))
        assert ns == o.num_slots(g) == [7, 192, 512, 65, 121][g]
        for p in range(2):
            real = 0
            for slot in range(ns):
                code = slot_action(mb.Game(g), p, slot)
                if code < 0:
                    continue
                real += 1
                assert action_slot(mb.Game(g), p, code) == slot == o.action_slot(g, p, code)
            assert real == [7, 154, 336, 65, 121][g]  # 7*8*3 - 2*7 diagonal-off-board; knight jumps on an 8x8 board
            assert slot_action(mb.Game(g), p, ns) == -1 and action_slot(mb.Game(g), p, 0xFFFF) == -1
        # every action the rules generate has a slot
        for s in o.random_positions(g, 1, 9, 6):
            st = np.array([s])
            for a in o.generate_actions(g, st):
                assert action_slot(mb.Game(g), int(s["turn"]), a) >= 0


def _compile_host_example(tmp_path):
    import subprocess

    from mcts_b200 import build

    build.build()
    exe = tmp_path / "host_loop"
    subprocess.run(
        ["gcc", "-O2", "-std=c99", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "examples" / "host_loop.c"),
         f"-L{build.OUT_DIR}", "-lmcts_b200", f"-Wl,-rpath,{build.OUT_DIR}", "-o", str(exe)],
        check=True,
    )
    return exe


def test_header_is_plain_c_and_struct_sizes(tmp_path):
    """include/mcts_rollout.h compiles as C99 (no C++ / torch types at the boundary) and its PODs have the sizes the
    bindings assume."""
    import subprocess

    src = tmp_path / "sizes.c"
    src.write_text(
        '#include <stdio.h>\n#include "mcts_rollout.h"\n'
        "int main(void) { printf(\"%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(mr_leaf), sizeof(mr_leaf_result),\n"
        "  sizeof(mr_action_stat), sizeof(mr_playout_record), sizeof(mr_batch_desc), sizeof(mr_search_desc),\n"
        "  sizeof(mr_root_child), sizeof(mr_search_stats), sizeof(mr_reply_update)); return 0; }\n"
    )
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    from mcts_b200._lib import BatchDesc, SearchDesc, SearchStats, REPLY_UPDATE_DTYPE, ROOT_CHILD_DTYPE

    assert [int(x) for x in out] == [40, 24, 8, 8, C.sizeof(BatchDesc), C.sizeof(SearchDesc), ROOT_CHILD_DTYPE.itemsize,
                                     C.sizeof(SearchStats), REPLY_UPDATE_DTYPE.itemsize]


def test_c_host_example_builds_and_fails_loudly_without_a_gpu(tmp_path):
    """examples/host_loop.c: a plain-C caller of the ABI.  Without a device it must stop at mr_init with
    MR_ERR_NO_DEVICE (exit code 3), never compute on the CPU."""
    import subprocess

    import torch

    exe = _compile_host_example(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_host_example_runs_on_the_gpu(tmp_path):
    import subprocess

    exe = _compile_host_example(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "best column 3" in r.stdout
