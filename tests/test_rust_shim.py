"""The Rust side of the boundary (rust/mcts-gpu) cannot be compiled in this image (no cargo / rustc), so what CAN be
checked is checked: the `extern "C"` binding is generated from include/mcts_rollout.h and must match the committed file;
it declares every symbol the library exports; its #[repr(C)] structs have the sizes and field offsets the C compiler
gives the header's; and the hand-written lib.rs only calls functions and names constants the binding declares."""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
FFI = ROOT / "rust" / "mcts-gpu" / "src" / "ffi.rs"
LIB = ROOT / "rust" / "mcts-gpu" / "src" / "lib.rs"

RUST_SIZES = {"u8": 1, "u16": 2, "u32": 4, "u64": 8, "i32": 4, "i64": 8, "f32": 4, "f64": 8, "c_int": 4, "usize": 8}


def test_ffi_rs_is_the_generators_output():
    import gen_rust_ffi

    assert FFI.read_text() == gen_rust_ffi.generate(), "include/mcts_rollout.h changed: run python tools/gen_rust_ffi.py"


def test_ffi_rs_declares_every_exported_symbol():
    import mcts_b200

    declared = set(re.findall(r"pub fn (mr_\w+)\(", FFI.read_text()))
    assert declared == set(mcts_b200.EXPORTS), declared ^ set(mcts_b200.EXPORTS)


def _rust_layout(fields):
    """#[repr(C)] layout of (name, type) fields: offsets and total size, natural alignment."""
    off, align_max, out = 0, 1, {}
    for name, ty in fields:
        m = re.fullmatch(r"\[(\w+); (\d+)\]", ty)
        base, count = (m.group(1), int(m.group(2))) if m else (ty, 1)
        size = RUST_SIZES[base]
        off = (off + size - 1) // size * size
        out[name] = off
        off += size * count
        align_max = max(align_max, size)
    return out, (off + align_max - 1) // align_max * align_max


def test_ffi_struct_layouts_match_the_c_compiler(tmp_path):
    text = FFI.read_text()
    structs = {}
    for m in re.finditer(r"pub struct (Mr\w+) \{\n(.*?)\n\}", text, flags=re.S):
        fields = re.findall(r"pub (\w+): ([^,\n]+),", m.group(2))
        if fields:
            structs[m.group(1)] = fields
    c_names = {"MrLeaf": "mr_leaf", "MrLeafResult": "mr_leaf_result", "MrActionStat": "mr_action_stat", "MrPlayoutRecord": "mr_playout_record",
               "MrBatchDesc": "mr_batch_desc", "MrReplyUpdate": "mr_reply_update", "MrSearchDesc": "mr_search_desc", "MrRootChild": "mr_root_child",
               "MrSearchStats": "mr_search_stats", "MrEdgeStats": "mr_edge_stats"}
    assert set(structs) == set(c_names)
    prog = ['#include <stdio.h>', '#include <stddef.h>', '#include "mcts_rollout.h"', "int main(void) {"]
    for rs, c in c_names.items():
        prog.append(f'  printf("{rs} size %zu\\n", sizeof({c}));')
        for fname, _ in structs[rs]:
            prog.append(f'  printf("{rs} {fname} %zu\\n", offsetof({c}, {fname}));')
    prog += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", f"-I{ROOT / 'include'}", str(src), "-o", str(exe)], check=True)
    got = {}
    for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines():
        s, f, v = line.split()
        got[(s, f)] = int(v)
    for rs, fields in structs.items():
        offsets, size = _rust_layout(fields)
        assert size == got[(rs, "size")], (rs, size, got[(rs, "size")])
        for fname, off in offsets.items():
            assert off == got[(rs, fname)], (rs, fname, off, got[(rs, fname)])


def test_lib_rs_uses_only_what_the_binding_declares():
    ffi, lib = FFI.read_text(), LIB.read_text()
    declared_fns = set(re.findall(r"pub fn (mr_\w+)\(", ffi))
    declared_consts = set(re.findall(r"pub const (MR_\w+):", ffi))
    used_fns = set(re.findall(r"\b(mr_[a-z0-9_]+)\(", lib))
    used_consts = set(re.findall(r"\b(MR_[A-Z0-9_]+)\b", lib))
    assert used_fns <= declared_fns, used_fns - declared_fns
    assert used_consts <= declared_consts, used_consts - declared_consts
    # the drop-in implements the reference's trait with the reference's signature (simulate.rs:84-91)
    assert re.search(r"impl<G: GpuGame> SimulateStrategy<G> for GpuRollout<G>", lib)
    assert re.search(r"fn playout\(&mut self, state: G::S, max_playout_depth: usize, stats: &TreeStats<G>, prev_action: Option<G::A>, rng: &mut SmallRng\) -> Trial<G>", lib)
    assert "fn backprop_flags(&self) -> BackpropFlags" in lib
    for game in ("c4::Standard", "bt::Breakthrough<8, 8>", "kt::Knightthrough<8, 8>", "oth::Othello", "hex::Hex<11, 2>"):
        assert f"impl GpuGame for {game}" in lib, game
