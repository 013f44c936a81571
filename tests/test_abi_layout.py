"""The ctypes mirror (ipc_filecoin_proofs_b200/_abi.py) must have exactly the C layout of include/ipcfp.h,
and the Rust binding source must declare every exported function."""
import ctypes as C
import os
import re
import subprocess
import tempfile

from ipc_filecoin_proofs_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

STRUCTS = {
    "ipcfp_tipset_desc": A.TipsetDesc, "ipcfp_event_spec": A.EventSpec, "ipcfp_storage_spec": A.StorageSpec, "ipcfp_witness": A.Witness,
    "ipcfp_event_proof": A.EventProofC, "ipcfp_event_result": A.EventResultC, "ipcfp_storage_proof": A.StorageProofC,
    "ipcfp_storage_result": A.StorageResultC, "ipcfp_slot_result": A.SlotResultC, "ipcfp_bundle": A.BundleC,
    "ipcfp_parsed_bundle": A.ParsedBundleC,
}


def test_ctypes_layout_matches_c_header():
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "ipcfp.h"', "int main(void) {"]
    for cname, st in STRUCTS.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in st._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["return 0; }"]
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "layout.c")
        exe = os.path.join(td, "layout")
        open(src, "w").write("\n".join(lines))
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, src])
        out = subprocess.check_output([exe], text=True)
    got = dict(l.split() for l in out.strip().splitlines())
    for cname, st in STRUCTS.items():
        assert int(got[cname]) == C.sizeof(st), cname
        for fname, _ in st._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(st, fname).offset, f"{cname}.{fname}"


def test_rust_sys_source_declares_every_export():
    hdr = open(os.path.join(ROOT, "include", "ipcfp.h")).read()
    declared = set(re.findall(r"\b(ipcfp_[a-z0-9_]+)\s*\(", hdr)) - {"ipcfp_store", "ipcfp_tipset"}
    rs = open(os.path.join(ROOT, "integration", "rust", "ipcfp-sys", "src", "lib.rs")).read()
    rust = set(re.findall(r"pub fn (ipcfp_[a-z0-9_]+)\s*\(", rs))
    assert declared == rust, (declared - rust, rust - declared)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` prints exactly one JSON line with the contract's keys (tiny workload)."""
    import json
    import sys
    env = dict(os.environ, IPCFP_BENCH_RECEIPTS="3000")
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                                  env=env, text=True, stderr=subprocess.DEVNULL)
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "receipts/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


def test_rust_shim_uses_only_declared_bindings():
    """integration/rust/gpu.rs cannot be compiled here (no Rust toolchain); at least every `sys::` item it names must exist in the -sys
    crate and every call must pass as many arguments as the declaration takes."""
    sys_rs = open(os.path.join(ROOT, "integration", "rust", "ipcfp-sys", "src", "lib.rs")).read()
    shim = open(os.path.join(ROOT, "integration", "rust", "gpu.rs")).read()
    fns = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (ipcfp_[a-z0-9_]+)\s*\(([^;]*?)\)\s*(?:->[^;]*)?;", sys_rs, re.S)}
    consts = set(re.findall(r"pub const (IPCFP_[A-Z0-9_]+)", sys_rs))
    types = set(re.findall(r"pub (?:struct|type) (ipcfp_[a-z0-9_]+)", sys_rs))
    used = set(re.findall(r"sys::([A-Za-z0-9_]+)", shim))
    assert used, "the shim names no binding at all?"
    unknown = {u for u in used if u not in fns and u not in consts and u not in types}
    assert not unknown, unknown

    def n_args(text):
        depth, n, any_tok = 0, 0, False
        for ch in text:
            if ch in "([{<":
                depth += 1
            elif ch in ")]}>":
                depth -= 1
            elif ch == "," and depth == 0:
                n += 1
                continue
            if not ch.isspace():
                any_tok = True
        return (n + 1) if any_tok and not text.rstrip().endswith(",") else n

    for m in re.finditer(r"sys::(ipcfp_[a-z0-9_]+)\s*\(", shim):
        name = m.group(1)
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(shim[i], 0)
            i += 1
        call_args = shim[m.end():i - 1].replace("->", "")
        assert n_args(call_args) == n_args(fns[name].replace("->", "")), (name, call_args)
