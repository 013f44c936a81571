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


def test_hidden_internals_build_exports_only_the_c_abi_and_survives_a_clashing_cxx_host():
    """`make HIDE_INTERNALS=1` (linker version script csrc/exports.map): the dynamic symbol table holds exactly the functions
    include/ipcfp.h declares, and a C++ host that defines its own `ipcfp::Error` — the clash that corrupted the heap against the default
    build — gets a clean IPCFP_ERR_NO_DEVICE / a working store. Links the objects `make` already built; CPU only."""
    import shutil
    objs = [os.path.join(ROOT, "ipc_filecoin_proofs_b200", "csrc", n) for n in os.listdir(os.path.join(ROOT, "ipc_filecoin_proofs_b200", "csrc")) if n.endswith(".o")]
    if not objs or not shutil.which("nvcc") or not shutil.which("g++"):
        import pytest
        pytest.skip("objects of libipcfp.so / nvcc / g++ not available")
    with tempfile.TemporaryDirectory() as td:
        lib = os.path.join(td, "libipcfp.so")
        subprocess.check_call(["make", "-C", ROOT, "-s", "HIDE_INTERNALS=1", f"LIB_OUT={lib}", lib])
        syms = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True, check=True).stdout.split("\n")
        exported = {l.split()[-1] for l in syms if l.strip()}
        hdr = open(os.path.join(ROOT, "include", "ipcfp.h")).read()
        declared = set(re.findall(r"\b(ipcfp_[a-z0-9_]+)\s*\(", hdr)) - {"ipcfp_store", "ipcfp_tipset"}
        assert exported == declared, (sorted(exported - declared), sorted(declared - exported))
        src = os.path.join(td, "clash.cpp")
        with open(src, "w") as f:
            f.write(r'''
#include <cstdio>
#include <string>
#include "ipcfp.h"
static int g_host_dtor_calls = 0;
namespace ipcfp {   // a host that (wrongly) defines a type where the library keeps its internal error type (csrc/common.cuh). Same layout here,
// so that being interposed is harmless and can be COUNTED: every time the library destroys one of ITS exceptions through this
// destructor, the host's symbol has replaced the library's own.
struct Error {
    int status; std::string msg; unsigned long index;
    Error(int s, std::string m, unsigned long i = ~0ul) : status(s), msg(std::move(m)), index(i) {}
    ~Error() { g_host_dtor_calls++; }
};
}
int main() {
    int no_device = 0;
    for (int k = 0; k < 50; k++) {
        ipcfp_store* s = nullptr;
        unsigned char cid[38] = {1, 0x71, 0xa0, 0xe4, 2, 0x20}, blob[8] = {0x80};
        unsigned long long off = 0;
        unsigned int len = 1;
        ipcfp_status st = ipcfp_store_create(cid, (const uint64_t*)&off, &len, blob, 1, 1, 0, 0, &s);   // without a device: throws and catches its own ipcfp::Error inside
        if (st != IPCFP_OK && st != IPCFP_ERR_NO_DEVICE) { printf("status %d\n", (int)st); return 1; }
        if (st == IPCFP_OK) ipcfp_store_destroy(s); else no_device++;
    }
    const int from_library = g_host_dtor_calls;
    try { throw ipcfp::Error(3, std::string(100, 'x')); } catch (const ipcfp::Error& e) { if (e.status != 3) return 2; }
    printf("library-internal exceptions destroyed by the HOST's destructor: %d of %d\n", from_library, no_device);
    return 0;
}
''')
        exe = os.path.join(td, "clash")
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-o", exe, src, "-L" + td, "-lipcfp", "-Wl,-rpath," + td])
        out = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, MALLOC_CHECK_="3"))
        assert out.returncode == 0 and "destroyed by the HOST's destructor: 0 of" in out.stdout, (out.returncode, out.stdout, out.stderr[-2000:])
        # the same program against the DEFAULT build shows the hazard the version script removes (only visible without a device, when the
        # library throws internally): every one of its exceptions goes through the host's destructor
        default_dir = os.path.join(ROOT, "ipc_filecoin_proofs_b200")
        exe2 = os.path.join(td, "clash_default")
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-o", exe2, src, "-L" + default_dir, "-lipcfp", "-Wl,-rpath," + default_dir])
        out2 = subprocess.run([exe2], capture_output=True, text=True)
        assert out2.returncode == 0, (out2.stdout, out2.stderr[-2000:])
        n_host, n_throw = [int(x) for x in re.findall(r"(\d+) of (\d+)", out2.stdout)[0]]
        assert n_host == n_throw, out2.stdout   # documents the default build's behaviour; becomes 0 once HIDE_INTERNALS is the default
