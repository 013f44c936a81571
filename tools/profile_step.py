#!/usr/bin/env python
"""Minimal driver for ncu: build the configs[3] tipset, ingest, run W warm-up + K resident steps.
Usage: python tools/profile_step.py [--receipts N] [--steps K] [--warmup W] [--verify]"""
import argparse
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
T0 = time.time()


def log(*a):
    print(f"[{time.time() - T0:7.2f}s]", *a, file=sys.stderr, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--receipts", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--verify", action="store_true")
    ap.add_argument("--storage", type=int, default=0, help="also run N storage-slot lookups on a 1M-entry HAMT")
    ap.add_argument("--sweep", default="", help="comma-separated pass-1 variants, e.g. MINB=8,STAGE=128x4x1,RING=128x4: every variant runs "
                                                 "warmup+steps resident steps in THIS process; prints the median device ms of pass 1")
    args = ap.parse_args()
    import synth
    from ipc_filecoin_proofs_b200 import _abi as A
    from ipc_filecoin_proofs_b200 import api
    log("imports done")
    ts = synth.Tipset(synth.config_params(4, n_receipts=args.receipts))
    log(f"tipset built: {ts.n_blocks} blocks {len(ts.blob) / 1e9:.3f} GB")
    L = api.lib()
    st = api.BlockStore.from_tipset(ts, verify_cids=args.verify)
    log("store created")
    spec = A.make_event_spec(ts.event_signature, ts.topic1, ts.actor_filter)
    d, keep = A.make_tipset_desc(ts)
    L.ipcfp_tipset_upload.restype = C.c_int32
    L.ipcfp_tipset_upload.argtypes = [C.c_void_p, C.POINTER(A.TipsetDesc), C.POINTER(C.c_void_p)]
    L.ipcfp_generate_event_proof_resident.restype = C.c_int32
    L.ipcfp_generate_event_proof_resident.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(A.EventSpec), C.c_uint32, C.POINTER(C.POINTER(A.EventResultC))]
    tip = C.c_void_p()
    assert L.ipcfp_tipset_upload(st._h, C.byref(d), C.byref(tip)) == 0
    log("tipset uploaded")
    if args.sweep:
        import numpy as np
        ref = None
        for var in args.sweep.split(","):
            for k in ("IPCFP_PASS1_STAGE", "IPCFP_PASS1_RING", "IPCFP_PASS1_MINB", "IPCFP_PASS1_TUNE", "IPCFP_PASS1_W16"):
                os.environ.pop(k, None)
            for kv in var.split("+"):
                name, val = kv.split("=")
                os.environ["IPCFP_PASS1_" + name] = val
            p1, tot = [], []
            sig = None
            for k in range(args.warmup + args.steps):
                out = C.POINTER(A.EventResultC)()
                rc = L.ipcfp_generate_event_proof_resident(st._h, tip, C.byref(spec), 0, C.byref(out))
                if rc != 0:
                    log(f"variant {var}: FAILED {rc} {L.ipcfp_last_error()}")
                    break
                r = out.contents
                if k >= args.warmup:
                    p1.append(r.ms_pass1); tot.append(r.ms_total)
                sig = (int(r.n_matching), int(r.n_proofs), int(r.witness.n_blocks), int(r.pass1_bytes))
                L.ipcfp_event_result_free(out)
            if ref is None:
                ref = sig
            print(f"SWEEP {var:24s} pass1 median {np.median(p1):.4f} ms min {np.min(p1):.4f} | step median {np.median(tot):.3f} ms | "
                  f"result {sig} {'same' if sig == ref else 'DIFFERENT from the first variant'}", flush=True)
        log("sweep done")
        return
    for k in range(args.warmup + args.steps):
        out = C.POINTER(A.EventResultC)()
        t = time.time()
        assert L.ipcfp_generate_event_proof_resident(st._h, tip, C.byref(spec), 0, C.byref(out)) == 0, L.ipcfp_last_error()
        r = out.contents
        log(f"step {k}: wall {1e3 * (time.time() - t):.2f} ms; device total {r.ms_total:.3f} txamt {r.ms_txamt:.3f} pass1 {r.ms_pass1:.3f} "
            f"pass2 {r.ms_pass2:.3f} witness {r.ms_witness:.3f}; matching {r.n_matching} witness {r.witness.n_blocks} blocks")
        L.ipcfp_event_result_free(out)
    if args.storage:
        import numpy as np
        ts3 = synth.Tipset(synth.config_params(3))
        st3 = api.BlockStore.from_tipset(ts3, verify_cids=args.verify)
        keys = [ts3.storage_entry(k)[0] for k in range(args.storage)]
        slots = np.frombuffer(b"".join(api.compute_mapping_slots(keys, [0] * len(keys))), dtype=np.uint8)
        for mode in ("fast", "strict", "fast"):
            os.environ.pop("IPCFP_HAMT_STRICT", None)
            if mode == "strict":
                os.environ["IPCFP_HAMT_STRICT"] = "1"
            for k in range(4):
                t = time.time()
                r = st3.read_storage_slots(ts3.storage_root, slots)
                print(f"STORAGE {mode} x{args.storage}: wall {1e3 * (time.time() - t):.2f} ms device {r.ms_total:.3f} ms lookup kernel {r.ms_lookup:.4f} ms nodes {r.lookup_nodes} found {int(r.found.sum())}", flush=True)
    log("done")


if __name__ == "__main__":
    main()
