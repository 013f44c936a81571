#!/usr/bin/env python
"""World-1 run of the in-library sharded call (every kernel and NCCL call of the cross-shard protocol on ONE GPU) — for ncu launch lists.
Usage: python tools/profile_sharded.py [--receipts N] [--steps K]"""
import argparse
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--receipts", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    import numpy as np
    import synth
    from ipc_filecoin_proofs_b200 import _abi as A
    from ipc_filecoin_proofs_b200 import api
    from ipc_filecoin_proofs_b200 import parallel as PL
    L = api.lib()
    ts = synth.Tipset(synth.config_params(4, n_receipts=args.receipts))
    st = api.BlockStore.from_tipset(ts)
    spec = A.make_event_spec(ts.event_signature, ts.topic1, ts.actor_filter)
    d, keep = A.make_tipset_desc(ts)
    tip = C.c_void_p()
    assert L.ipcfp_tipset_upload(st._h, C.byref(d), C.byref(tip)) == 0
    comm = PL.ShardedComm(L, 1, 0, 0, PL.ShardedComm.unique_id(L))
    bounds = np.array([0, args.receipts], dtype=np.uint64)
    for k in range(args.steps):
        t = time.time()
        out = comm.generate_event_proof(st._h, tip, spec, bounds)
        r = out.contents
        print(f"SHARDED step {k}: wall {1e3 * (time.time() - t):.2f} ms; device total {r.ms_total:.3f} txamt {r.ms_txamt:.3f} pass1 {r.ms_pass1:.3f} pass2 {r.ms_pass2:.3f} "
              f"witness {r.ms_witness:.3f} exchange {r.ms_exchange:.3f} fetch {r.ms_fetch:.3f} union {r.ms_union:.3f}; n_exec {r.n_exec} union {r.n_union_cids}", flush=True)
        L.ipcfp_event_result_free(out)
    comm.close()


if __name__ == "__main__":
    main()
