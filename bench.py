#!/usr/bin/env python
"""bench.py — receipts/sec scanned (+ witness bytes/sec) of the event-proof hot path on B200.

  python bench.py --gpus N --steps K --warmup W            # this engine (CUDA, through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm (oracle), host cores

One "step" = one generate_event_proof over the synthetic tipset of BASELINE.json configs[3]
(1 M receipts x 8 events, 0.1 % match rate, events-AMT bit widths 3/5): message-AMT walk + execution
order, pass 1 over every receipt, pass 2 over the matches, witness sort + gather, results to the host.
N > 1 (torchrun): weak scaling — every rank holds a 1 M-receipt shard of an N x 1 M tipset
(configs[4] at N = 8), scans it, and the per-shard witness CID sets are all-gathered (NCCL) and merged.

`value`  : receipts/s with the block store and the tipset descriptor resident in HBM.
`e2e`    : the same metric through the plain C-ABI call sequence a reference-side binding makes with HOST
           buffers: ipcfp_store_create (H2D of every block, index build, Blake2b CID check) +
           ipcfp_generate_event_proof (H2D of the events roots, scan, D2H of the results).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RECEIPTS_PER_GPU = int(os.environ.get("IPCFP_BENCH_RECEIPTS", 1_000_000))


_T0 = time.time()


def log(*a):
    print(f"[{time.time() - _T0:8.2f}s]", *a, file=sys.stderr, flush=True)


def build_tipset(world, rank):
    import synth
    n_total = RECEIPTS_PER_GPU * world
    lo, hi = RECEIPTS_PER_GPU * rank, RECEIPTS_PER_GPU * (rank + 1)
    kw = {}
    if world > 1:
        kw.update(shard_lo=lo, shard_hi=hi)
    p = synth.config_params(4, n_receipts=n_total, **kw)
    t0 = time.time()
    ts = synth.Tipset(p)
    log(f"[rank {rank}] synthetic tipset: {ts.n_blocks} blocks, {len(ts.blob) / 1e9:.3f} GB, built in {time.time() - t0:.1f}s")
    return ts, lo, hi


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                mx = float(f[2])
                if t0 - 0.05 <= t <= t1 + 0.05:
                    sm.append(float(f[1]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                pass
        if not sm:  # timed region shorter than the sampling period: fall back to all samples
            for t, line in self.rows:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def pass1_traffic():
    """dram bytes per k_pass1 launch from the committed ncu --set full capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "pass1_traffic.json")) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args, world, rank):
    if rank != 0:
        return
    import oracle
    from ipc_filecoin_proofs_b200 import _abi as A
    ts, lo, hi = build_tipset(1, 0)
    spec = A.make_event_spec(ts.event_signature, ts.topic1, ts.actor_filter)
    cores = os.cpu_count() or 1
    st = oracle.Store.from_tipset(ts)
    d, keep = A.make_tipset_desc(ts)
    L = oracle.lib()

    def step():
        out = C.POINTER(A.EventResultC)()
        rc = L.oracle_generate_event_proof(st._h, C.byref(d), C.byref(spec), 0, cores, C.byref(out))
        assert rc == 0, L.oracle_last_error()
        r = out.contents
        res = (int(r.n_matching), int(r.witness.n_blocks), int(r.witness.blob_size))
        L.oracle_event_result_free(out)
        return res

    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        nm, wb, wbytes = step()
    dt = time.time() - t0
    val = ts.n_receipts * args.steps / dt
    line = {
        "impl": "reference", "metric": "receipts/sec scanned", "value": val, "unit": "receipts/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "BASELINE.json configs[3]: 1M receipts x 8 events, 0.1% match, AMT bit-widths 3/5 (generate_event_proof)",
                   "receipts": int(ts.n_receipts), "matching": nm, "witness_blocks": wb},
        "witness_bytes_per_s": wbytes * args.steps / dt,
        "cpu_baseline": {"value": val, "unit": "receipts/s", "cores": cores, "kind": "port",
                         "sample": "full workload per step; C++ restatement of the reference (the Rust crate cannot be built here), "
                                   "pass 1 parallelised over receipts on all host threads, the rest single-threaded like the reference"},
        "e2e": {"value": val, "unit": "receipts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ engine arm
def run_engine(args, world, rank, local):
    import torch
    from ipc_filecoin_proofs_b200 import _abi as A
    from ipc_filecoin_proofs_b200 import api

    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the engine has no CPU path"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = api.lib()
    ts, lo, hi = build_tipset(world, rank)
    spec = A.make_event_spec(ts.event_signature, ts.topic1, ts.actor_filter)
    N_local = hi - lo

    # pinned host copies of the flat arrays (what a binding would fill from RPC responses)
    def pinned(a):
        a = np.ascontiguousarray(a)
        pa = api.PinnedArray(a.nbytes)
        pa.array[:] = a.view(np.uint8).reshape(-1)
        return pa
    log("allocating pinned host buffers")
    p_cids, p_offs, p_lens, p_blob = pinned(ts.cids), pinned(ts.offsets), pinned(ts.lengths), pinned(ts.blob)
    p_roots, p_has = pinned(ts.events_roots), pinned(ts.has_events_root)
    log("pinned buffers ready")
    d, keep = A.make_tipset_desc(ts)
    d.events_roots = p_roots.array.ctypes.data
    d.has_events_root = p_has.array.ctypes.data
    h2d_bytes = p_cids.array.nbytes + p_offs.array.nbytes + p_lens.array.nbytes + p_blob.array.nbytes + p_roots.array.nbytes + p_has.array.nbytes

    def store_create(flags):
        h = C.c_void_p()
        rc = L.ipcfp_store_create(p_cids.array.ctypes.data, p_offs.array.ctypes.data, p_lens.array.ctypes.data, p_blob.array.ctypes.data,
                                  p_blob.array.nbytes, ts.n_blocks, local, flags, C.byref(h))
        assert rc == 0, L.ipcfp_last_error()
        return h

    # ---- resident-state objects
    store = store_create(A.STORE_VERIFY_CIDS)
    tip = C.c_void_p()
    assert L.ipcfp_tipset_upload(store, C.byref(d), C.byref(tip)) == 0, L.ipcfp_last_error()
    ext_stream = torch.cuda.ExternalStream(L.ipcfp_store_stream(store), device=torch.device("cuda", local))

    coll = ops = None
    if world > 1:
        from ipc_filecoin_proofs_b200 import parallel as PL
        coll = PL.Collectives(dist, torch.device("cuda", local))
        ops = PL.CudaShardOps(L, local)

    def run_shard(store_h, tip_h):
        """One step on this rank: local shard scan + (N > 1) cross-shard execution order and witness-CID union."""
        if world == 1:
            out = C.POINTER(A.EventResultC)()
            rc = L.ipcfp_generate_event_proof_resident(store_h, tip_h, C.byref(spec), 0, C.byref(out))
            assert rc == 0, L.ipcfp_last_error()
            return out, int(out.contents.n_exec), int(out.contents.witness.n_blocks)
        out, n_exec, merged = PL.generate_event_proof_distributed(L, store_h, tip_h, spec, lo, hi, coll, ops)
        return out, n_exec, int(merged.numel() // 38)

    stats = {}

    def step_resident():
        out, n_exec, merged = run_shard(store, tip)
        r = out.contents
        m = int(r.witness.n_blocks)
        wbytes = int(np.frombuffer((C.c_uint32 * m).from_address(r.witness.lengths), dtype=np.uint32).sum(dtype=np.uint64)) if m else 0
        stats.update(n_matching=int(r.n_matching), n_proofs=int(r.n_proofs), witness_blocks=m,
                     witness_bytes=wbytes, merged_witness_cids=merged, n_exec=n_exec,
                     ms=dict(total=r.ms_total, txamt=r.ms_txamt, pass1=r.ms_pass1, pass2=r.ms_pass2, witness=r.ms_witness),
                     pass1_bytes=int(r.pass1_bytes), pass1_nodes=int(r.pass1_nodes),
                     d2h_bytes=int(r.n_matching) * 4 + int(r.n_proofs) * C.sizeof(A.EventProofC) + int(r.data_blob_size) +
                     int(r.witness.n_blocks) * (38 + 8 + 4) + int(r.witness.blob_size))
        L.ipcfp_event_result_free(out)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- resident timing
    log("store + tipset resident; warm-up")
    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()   # every rank enters the timed region together (rank 0 just waited for the sampler)
    launches0 = api.kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pass1_ms, step_ms = [], []
    t_wall0 = time.time()
    ev0.record(ext_stream)
    step_wall = []
    for _ in range(args.steps):
        _t = time.perf_counter()
        step_resident()
        step_wall.append(1e3 * (time.perf_counter() - _t))
        pass1_ms.append(stats["ms"]["pass1"])
        step_ms.append(stats["ms"]["total"])
    ev1.record(ext_stream)
    torch.cuda.synchronize()
    t_wall1 = time.time()
    barrier()
    launches = api.kernel_launch_count() - launches0
    dev_ms = ev0.elapsed_time(ev1) if world == 1 else (t_wall1 - t_wall0) * 1e3
    t_local = torch.tensor([dev_ms, (t_wall1 - t_wall0) * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max = [float(x) for x in t_local.cpu()]
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    log(f"resident timing done: {dev_ms_max / args.steps:.3f} ms/step")
    if world > 1 and PL.PROFILE is not None and rank == 0:
        log("parallel phases (ms, median/max over calls): " + ", ".join(f"{k}={np.median(v):.2f}/{np.max(v):.1f}" for k, v in PL.PROFILE.items()))
        log("per-step wall ms: " + ", ".join(f"{x:.1f}" for x in step_wall))
    # ---- end-to-end timing (host buffers → results on the host), every step re-ingests the block set
    L.ipcfp_tipset_free(tip)
    L.ipcfp_store_destroy(store)
    e2e_steps = max(1, min(args.steps, 5))

    e2e_parts = []

    def step_e2e():
        t0 = time.time()
        h = store_create(A.STORE_VERIFY_CIDS)
        t1 = time.time()
        tp = C.c_void_p()
        assert L.ipcfp_tipset_upload(h, C.byref(d), C.byref(tp)) == 0, L.ipcfp_last_error()
        out, _, _ = run_shard(h, tp)
        t2 = time.time()
        L.ipcfp_event_result_free(out)
        L.ipcfp_tipset_free(tp)
        L.ipcfp_store_destroy(h)
        t3 = time.time()
        e2e_parts.append((1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)))

    step_e2e()
    barrier()
    t0 = time.time()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_ms = (time.time() - t0) * 1e3 / e2e_steps
    t_e2e = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_ms = float(t_e2e.cpu()[0])

    log(f"e2e timing done: {e2e_ms:.2f} ms/step; (store_create, generate, destroy) ms per step: {e2e_parts}")
    # ---- CPU baseline (rank 0, N = 1 only): the oracle, single-threaded like the reference
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        ost = oracle.Store.from_tipset(ts)
        OL = oracle.lib()
        out = C.POINTER(A.EventResultC)()
        t0 = time.time()
        rc = OL.oracle_generate_event_proof(ost._h, C.byref(d), C.byref(spec), 0, 1, C.byref(out))
        dt = time.time() - t0
        assert rc == 0
        same = int(out.contents.n_matching) == stats["n_matching"] and int(out.contents.witness.n_blocks) == stats["witness_blocks"] and \
            int(out.contents.witness.blob_size) == stats["witness_bytes"]
        OL.oracle_event_result_free(out)
        cpu_baseline = {"value": ts.n_receipts / dt, "unit": "receipts/s", "cores": 1, "kind": "port",
                        "sample": "the full 1M-receipt workload, 1 repetition, single-threaded C++ restatement of the reference "
                                  "(the reference is single-threaded; its Rust crate cannot be built in this image)",
                        "seconds": dt, "agrees_with_gpu": bool(same)}

    log("cpu baseline done")
    if rank == 0:
        n_total = N_local * world
        value = n_total * args.steps / (dev_ms_max / 1e3)
        peak, peak_src = peaks()
        p1 = float(np.mean(pass1_ms))
        achieved = stats["pass1_bytes"] / (p1 / 1e3) / 1e9
        line = {
            "metric": "receipts/sec scanned", "value": value, "unit": "receipts/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[3] per GPU: 1M receipts x 8 events, 0.1% match, events-AMT bit-widths 3/5; "
                                   "generate_event_proof (message-AMT walk + exec order, pass 1, pass 2, witness sort+gather, results to host)"
                                   + ("" if world == 1 else f"; N={world}: one {world}M-receipt tipset sharded by index range; NCCL all-to-all (distributed first-seen dedup of the execution order) + all-gather of witness CID sets"),
                       "receipts_per_gpu": N_local, "receipts_total": n_total, "store_blocks_per_gpu": int(ts.n_blocks),
                       "store_bytes_per_gpu": int(len(ts.blob)), "l2": "inputs (1.15 GB/GPU) exceed the 126 MB L2; no flush needed",
                       "matching": stats["n_matching"], "proofs": stats["n_proofs"], "witness_blocks": stats["witness_blocks"],
                       "merged_witness_cids": stats["merged_witness_cids"]},
            "witness_bytes_per_s": stats["witness_bytes"] * world * args.steps / (dev_ms_max / 1e3),
            "wall_ms_per_step": wall_ms_max / args.steps,
            "device_ms_breakdown": stats["ms"],
            "roofline": {"kernel": "k_pass1_occ8 (pass1_body, csrc/events.cu)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": pass1_traffic(), "algorithmic_bytes_per_launch": stats["pass1_bytes"], "ms_per_launch": p1,
                         "peak_source": peak_src},
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": n_total / (e2e_ms / 1e3), "unit": "receipts/s", "ms_per_step": e2e_ms, "steps": e2e_steps,
                    "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(stats["d2h_bytes"])},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    # stdout carries exactly one JSON line: everything else a library prints (e.g. NCCL's version banner) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    world, rank, local = dist_env()
    if args.impl == "reference":
        run_reference(args, world, rank)
    else:
        run_engine(args, world, rank, local)


if __name__ == "__main__":
    main()
