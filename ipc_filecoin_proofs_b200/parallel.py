"""Cross-shard plumbing of the event-proof path (one process per GPU, torch.distributed).

Receipts shard by index range; pass 1 / pass 2 are local (`ipcfp_generate_event_proof_shard`).
Two things span shards and are resolved here with a handful of small collectives:

  * the execution order (reference events/utils.rs:48-94: concatenate every message AMT, first
    occurrence of a CID wins) — a distributed hash join: bucketize → all-to-all → dedup →
    all-gather of the (tiny) duplicate position lists → exec index ↔ raw position arithmetic →
    fetch of the message CIDs the local proofs need;
  * the witness CID set (reference common/witness.rs:24-40 BTreeSet union) — all-gather of the
    per-shard sorted CID lists + merge (sort/unique) on the device.

The device work goes through the engine's C ABI (`ipcfp_exec_*`, `ipcfp_merge_witness_cids`); this
module only moves tensors. With the `nccl` backend collectives run on device tensors over NVLink;
with `gloo` (CPU tests, or several ranks sharing one GPU) tensors are staged through the host.
"""
import bisect
import ctypes as C

import numpy as np

from . import _abi as A

ENTRY = 48   # exec entry: 40-byte CID record + u64 global position
REC = 40


class Collectives:
    """Thin wrapper over torch.distributed that works for nccl (device tensors) and gloo (host staging)."""

    def __init__(self, dist, device=None):
        import torch
        self.torch = torch
        self.dist = dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        self.backend = dist.get_backend() if dist is not None else "none"
        self.device = device if (device is not None and self.backend == "nccl") else torch.device("cpu")

    def _t(self, a, dtype):
        return self.torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(self.device)

    def all_gather_i64(self, values):
        """values: 1-D int array of fixed length k → (world, k) numpy int64."""
        t = self._t(np.asarray(values, dtype=np.int64).reshape(-1), self.torch.int64)
        if self.world == 1:
            return t.cpu().numpy().reshape(1, -1)
        out = self.torch.empty(self.world * t.numel(), dtype=self.torch.int64, device=self.device)
        self.dist.all_gather_into_tensor(out, t)
        return out.cpu().numpy().reshape(self.world, -1)

    def all_gather_var_u64(self, values):
        """Variable-length u64 lists → concatenation over ranks (rank order)."""
        values = np.asarray(values, dtype=np.uint64).reshape(-1)
        counts = self.all_gather_i64([len(values)])[:, 0]
        cap = int(counts.max()) if len(counts) else 0
        if cap == 0:
            return np.zeros(0, dtype=np.uint64), counts
        pad = np.zeros(cap, dtype=np.int64)
        pad[:len(values)] = values.view(np.int64)
        allv = self.all_gather_i64(pad)
        return np.concatenate([allv[r, :counts[r]] for r in range(self.world)]).view(np.uint64), counts

    def all_to_all_bytes(self, send_t, chunk_bytes):
        """send_t: uint8 tensor of world*chunk_bytes (CUDA or CPU) → received tensor, same shape, same device."""
        if self.world == 1:
            return send_t
        src = send_t if send_t.device == self.device else send_t.to(self.device)
        out = self.torch.empty_like(src)
        self.dist.all_to_all_single(out, src)
        return out if out.device == send_t.device else out.to(send_t.device)

    def all_gather_bytes(self, t):
        """t: uint8 tensor (fixed size on every rank) → concatenation, on t's device."""
        if self.world == 1:
            return t
        src = t if t.device == self.device else t.to(self.device)
        out = self.torch.empty(self.world * src.numel(), dtype=self.torch.uint8, device=self.device)
        self.dist.all_gather_into_tensor(out, src)
        return out if out.device == t.device else out.to(t.device)

    def all_reduce_sum_i64(self, a):
        t = self._t(np.asarray(a, dtype=np.int64), self.torch.int64)
        if self.world > 1:
            self.dist.all_reduce(t)
        return t.cpu().numpy()

    def all_reduce_min_i64(self, v):
        t = self._t(np.asarray([v], dtype=np.int64), self.torch.int64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return int(t.cpu()[0])


class CudaShardOps:
    """Device helpers through the C ABI (torch CUDA tensors provide the buffers)."""

    def __init__(self, lib, device_index):
        import torch
        self.torch = torch
        self.L = lib
        self.dev = device_index
        self.tdev = torch.device("cuda", device_index)

    def bucketize(self, seg_ptr, nseg, pos0, world, cap):
        send = self.torch.zeros(world * cap * ENTRY, dtype=self.torch.uint8, device=self.tdev)
        counts = np.zeros(world, dtype=np.uint64)
        self.torch.cuda.synchronize(self.tdev)
        st = self.L.ipcfp_exec_bucketize(self.dev, C.c_void_p(seg_ptr), nseg, pos0, world, cap, C.c_void_p(send.data_ptr()), counts.ctypes.data)
        if st != A.OK:
            raise A.IpcfpError(st, self.L.ipcfp_last_error().decode(), self.L.ipcfp_last_error_index())
        return send, counts

    def dedup(self, recv, counts, world, cap):
        total = int(np.asarray(counts).sum())
        dup = self.torch.zeros(max(total, 1), dtype=self.torch.int64, device=self.tdev)
        n = C.c_uint64()
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        self.torch.cuda.synchronize(self.tdev)
        st = self.L.ipcfp_exec_dedup(self.dev, C.c_void_p(recv.data_ptr()), counts.ctypes.data, world, cap, C.c_void_p(dup.data_ptr()), max(total, 1),
                                     C.byref(n))
        if st != A.OK:
            raise A.IpcfpError(st, self.L.ipcfp_last_error().decode(), self.L.ipcfp_last_error_index())
        return dup[:n.value].cpu().numpy().view(np.uint64)

    def fetch(self, seg_ptr, nseg, pos0, req):
        req = np.ascontiguousarray(req, dtype=np.uint64)
        out = self.torch.zeros(max(len(req), 1) * REC, dtype=self.torch.uint8, device=self.tdev)
        if len(req):
            r = self.torch.as_tensor(req.view(np.int64)).to(self.tdev)
            self.torch.cuda.synchronize(self.tdev)
            st = self.L.ipcfp_exec_fetch(self.dev, C.c_void_p(seg_ptr), nseg, pos0, C.c_void_p(r.data_ptr()), len(req), C.c_void_p(out.data_ptr()))
            if st != A.OK:
                raise A.IpcfpError(st, self.L.ipcfp_last_error().decode(), self.L.ipcfp_last_error_index())
        return out[:len(req) * REC].cpu().numpy().reshape(-1, REC)

    def buffer_device(self):
        return self.tdev

    def merge_witness(self, gathered, counts, world, cap):
        total = int(np.asarray(counts).sum())
        out = self.torch.empty((total + 1) * 38, dtype=self.torch.uint8, device=self.tdev)
        n = C.c_uint64()
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        self.torch.cuda.synchronize(self.tdev)
        st = self.L.ipcfp_merge_witness_cids(self.dev, C.c_void_p(gathered.data_ptr()), counts.ctypes.data, world, cap, C.c_void_p(out.data_ptr()),
                                             total + 1, C.byref(n))
        if st != A.OK:
            raise A.IpcfpError(st, self.L.ipcfp_last_error().decode(), self.L.ipcfp_last_error_index())
        return out[:n.value * 38]

    def upload(self, a):
        return self.torch.as_tensor(np.ascontiguousarray(a)).to(self.tdev)


def raw_position_of(exec_index, dups_sorted):
    """exec index i ↔ position p in the concatenated message list, given the sorted duplicate positions D:
    p is the (i+1)-th position that is not in D, i.e. the fixed point of p = i + |{d ∈ D : d ≤ p}|."""
    p = exec_index
    while True:
        k = bisect.bisect_right(dups_sorted, p)
        q = exec_index + k
        if q == p:
            return p
        p = q


def resolve_execution_order(ops, coll, seg_ptr, nseg, matching, proof_exec_indices):
    """Runs the distributed first-seen dedup and returns (n_exec, {exec_index: 40-byte record}).

    seg_ptr/nseg: this rank's slice of the raw message list (device pointer for CudaShardOps);
    matching: the rank's matching receipt indices (for the MISSING_EXEC check, events/generator.rs:244-246);
    proof_exec_indices: exec indices whose message CID the rank's proofs need."""
    world, rank = coll.world, coll.rank
    counts = coll.all_gather_i64([nseg])[:, 0]
    pos0 = int(counts[:rank].sum())
    nraw = int(counts.sum())
    cap = int(counts.max()) // world + int(counts.max()) // (4 * world) + 1024
    send, cnt = ops.bucketize(seg_ptr, nseg, pos0, world, cap)
    cnt_matrix = coll.all_gather_i64(cnt.view(np.int64))          # [sender, owner]
    recv_counts = cnt_matrix[:, rank].astype(np.uint64)
    recv = coll.all_to_all_bytes(send, cap * ENTRY)
    dups_local = ops.dedup(recv, recv_counts, world, cap)
    dups, _ = coll.all_gather_var_u64(dups_local)
    D = np.sort(dups).tolist()
    n_exec = nraw - len(D)
    # exec.get(i) must exist for every matching receipt (checked in ascending order by the reference)
    bad = [int(i) for i in matching if int(i) >= n_exec]
    first_bad = coll.all_reduce_min_i64(min(bad) if bad else np.iinfo(np.int64).max)
    if first_bad != np.iinfo(np.int64).max:
        raise A.IpcfpError(A.ERR_MISSING_EXEC, "Missing message at index", first_bad)
    need = sorted(set(int(i) for i in proof_exec_indices))
    pos = np.array([raw_position_of(i, D) for i in need], dtype=np.uint64)
    req_all, req_counts = coll.all_gather_var_u64(pos)
    ans = ops.fetch(seg_ptr, nseg, pos0, req_all)                # zeros where another rank owns the position
    ans = coll.all_reduce_sum_i64(ans.reshape(-1).view(np.int64)).view(np.uint8).reshape(-1, REC)
    start = int(req_counts[:rank].sum())
    mine = ans[start:start + len(need)]
    return n_exec, {i: bytes(mine[k]) for k, i in enumerate(need)}


def record_to_cid(rec40):
    """40-byte record {digest[32], prefix[6], 0, 0} → 38-byte CID."""
    return bytes(rec40[32:38]) + bytes(rec40[:32])


def patch_message_cids(res_c, msg_of):
    """Writes EventProof.message_cid of a shard result in place (res_c: EventResultC)."""
    for k in range(int(res_c.n_proofs)):
        p = res_c.proofs[k]
        cid = record_to_cid(msg_of[int(p.exec_index)])
        C.memmove(C.addressof(p.message_cid), cid, 38)


def gather_witness_cids(ops, coll, local_sorted_cids):
    """all-gather of the per-shard sorted witness CID lists + device merge → (m, 38) uint8 tensor/array of the union."""
    local = np.ascontiguousarray(local_sorted_cids, dtype=np.uint8).reshape(-1, 38)
    counts = coll.all_gather_i64([len(local)])[:, 0].astype(np.uint64)
    cap = int(counts.max()) + 1
    buf = np.zeros((cap, 38), dtype=np.uint8)
    buf[:len(local)] = local
    mine = ops.upload(buf.reshape(-1))
    gathered = coll.all_gather_bytes(mine)
    merged = ops.merge_witness(gathered, counts, coll.world, cap)
    return merged


def generate_event_proof_distributed(lib, store_handle, tipset_handle, spec_c, lo, hi, coll, ops, flags=0):
    """One rank's part of a sharded generate_event_proof + the cross-shard resolution.
    Returns (POINTER(EventResultC) with message CIDs patched — caller frees it, n_exec, merged witness CIDs)."""
    out = C.POINTER(A.EventResultC)()
    st = lib.ipcfp_generate_event_proof_shard_resident(store_handle, tipset_handle, C.byref(spec_c), lo, hi, coll.world, coll.rank, flags, C.byref(out))
    # a failing rank must not leave the others hanging in a collective: agree on the status first
    worst = coll.all_reduce_min_i64(st)
    if worst != A.OK:
        if st != A.OK:
            raise A.IpcfpError(st, lib.ipcfp_last_error().decode(errors="replace"), lib.ipcfp_last_error_index())
        raise A.IpcfpError(worst, "another rank failed", 0xFFFFFFFFFFFFFFFF)
    r = out.contents
    matching = np.frombuffer((C.c_uint64 * int(r.n_matching)).from_address(r.matching_indices), dtype=np.uint64) if r.n_matching else np.zeros(0, np.uint64)
    need = [int(r.proofs[k].exec_index) for k in range(int(r.n_proofs))]
    n_exec, msg_of = resolve_execution_order(ops, coll, r.shard_exec_dev, int(r.shard_exec_count), matching, need)
    patch_message_cids(r, msg_of)
    m = int(r.witness.n_blocks)
    local_cids = np.frombuffer((C.c_uint8 * (m * 38)).from_address(r.witness.cids), dtype=np.uint8) if m else np.zeros(0, np.uint8)
    merged = gather_witness_cids(ops, coll, local_cids)
    return out, n_exec, merged
