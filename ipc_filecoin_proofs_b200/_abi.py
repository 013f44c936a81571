"""ctypes mirror of include/ipcfp.h (POD structs only) + converters to Python result objects.

Shared by the product binding (api.py) and by the test oracle's binding (oracle/__init__.py):
both libraries fill the same structs so parity tests compare field by field.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

CID_LEN = 38

OK = 0
ERR_INVALID_ARG = -1
ERR_MISSING_BLOCK = -2
ERR_DECODE = -3
ERR_CID_MISMATCH = -4
ERR_MISSING_EXEC = -5
ERR_CUDA = -6
ERR_NCCL = -7
ERR_STATE_ROOT_MISMATCH = -8
ERR_ACTOR_NOT_FOUND = -9
ERR_NO_DEVICE = -10
ERR_UNSUPPORTED = -11

STORE_VERIFY_CIDS = 0x1
SCAN_SKIP_TX_AMTS = 0x1
SHARDED_UNION_TO_HOST = 0x2
SHARDED_UNION_FULL = 0x4
WITNESS_BY_REFERENCE = 0x8
COMM_ID_BYTES = 128


class TipsetDesc(C.Structure):
    _fields_ = [
        ("parent_epoch", C.c_int64),
        ("child_epoch", C.c_int64),
        ("n_parents", C.c_uint32),
        ("parent_cids", C.c_void_p),
        ("parent_txmeta_cids", C.c_void_p),
        ("child_cid", C.c_void_p),
        ("receipts_root", C.c_void_p),
        ("child_parent_state_root", C.c_void_p),
        ("n_receipts", C.c_uint64),
        ("events_roots", C.c_void_p),
        ("has_events_root", C.c_void_p),
    ]


class EventSpec(C.Structure):
    _fields_ = [
        ("event_signature", C.c_char_p),
        ("topic_1", C.c_char_p),
        ("has_actor_id_filter", C.c_uint8),
        ("actor_id_filter", C.c_uint64),
    ]


class StorageSpec(C.Structure):
    _fields_ = [("actor_id", C.c_uint64), ("slot", C.c_uint8 * 32)]


class Witness(C.Structure):
    _fields_ = [
        ("n_blocks", C.c_uint64),
        ("cids", C.c_void_p),
        ("offsets", C.c_void_p),
        ("lengths", C.c_void_p),
        ("blob", C.c_void_p),
        ("blob_size", C.c_uint64),
    ]


class EventProofC(C.Structure):
    _fields_ = [
        ("exec_index", C.c_uint64),
        ("event_index", C.c_uint64),
        ("emitter", C.c_uint64),
        ("n_topics", C.c_uint32),
        ("data_len", C.c_uint32),
        ("data_off", C.c_uint64),
        ("topics_off", C.c_uint64),
        ("message_cid", C.c_uint8 * CID_LEN),
        ("_pad", C.c_uint8 * 2),
    ]


class EventResultC(C.Structure):
    _fields_ = [
        ("n_matching", C.c_uint64),
        ("matching_indices", C.c_void_p),
        ("n_proofs", C.c_uint64),
        ("proofs", C.POINTER(EventProofC)),
        ("data_blob", C.c_void_p),
        ("data_blob_size", C.c_uint64),
        ("witness", Witness),
        ("n_exec", C.c_uint64),
        ("ms_total", C.c_float),
        ("ms_pass1", C.c_float),
        ("ms_pass2", C.c_float),
        ("ms_txamt", C.c_float),
        ("ms_witness", C.c_float),
        ("pass1_bytes", C.c_uint64),
        ("pass1_nodes", C.c_uint64),
        ("shard_exec_dev", C.c_void_p),
        ("shard_exec_count", C.c_uint64),
        ("shard_raw_total", C.c_uint64),
        ("union_cids_dev", C.c_void_p),
        ("n_union_cids", C.c_uint64),
        ("union_cids", C.c_void_p),
        ("total_matching", C.c_uint64),
        ("total_proofs", C.c_uint64),
        ("ms_exchange", C.c_float),
        ("ms_fetch", C.c_float),
        ("ms_union", C.c_float),
        ("_pad0", C.c_float),
        ("union_part_first", C.c_uint64),
        ("n_union_part", C.c_uint64),
    ]


class StorageProofC(C.Structure):
    _fields_ = [
        ("actor_id", C.c_uint64),
        ("actor_state_cid", C.c_uint8 * CID_LEN),
        ("storage_root", C.c_uint8 * CID_LEN),
        ("slot", C.c_uint8 * 32),
        ("value", C.c_uint8 * 32),
        ("found", C.c_uint8),
        ("_pad", C.c_uint8 * 3),
        ("raw_len", C.c_uint32),
    ]


class ParsedBundleC(C.Structure):
    """ipcfp_parsed_bundle (ipcfp_bundle_from_json)."""
    _fields_ = [
        ("tipset", TipsetDesc),
        ("n_storage_proofs", C.c_uint64),
        ("storage_proofs", C.c_void_p),
        ("n_event_proofs", C.c_uint64),
        ("event_proofs", C.c_void_p),
        ("data_blob", C.c_void_p),
        ("data_blob_size", C.c_uint64),
        ("witness", Witness),
    ]


class StorageResultC(C.Structure):
    _fields_ = [
        ("n_proofs", C.c_uint64),
        ("proofs", C.POINTER(StorageProofC)),
        ("witness", Witness),
        ("spec_witness_offsets", C.c_void_p),
        ("spec_witness_index", C.c_void_p),
        ("ms_total", C.c_float),
    ]


class SlotResultC(C.Structure):
    _fields_ = [
        ("n", C.c_uint64),
        ("found", C.c_void_p),
        ("raw_len", C.c_void_p),
        ("values", C.c_void_p),
        ("witness", Witness),
        ("ms_total", C.c_float),
        ("ms_lookup", C.c_float),
        ("lookup_nodes", C.c_uint64),
        ("lookup_bytes", C.c_uint64),
    ]


class BundleC(C.Structure):
    _fields_ = [
        ("storage", C.POINTER(StorageResultC)),
        ("n_event_results", C.c_uint64),
        ("events", C.POINTER(C.POINTER(EventResultC))),
        ("witness", Witness),
    ]


def _arr(ptr, n, dtype):
    """Copy n items of dtype from a raw pointer into a fresh numpy array."""
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    nbytes = n * np.dtype(dtype).itemsize
    buf = (C.c_uint8 * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).copy()


@dataclass
class WitnessPy:
    cids: np.ndarray      # (m, 38) uint8, sorted in Cid Ord
    offsets: np.ndarray   # (m,) uint64
    lengths: np.ndarray   # (m,) uint32
    blob: np.ndarray      # uint8 (blocks in any order)

    @property
    def n_blocks(self):
        return len(self.cids)

    @property
    def total_bytes(self):
        return int(self.lengths.sum())

    def block(self, i):
        o = int(self.offsets[i])
        return bytes(self.blob[o:o + int(self.lengths[i])])

    def blocks(self):
        return [self.block(i) for i in range(self.n_blocks)]

    def as_dict(self):
        return {bytes(self.cids[i]): self.block(i) for i in range(self.n_blocks)}

    def as_c(self):
        """Returns (Witness struct, keepalive) for passing back into C."""
        cids = np.ascontiguousarray(self.cids)
        offs = np.ascontiguousarray(self.offsets)
        lens = np.ascontiguousarray(self.lengths)
        blob = np.ascontiguousarray(self.blob)
        w = Witness(len(cids), cids.ctypes.data, offs.ctypes.data, lens.ctypes.data, blob.ctypes.data, len(blob))
        return w, (cids, offs, lens, blob)


def witness_from_c(w):
    m = int(w.n_blocks)
    return WitnessPy(_arr(w.cids, m * CID_LEN, np.uint8).reshape(m, CID_LEN), _arr(w.offsets, m, np.uint64), _arr(w.lengths, m, np.uint32),
                     _arr(w.blob, int(w.blob_size), np.uint8))


@dataclass
class EventProofPy:
    exec_index: int
    event_index: int
    emitter: int
    topics: list          # list of 32-byte bytes
    data: bytes
    message_cid: bytes

    def key(self):
        return (self.exec_index, self.event_index, self.emitter, tuple(self.topics), self.data, self.message_cid)


@dataclass
class EventResultPy:
    matching: np.ndarray
    proofs: list
    witness: WitnessPy
    n_exec: int
    timings: dict = field(default_factory=dict)
    pass1_bytes: int = 0
    pass1_nodes: int = 0
    raw_proofs: np.ndarray = None   # packed ipcfp_event_proof records (for the verifier)
    data_blob: np.ndarray = None


def event_result_from_c(r):
    matching = _arr(r.matching_indices, int(r.n_matching), np.uint64)
    data = _arr(r.data_blob, int(r.data_blob_size), np.uint8)
    n = int(r.n_proofs)
    raw = _arr(C.cast(r.proofs, C.c_void_p).value, n * C.sizeof(EventProofC), np.uint8)
    proofs = []
    db = data.tobytes()
    for i in range(n):
        p = r.proofs[i]
        topics = [db[p.topics_off + 32 * k: p.topics_off + 32 * (k + 1)] for k in range(p.n_topics)]
        proofs.append(EventProofPy(int(p.exec_index), int(p.event_index), int(p.emitter), topics,
                                   db[p.data_off:p.data_off + p.data_len], bytes(p.message_cid)))
    return EventResultPy(matching, proofs, witness_from_c(r.witness), int(r.n_exec),
                         dict(total=r.ms_total, pass1=r.ms_pass1, pass2=r.ms_pass2, txamt=r.ms_txamt, witness=r.ms_witness),
                         int(r.pass1_bytes), int(r.pass1_nodes), raw, data)


def pack_event_proofs(proofs):
    """list of EventProofPy → (packed ipcfp_event_proof records as uint8, data blob as uint8): the layout the C ABI returns."""
    n = len(proofs)
    recs = (EventProofC * max(n, 1))()
    blob = bytearray()
    for i, p in enumerate(proofs):
        r = recs[i]
        r.exec_index, r.event_index, r.emitter = int(p.exec_index), int(p.event_index), int(p.emitter)
        r.n_topics, r.data_len = len(p.topics), len(p.data)
        r.topics_off = len(blob)
        for t in p.topics:
            blob += bytes(t)
        r.data_off = len(blob)
        blob += bytes(p.data)
        r.message_cid[:] = list(bytes(p.message_cid))
    raw = np.frombuffer(bytes(recs)[:n * C.sizeof(EventProofC)], dtype=np.uint8).copy()
    return raw, np.frombuffer(bytes(blob) + bytes(16), dtype=np.uint8).copy()


def pack_storage_proofs(proofs):
    """list of StorageProofPy → packed ipcfp_storage_proof records as uint8."""
    n = len(proofs)
    recs = (StorageProofC * max(n, 1))()
    for i, p in enumerate(proofs):
        r = recs[i]
        r.actor_id = int(p.actor_id)
        r.actor_state_cid[:] = list(bytes(p.actor_state_cid))
        r.storage_root[:] = list(bytes(p.storage_root))
        r.slot[:] = list(bytes(p.slot))
        r.value[:] = list(bytes(p.value))
        r.found, r.raw_len = int(bool(p.found)), int(p.raw_len)
    return np.frombuffer(bytes(recs)[:n * C.sizeof(StorageProofC)], dtype=np.uint8).copy()


@dataclass
class StorageProofPy:
    actor_id: int
    actor_state_cid: bytes
    storage_root: bytes
    slot: bytes
    value: bytes
    found: bool
    raw_len: int


@dataclass
class StorageResultPy:
    proofs: list
    witness: WitnessPy
    spec_witness: list      # per spec: list of indices into witness
    ms_total: float = 0.0
    raw_proofs: np.ndarray = None


def storage_result_from_c(r):
    n = int(r.n_proofs)
    proofs = []
    for i in range(n):
        p = r.proofs[i]
        proofs.append(StorageProofPy(int(p.actor_id), bytes(p.actor_state_cid), bytes(p.storage_root), bytes(p.slot), bytes(p.value),
                                     bool(p.found), int(p.raw_len)))
    offs = _arr(r.spec_witness_offsets, n + 1, np.uint64)
    idx = _arr(r.spec_witness_index, int(offs[-1]) if n else 0, np.uint32)
    spec_w = [idx[int(offs[i]):int(offs[i + 1])].tolist() for i in range(n)]
    raw = _arr(C.cast(r.proofs, C.c_void_p).value, n * C.sizeof(StorageProofC), np.uint8)
    return StorageResultPy(proofs, witness_from_c(r.witness), spec_w, float(r.ms_total), raw)


@dataclass
class SlotResultPy:
    found: np.ndarray
    raw_len: np.ndarray
    values: np.ndarray   # (n, 32)
    witness: WitnessPy
    ms_total: float = 0.0
    ms_lookup: float = 0.0
    lookup_nodes: int = 0
    lookup_bytes: int = 0


def slot_result_from_c(r):
    n = int(r.n)
    return SlotResultPy(_arr(r.found, n, np.uint8), _arr(r.raw_len, n, np.uint32), _arr(r.values, n * 32, np.uint8).reshape(n, 32),
                        witness_from_c(r.witness), float(r.ms_total), float(r.ms_lookup), int(r.lookup_nodes), int(r.lookup_bytes))


@dataclass
class BundlePy:
    storage: StorageResultPy
    events: list
    witness: WitnessPy


def bundle_from_c(b):
    st = storage_result_from_c(b.storage.contents) if b.storage else None
    ev = [event_result_from_c(b.events[i].contents) for i in range(int(b.n_event_results))]
    return BundlePy(st, ev, witness_from_c(b.witness))


class IpcfpError(RuntimeError):
    def __init__(self, status, msg, index):
        super().__init__(f"ipcfp status {status}: {msg} (index {index})")
        self.status = status
        self.msg = msg
        self.index = index


def make_tipset_desc(ts):
    """Build a TipsetDesc from any object with the synth.Tipset attribute names.
    Returns (desc, keepalive)."""
    keep = []

    def ptr(a):
        a = np.ascontiguousarray(a, dtype=np.uint8)
        keep.append(a)
        return a.ctypes.data if a.size else None

    d = TipsetDesc()
    d.parent_epoch = int(ts.parent_epoch)
    d.child_epoch = int(ts.child_epoch)
    d.n_parents = int(ts.n_parents)
    d.parent_cids = ptr(ts.parent_cids)
    d.parent_txmeta_cids = ptr(ts.parent_txmeta_cids)
    d.child_cid = ptr(ts.child_cid)
    d.receipts_root = ptr(ts.receipts_root)
    d.child_parent_state_root = ptr(ts.parent_state_root)
    d.n_receipts = int(ts.n_receipts)
    d.events_roots = ptr(ts.events_roots)
    d.has_events_root = ptr(ts.has_events_root)
    return d, keep


def make_event_spec(event_signature, topic_1, actor_id_filter=None):
    s = EventSpec()
    s.event_signature = event_signature.encode()
    s.topic_1 = topic_1.encode()
    s.has_actor_id_filter = 0 if actor_id_filter is None else 1
    s.actor_id_filter = 0 if actor_id_filter is None else int(actor_id_filter)
    return s


def make_storage_specs(specs):
    arr = (StorageSpec * len(specs))()
    for i, (actor, slot) in enumerate(specs):
        arr[i].actor_id = int(actor)
        arr[i].slot[:] = list(slot)
    return arr
