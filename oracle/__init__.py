"""ctypes front-end of the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE — parity unpinned by the reference (see oracle/oracle.h). Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from ipc_filecoin_proofs_b200 import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build_lib(force=False):
    srcs = [os.path.join(_HERE, "oracle.cpp"), os.path.join(_HERE, "oracle.h"), os.path.join(_HERE, "..", "synth", "cpu_crypto.h"),
            os.path.join(_HERE, "..", "include", "ipcfp.h")]
    if not force and os.path.exists(_LIB_PATH) and all(
            os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs if os.path.exists(s)):
        return _LIB_PATH
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", _LIB_PATH,
                           os.path.join(_HERE, "oracle.cpp")])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_lib()
        L = C.CDLL(_LIB_PATH)
        L.oracle_store_create.restype = C.c_void_p
        L.oracle_store_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.oracle_store_destroy.argtypes = [C.c_void_p]
        L.oracle_store_verify_cids.restype = C.c_uint64
        L.oracle_store_verify_cids.argtypes = [C.c_void_p, C.c_uint32]
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_last_error_index.restype = C.c_uint64
        L.oracle_generate_event_proof.restype = C.c_int32
        L.oracle_generate_event_proof.argtypes = [C.c_void_p, C.POINTER(A.TipsetDesc), C.POINTER(A.EventSpec), C.c_uint32, C.c_uint32,
                                                  C.POINTER(C.POINTER(A.EventResultC))]
        L.oracle_generate_event_proof_shard.restype = C.c_int32
        L.oracle_generate_event_proof_shard.argtypes = [C.c_void_p, C.POINTER(A.TipsetDesc), C.POINTER(A.EventSpec), C.c_uint64,
                                                        C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                        C.POINTER(C.POINTER(A.EventResultC))]
        L.oracle_event_result_free.argtypes = [C.POINTER(A.EventResultC)]
        L.oracle_read_storage_slots.restype = C.c_int32
        L.oracle_read_storage_slots.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.POINTER(A.SlotResultC))]
        L.oracle_slot_result_free.argtypes = [C.POINTER(A.SlotResultC)]
        L.oracle_generate_storage_proofs.restype = C.c_int32
        L.oracle_generate_storage_proofs.argtypes = [C.c_void_p, C.POINTER(A.TipsetDesc), C.c_void_p, C.c_uint64,
                                                     C.POINTER(C.POINTER(A.StorageResultC))]
        L.oracle_storage_result_free.argtypes = [C.POINTER(A.StorageResultC)]
        L.oracle_generate_proof_bundle.restype = C.c_int32
        L.oracle_generate_proof_bundle.argtypes = [C.c_void_p, C.POINTER(A.TipsetDesc), C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                                   C.POINTER(C.POINTER(A.BundleC))]
        L.oracle_bundle_free.argtypes = [C.POINTER(A.BundleC)]
        L.oracle_verify_event_proofs.restype = C.c_int32
        L.oracle_verify_event_proofs.argtypes = [C.POINTER(A.Witness), C.POINTER(A.TipsetDesc), C.c_void_p, C.c_uint64, C.c_void_p,
                                                 C.POINTER(A.EventSpec), C.c_void_p]
        L.oracle_verify_storage_proofs.restype = C.c_int32
        L.oracle_verify_storage_proofs.argtypes = [C.POINTER(A.Witness), C.POINTER(A.TipsetDesc), C.c_void_p, C.c_uint64, C.c_void_p]
        for name in ("oracle_keccak256", "oracle_blake2b256", "oracle_sha256"):
            getattr(L, name).argtypes = [C.c_char_p, C.c_uint64, C.c_void_p]
        L.oracle_compute_mapping_slot.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p]
        L.oracle_sort_unique_cids.restype = C.c_uint64
        L.oracle_sort_unique_cids.argtypes = [C.c_void_p, C.c_uint64]
        _lib = L
    return _lib


def _check(st):
    if st != A.OK:
        L = lib()
        raise A.IpcfpError(st, L.oracle_last_error().decode(errors="replace"), L.oracle_last_error_index())


class Store:
    """MemoryBlockstore over flat arrays (borrowed: the arrays are kept alive here)."""

    def __init__(self, cids, offsets, lengths, blob):
        self._keep = (np.ascontiguousarray(cids, dtype=np.uint8), np.ascontiguousarray(offsets, dtype=np.uint64),
                      np.ascontiguousarray(lengths, dtype=np.uint32), np.ascontiguousarray(blob, dtype=np.uint8))
        c, o, l, b = self._keep
        self.n_blocks = len(l)
        self._h = lib().oracle_store_create(c.ctypes.data, o.ctypes.data, l.ctypes.data, b.ctypes.data, self.n_blocks)

    @classmethod
    def from_tipset(cls, ts):
        return cls(ts.cids, ts.offsets, ts.lengths, ts.blob)

    def verify_cids(self, threads=1):
        r = lib().oracle_store_verify_cids(self._h, threads)
        return None if r == 0xFFFFFFFFFFFFFFFF else int(r)

    def generate_event_proof(self, ts, spec, flags=0, threads=1):
        d, keep = A.make_tipset_desc(ts)
        out = C.POINTER(A.EventResultC)()
        _check(lib().oracle_generate_event_proof(self._h, C.byref(d), C.byref(spec), flags, threads, C.byref(out)))
        try:
            return A.event_result_from_c(out.contents)
        finally:
            lib().oracle_event_result_free(out)

    def generate_event_proof_shard(self, ts, spec, lo, hi, world, rank, flags=0, threads=1):
        d, keep = A.make_tipset_desc(ts)
        out = C.POINTER(A.EventResultC)()
        _check(lib().oracle_generate_event_proof_shard(self._h, C.byref(d), C.byref(spec), lo, hi, world, rank, flags, threads,
                                                       C.byref(out)))
        try:
            return A.event_result_from_c(out.contents)
        finally:
            lib().oracle_event_result_free(out)

    def read_storage_slots(self, root, slots):
        root = np.ascontiguousarray(root, dtype=np.uint8)
        slots = np.ascontiguousarray(slots, dtype=np.uint8).reshape(-1, 32)
        out = C.POINTER(A.SlotResultC)()
        _check(lib().oracle_read_storage_slots(self._h, root.ctypes.data, slots.ctypes.data, len(slots), C.byref(out)))
        try:
            return A.slot_result_from_c(out.contents)
        finally:
            lib().oracle_slot_result_free(out)

    def generate_storage_proofs(self, ts, specs):
        d, keep = A.make_tipset_desc(ts)
        arr = A.make_storage_specs(specs)
        out = C.POINTER(A.StorageResultC)()
        _check(lib().oracle_generate_storage_proofs(self._h, C.byref(d), arr, len(specs), C.byref(out)))
        try:
            return A.storage_result_from_c(out.contents)
        finally:
            lib().oracle_storage_result_free(out)

    def generate_proof_bundle(self, ts, storage_specs, event_specs):
        d, keep = A.make_tipset_desc(ts)
        sarr = A.make_storage_specs(storage_specs)
        earr = (A.EventSpec * len(event_specs))(*event_specs)
        out = C.POINTER(A.BundleC)()
        _check(lib().oracle_generate_proof_bundle(self._h, C.byref(d), sarr, len(storage_specs), earr, len(event_specs), C.byref(out)))
        try:
            return A.bundle_from_c(out.contents)
        finally:
            lib().oracle_bundle_free(out)

    def close(self):
        if self._h:
            lib().oracle_store_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def verify_event_proofs(witness, ts, result, filter_spec=None):
    """Restated events/verifier.rs over a WitnessPy; returns a list of bools."""
    d, keep = A.make_tipset_desc(ts)
    w, keep2 = witness.as_c()
    n = len(result.proofs)
    res = np.zeros(n, dtype=np.uint8)
    raw = np.ascontiguousarray(result.raw_proofs)
    blob = np.ascontiguousarray(result.data_blob)
    _check(lib().oracle_verify_event_proofs(C.byref(w), C.byref(d), raw.ctypes.data if n else None, n,
                                            blob.ctypes.data if blob.size else None,
                                            C.byref(filter_spec) if filter_spec is not None else None, res.ctypes.data))
    return [bool(x) for x in res]


def verify_storage_proofs(witness, ts, result):
    d, keep = A.make_tipset_desc(ts)
    w, keep2 = witness.as_c()
    n = len(result.proofs)
    res = np.zeros(n, dtype=np.uint8)
    raw = np.ascontiguousarray(result.raw_proofs)
    _check(lib().oracle_verify_storage_proofs(C.byref(w), C.byref(d), raw.ctypes.data if n else None, n, res.ctypes.data))
    return [bool(x) for x in res]


def _hash(fn, data):
    out = (C.c_uint8 * 32)()
    getattr(lib(), fn)(bytes(data), len(data), out)
    return bytes(out)


def keccak256(data):
    return _hash("oracle_keccak256", data)


def blake2b256(data):
    return _hash("oracle_blake2b256", data)


def sha256(data):
    return _hash("oracle_sha256", data)


def compute_mapping_slot(key32, slot_index):
    out = (C.c_uint8 * 32)()
    lib().oracle_compute_mapping_slot(bytes(key32), slot_index, out)
    return bytes(out)


def sort_unique_cids(cids):
    a = np.ascontiguousarray(cids, dtype=np.uint8).reshape(-1, 38).copy()
    n = lib().oracle_sort_unique_cids(a.ctypes.data, len(a))
    return a[:n]
