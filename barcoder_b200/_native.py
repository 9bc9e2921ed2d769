"""ctypes binding of libbarcoder_b200.so (C ABI: include/barcoder_b200.h).

This is the only compute path of the package.  There is no CPU fallback: if the shared
library is missing, cannot be loaded, or no CUDA device is usable, the calls raise
(``NativeLibraryError`` / ``BowtieError`` upstream) instead of degrading.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BARCODER_B200_LIB") or os.path.join(_PKG, "libbarcoder_b200.so")

BC_OK = 0
BC_EINVAL, BC_ECUDA, BC_ENODEV, BC_ENOMEM, BC_ELIMIT = -1, -2, -3, -4, -5
BC_PAM_IUPAC = 1
BC_PAM_GATE = 2
BC_PARAM_BLOCKS = 1
BC_PARAM_PATH = 2
BC_PARAM_COUNT_CANDIDATES = 3
BC_PARAM_HIT_CAPACITY = 4
BC_PARAM_SPACER_ID_BASE = 5
BC_PARAM_SCAN_PART = 6
BC_PARAM_WINDOW_SORT = 7
BC_PARAM_JOIN_CHUNK = 8
BC_PARAM_KEY_NT = 9
BC_PARAM_SLOT_PART = 10
BC_PARAM_INDEX_SORT = 11
BC_PARAM_KEY_CAP = 12
BC_PARAM_COMPACT_DIR = 13
PATH_AUTO, PATH_PROBE, PATH_JOIN, PATH_CJOIN = 0, 1, 2, 3

META_PAM_OK = 1 << 3
META_PAM_FULL = 1 << 4
META_PAM_AMB = 1 << 5

HIT_DTYPE = np.dtype([("spacer_id", "<u4"), ("gpos", "<u4"), ("mm_mask", "<u4"), ("meta", "<u4")])

EXPORTS = (
    "bc_abi_version", "bc_device_count", "bc_create", "bc_destroy", "bc_set_genome", "bc_set_genome_dev", "bc_set_library",
    "bc_set_library_dev", "bc_set_pam", "bc_set_param", "bc_build_index", "bc_search", "bc_copy_hits",
    "bc_set_hit_sink", "bc_sort_hits", "bc_set_slice_callback", "bc_peer_export", "bc_peer_open", "bc_peer_close", "bc_hits_device", "bc_get_stats", "bc_last_error", "bc_enumerate_guides", "bc_copy_guides",
)


# void fn(void* user, const bc_hit* d_hits, uint64_t begin, uint64_t end)
SLICE_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64)


class NativeLibraryError(RuntimeError):
    pass


class BcStats(ctypes.Structure):
    _fields_ = [
        ("genome_bases", ctypes.c_uint64), ("library_spacers", ctypes.c_uint64),
        ("spacer_len", ctypes.c_uint32), ("k", ctypes.c_uint32), ("blocks", ctypes.c_uint32),
        ("combos", ctypes.c_uint32), ("path", ctypes.c_uint32), ("scan_launches", ctypes.c_uint32),
        ("hits", ctypes.c_uint64), ("candidates", ctypes.c_uint64), ("probes", ctypes.c_uint64),
        ("ms_pack_genome", ctypes.c_float), ("ms_pack_library", ctypes.c_float),
        ("ms_build_index", ctypes.c_float), ("ms_search", ctypes.c_float),
        ("ms_scan_kernel", ctypes.c_float), ("ms_genome_bucket", ctypes.c_float),
        ("index_launches", ctypes.c_uint32), ("key_nt", ctypes.c_uint32), ("ms_sort_hits", ctypes.c_float),
        ("ms_win_count", ctypes.c_float), ("ms_win_bin", ctypes.c_float), ("ms_win_place", ctypes.c_float),
        ("ms_finish", ctypes.c_float), ("search_attempts", ctypes.c_uint32), ("reserved0", ctypes.c_uint32),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None


def load():
    """Load the extension once; raise loudly if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C barcoder_b200/csrc`. barcoder_b200 has no CPU fallback."
        )
    try:
        L = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover - depends on the box
        raise NativeLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    vp, u8p, u32, u64, i32, i64 = (ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64,
                                   ctypes.c_int, ctypes.c_int64)
    L.bc_abi_version.restype = i32
    L.bc_device_count.restype = i32
    L.bc_create.argtypes = [ctypes.POINTER(vp), i32]
    L.bc_destroy.argtypes = [vp]
    L.bc_destroy.restype = None
    L.bc_set_genome.argtypes = [vp, u8p, vp, u32]
    L.bc_set_genome_dev.argtypes = [vp, u8p, vp, u32, vp]
    L.bc_set_library.argtypes = [vp, u8p, u32, u32]
    L.bc_set_library_dev.argtypes = [vp, u8p, u32, u32, vp]
    L.bc_set_pam.argtypes = [vp, ctypes.c_char_p, i32, u32]
    L.bc_set_param.argtypes = [vp, i32, i64]
    L.bc_build_index.argtypes = [vp, i32]
    L.bc_search.argtypes = [vp, i32, ctypes.POINTER(u64)]
    L.bc_copy_hits.argtypes = [vp, vp, u64]
    L.bc_set_hit_sink.argtypes = [vp, vp, u64]
    L.bc_sort_hits.argtypes = [vp, i32]
    L.bc_set_slice_callback.argtypes = [vp, SLICE_FN, vp]
    L.bc_peer_export.argtypes = [vp, u64, ctypes.POINTER(vp), ctypes.c_char_p]
    L.bc_peer_open.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(vp)]
    L.bc_peer_close.argtypes = [vp, vp, ctypes.c_int]
    L.bc_hits_device.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(u64)]
    L.bc_get_stats.argtypes = [vp, ctypes.POINTER(BcStats)]
    L.bc_enumerate_guides.argtypes = [vp, u32, ctypes.c_char_p, i32, u32, ctypes.POINTER(u64)]
    L.bc_copy_guides.argtypes = [vp, vp, u64]
    L.bc_last_error.argtypes = [vp]
    L.bc_last_error.restype = ctypes.c_char_p
    for name in EXPORTS:
        if name not in ("bc_destroy", "bc_last_error"):
            getattr(L, name).restype = i32
    _lib = L
    return L


def device_count():
    """Usable CUDA devices."""
    return int(load().bc_device_count())


class NativeError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"barcoder_b200 native error {code}: {message}")
        self.code = code
        self.message = message


class Searcher:
    """Thin object wrapper over one bc_ctx (one GPU)."""

    def __init__(self, device=0):
        self._L = load()
        self._ctx = ctypes.c_void_p()
        rc = self._L.bc_create(ctypes.byref(self._ctx), int(device))
        if rc != BC_OK:
            msg = self._L.bc_last_error(None)
            self._ctx = None
            raise NativeError(rc, (msg or b"").decode())
        self.device = device
        self.contig_offsets = None

    # -- lifetime
    def close(self):
        if getattr(self, "_ctx", None):
            self._L.bc_destroy(self._ctx)
            self._ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != BC_OK:
            raise NativeError(rc, (self._L.bc_last_error(self._ctx) or b"").decode())

    # -- inputs
    def set_genome(self, contigs):
        """contigs: list of str/bytes.  Host buffers; the copy to the device is inside the call."""
        bs = [c.encode("ascii") if isinstance(c, str) else bytes(c) for c in contigs]
        off = np.zeros(len(bs) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
        blob = np.frombuffer(b"".join(bs), dtype=np.uint8) if off[-1] else np.zeros(1, np.uint8)
        self.set_genome_array(blob, off)

    def set_genome_array(self, ascii_u8, offsets_u64):
        ascii_u8 = np.ascontiguousarray(ascii_u8, dtype=np.uint8)
        off = np.ascontiguousarray(offsets_u64, dtype=np.uint64)
        self._check(self._L.bc_set_genome(self._ctx, ascii_u8.ctypes.data, off.ctypes.data, len(off) - 1))
        self.contig_offsets = off.copy()

    def set_genome_device(self, data_ptr, offsets_u64, stream=0):
        off = np.ascontiguousarray(offsets_u64, dtype=np.uint64)
        self._check(self._L.bc_set_genome_dev(self._ctx, data_ptr, off.ctypes.data, len(off) - 1, stream or None))
        self.contig_offsets = off.copy()

    def set_library(self, spacers):
        """spacers: list of equal-length str, or a uint8 array [n, L]."""
        if isinstance(spacers, np.ndarray):
            arr = np.ascontiguousarray(spacers, dtype=np.uint8)
            n, L = arr.shape
        else:
            spacers = list(spacers)
            n = len(spacers)
            L = len(spacers[0]) if n else 1
            if any(len(s) != L for s in spacers):
                raise ValueError("all spacers passed to one set_library call must have the same length")
            arr = np.frombuffer("".join(spacers).encode("ascii"), dtype=np.uint8) if n else np.zeros(1, np.uint8)
        self._check(self._L.bc_set_library(self._ctx, arr.ctypes.data, n, L))
        self.n, self.L = n, L

    def set_library_device(self, data_ptr, n, L, stream=0):
        self._check(self._L.bc_set_library_dev(self._ctx, data_ptr, n, L, stream or None))
        self.n, self.L = n, L

    def set_pam(self, pam="", direction="downstream", iupac=False, gate=False):
        d = {"downstream": 0, "upstream": 1}.get(direction)
        if d is None:
            raise ValueError("direction must be 'upstream' or 'downstream'")
        flags = (BC_PAM_IUPAC if iupac else 0) | (BC_PAM_GATE if gate else 0)
        self._check(self._L.bc_set_pam(self._ctx, pam.upper().encode("ascii"), d, flags))

    def set_param(self, key, value):
        self._check(self._L.bc_set_param(self._ctx, key, int(value)))

    # -- compute
    def build_index(self, k):
        self._check(self._L.bc_build_index(self._ctx, int(k)))

    def search(self, k):
        n = ctypes.c_uint64()
        self._check(self._L.bc_search(self._ctx, int(k), ctypes.byref(n)))
        return n.value

    def sort_hits(self, order="canonical"):
        """Order the device hit buffer before copying it out: "canonical" = (spacer_id, gpos, strand),
        "best" = (spacer_id, mismatches, gpos, strand), the order of `bowtie --best`."""
        self._check(self._L.bc_sort_hits(self._ctx, {"canonical": 0, "best": 1}[order]))

    def hits(self):
        """Copy the result records of the last search to the host (unordered unless sort_hits() was called)."""
        st = self.stats()
        out = np.empty(st["hits"], dtype=HIT_DTYPE)
        if len(out):
            self._check(self._L.bc_copy_hits(self._ctx, out.ctypes.data, len(out)))
        return out

    def hits_into(self, host_ptr, cap_records):
        """Copy the records into caller-owned host memory (e.g. a pinned buffer)."""
        self._check(self._L.bc_copy_hits(self._ctx, host_ptr, int(cap_records)))

    def set_hit_sink(self, host_ptr, cap_records):
        """Stream the records of every following search() into caller-owned host memory while the
        search runs (pinned memory lets the copies overlap it); host_ptr=None removes the sink."""
        self._check(self._L.bc_set_hit_sink(self._ctx, host_ptr or None, int(cap_records) if host_ptr else 0))

    def set_slice_callback(self, fn):
        """fn(d_hits_base_pointer, begin, end) is called from inside search() whenever records
        [begin, end) of the device hit buffer are final; None removes it.  With a callback a hit
        buffer overflow is an error (code BC_ELIMIT, stats()['hits'] = needed capacity)."""
        if fn is None:
            self._slice_cb = SLICE_FN()   # NULL function pointer
        else:
            self._slice_cb = SLICE_FN(lambda user, base, begin, end: fn(base or 0, begin, end))
        self._check(self._L.bc_set_slice_callback(self._ctx, self._slice_cb, None))

    def peer_export(self, n_records):
        """Allocate a device buffer of n_records hit records and export it: (pointer, 64-byte CUDA IPC
        handle).  Other processes of the box open it with peer_open() and use a slice of it as their
        hit sink."""
        ptr = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        self._check(self._L.bc_peer_export(self._ctx, int(n_records), ctypes.byref(ptr), handle))
        return ptr.value, handle.raw

    def peer_open(self, handle):
        ptr = ctypes.c_void_p()
        self._check(self._L.bc_peer_open(self._ctx, bytes(handle), ctypes.byref(ptr)))
        return ptr.value

    def peer_close(self, ptr, owner):
        self._check(self._L.bc_peer_close(self._ctx, ptr, 1 if owner else 0))

    def hits_device(self):
        ptr, n = ctypes.c_void_p(), ctypes.c_uint64()
        self._check(self._L.bc_hits_device(self._ctx, ctypes.byref(ptr), ctypes.byref(n)))
        return ptr.value or 0, n.value

    def enumerate_guides(self, L, pam, direction="downstream", iupac=False, reference_range=False):
        """Distinct pure-ACGT L-mers next to a PAM match on either strand of the resident genome
        (design_guides.py:22-49).  Returns uint8 ASCII rows [n, L], sorted."""
        d = {"downstream": 0, "upstream": 1}[direction]
        flags = (BC_PAM_IUPAC if iupac else 0) | (4 if reference_range else 0)
        n = ctypes.c_uint64()
        self._check(self._L.bc_enumerate_guides(self._ctx, int(L), pam.upper().encode("ascii"), d, flags,
                                                ctypes.byref(n)))
        codes = np.empty(n.value, dtype=np.uint64)
        if n.value:
            self._check(self._L.bc_copy_guides(self._ctx, codes.ctypes.data, n.value))
        out = np.empty((n.value, L), dtype=np.uint8)
        letters = np.frombuffer(b"ACGT", dtype=np.uint8)
        for j in range(L):
            out[:, j] = letters[((codes >> np.uint64(2 * j)) & np.uint64(3)).astype(np.uint8)]
        if len(out):
            out = np.unique(np.ascontiguousarray(out).view(np.dtype((np.void, L))).ravel()).view(np.uint8).reshape(-1, L)
        return out

    def stats(self):
        st = BcStats()
        self._check(self._L.bc_get_stats(self._ctx, ctypes.byref(st)))
        return st.as_dict()


def canonical_sort(hits):
    """(spacer_id, gpos, strand) order - the order every parity comparison uses."""
    order = np.lexsort((hits["meta"] & 1, hits["gpos"], hits["spacer_id"]))
    return hits[order]
