"""oracle/xsi_ref.py -- TEST INFRASTRUCTURE ONLY.

ctypes door onto oracle/_ref/libxsi_ref.so: the UNMODIFIED reference (XsiFactoryExt writer and
Accessor reader) compiled by oracle/Makefile from /root/reference, driven by oracle/ref_shim.cpp.
Only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline legs may import this.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libxsi_ref.so")
REF_CLI = os.path.join(_HERE, "_ref", "xsqueezeit_ref")

_lib = None


def available():
    return os.path.exists(REF_SO)


def open_lib(path):
    """The shim's C interface from any build of oracle/ref_shim.cpp: the CPU reference (oracle/_ref/libxsi_ref.so) or the same
    file compiled with the B200 adapters (bindings/_out/libxsi_shim_b200.so)."""
    L = ctypes.CDLL(path)
    _declare(L)
    return L


def lib():
    global _lib
    if _lib is None:
        _lib = open_lib(REF_SO)
    return _lib


def _declare(L):
    if True:
        L.xsi_ref_encode_file.restype = ctypes.c_int
        L.xsi_ref_encode_file.argtypes = [
            ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p]
        L.xsi_ref_set_wah_encode_missing.restype = None
        L.xsi_ref_set_wah_encode_missing.argtypes = [ctypes.c_int]
        L.xsi_ref_accessor_open.restype = ctypes.c_void_p
        L.xsi_ref_accessor_open.argtypes = [ctypes.c_char_p]
        L.xsi_ref_hap_samples.restype = ctypes.c_uint64
        L.xsi_ref_hap_samples.argtypes = [ctypes.c_void_p]
        L.xsi_ref_fill_genotype_array.restype = ctypes.c_uint64
        L.xsi_ref_fill_genotype_array.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64,
                                                  ctypes.c_uint64, ctypes.c_uint64]
        L.xsi_ref_fill_allele_counts.restype = ctypes.c_int
        L.xsi_ref_fill_allele_counts.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64]
        L.xsi_ref_allele_counts.restype = ctypes.c_uint64
        L.xsi_ref_allele_counts.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
        L.xsi_ref_accessor_close.restype = None
        L.xsi_ref_accessor_close.argtypes = [ctypes.c_void_p]
        if hasattr(L, "xsi_ref_internal_access"):
            L.xsi_ref_internal_access.restype = ctypes.c_int
            L.xsi_ref_internal_access.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64,
                                                  ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]


def encode_file(path, gt, rec_off, ngt, n_allele, n_samples, block_len, mac_threshold,
                default_phased, zstd=False, zstd_level=7, sample_names=None, wah_encode_missing=False, L=None):
    """Run the reference writer on in-memory rows. gt: int32 flat, rec_off: uint64 row starts."""
    gt = np.ascontiguousarray(gt, dtype=np.int32)
    rec_off = np.ascontiguousarray(rec_off, dtype=np.uint64)
    ngt = np.ascontiguousarray(ngt, dtype=np.int32)
    n_allele = np.ascontiguousarray(n_allele, dtype=np.int32)
    blob = None
    if sample_names is not None:
        blob = b"".join(s.encode() + b"\0" for s in sample_names)
    L = L or lib()
    L.xsi_ref_set_wah_encode_missing(1 if wah_encode_missing else 0)
    try:
        rc = L.xsi_ref_encode_file(path.encode(), gt.ctypes.data, rec_off.ctypes.data, ngt.ctypes.data,
                                       n_allele.ctypes.data, len(ngt), n_samples, block_len, mac_threshold,
                                       int(default_phased), int(zstd), zstd_level, 1, blob)
    finally:
        L.xsi_ref_set_wah_encode_missing(0)
    if rc != 0:
        raise RuntimeError("reference encode failed rc=%d" % rc)


class RefAccessor:
    """The reference Accessor (accessor.hpp:31-124) on an .xsi file."""

    def __init__(self, path, L=None):
        self.L = L or lib()
        self.h = self.L.xsi_ref_accessor_open(path.encode())
        if not self.h:
            raise RuntimeError("reference Accessor failed to open " + path)
        self.hap_samples = int(self.L.xsi_ref_hap_samples(self.h))

    def internal_access(self, n_alleles, position, a_bytes, nbytes=8):
        """(a[uint32], sparse flags, first nbytes of every line, default allele) of Accessor::get_internal_access."""
        a = np.zeros(self.hap_samples, dtype=np.uint32)
        sp = np.zeros(max(1, n_alleles - 1), dtype=np.uint8)
        by = np.zeros((max(1, n_alleles - 1), nbytes), dtype=np.uint8)
        da = ctypes.c_int32(0)
        n = self.L.xsi_ref_internal_access(self.h, n_alleles, position, a.ctypes.data, a.size, a_bytes, sp.ctypes.data, by.ctypes.data,
                                           nbytes, ctypes.byref(da))
        if n < 0:
            raise RuntimeError("get_internal_access threw (%d)" % n)
        return a, sp[:n].copy(), by[:n].copy(), da.value

    def fill_genotype_array(self, n_alleles, position, out=None):
        if out is None:
            out = np.empty(self.hap_samples, dtype=np.int32)
        n = self.L.xsi_ref_fill_genotype_array(self.h, out.ctypes.data, out.size, n_alleles, position)
        if n == 2**64 - 1:
            raise RuntimeError("reference fill_genotype_array threw")
        return out, int(n)

    def fill_allele_counts(self, n_alleles, position):
        if self.L.xsi_ref_fill_allele_counts(self.h, n_alleles, position) != 0:
            raise RuntimeError("reference fill_allele_counts threw")
        return self.allele_counts()

    def allele_counts(self):
        buf = np.zeros(256, dtype=np.uint64)
        n = self.L.xsi_ref_allele_counts(self.h, buf.ctypes.data, buf.size)
        return buf[:n].copy()

    def close(self):
        if self.h:
            self.L.xsi_ref_accessor_close(self.h)
            self.h = None

    def __del__(self):
        self.close()


def load_gtdump(prefix):
    """Read oracle/gtdump.c output: returns (n_samples, n_allele[int32], ngt[int32], gt[int32 flat], names)."""
    meta = np.fromfile(prefix + ".meta", dtype=np.uint8)
    magic, n_samples = np.frombuffer(meta[:8], dtype=np.uint32)
    assert magic == 0x31445447
    n_records = int(np.frombuffer(meta[8:16], dtype=np.uint64)[0])
    m = np.frombuffer(meta[16:16 + 8 * n_records], dtype=np.uint32).reshape(n_records, 2)
    gt = np.fromfile(prefix + ".gt", dtype=np.int32)
    names = [l.rstrip("\n") for l in open(prefix + ".samples")]
    return int(n_samples), m[:, 0].astype(np.int32), m[:, 1].astype(np.int32), gt, names
