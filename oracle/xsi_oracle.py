"""oracle/xsi_oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes wrapper around oracle/libxsi_oracle.so (the plain-C restatement in xsi_oracle.c).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module;
the product package xsqueezeit_b200 never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "libxsi_oracle.so")
_lib = None


def build():
    src = os.path.join(_HERE, "xsi_oracle.c")
    if (not os.path.exists(SO)) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", SO, src], cwd=_HERE)


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(SO)
        vp, u64, i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int
        L.xo_wah_encode_bits.restype = u64
        L.xo_wah_encode_bits.argtypes = [vp, u64, vp]
        L.xo_wah_decode_bits.restype = u64
        L.xo_wah_decode_bits.argtypes = [vp, u64, u64, vp, vp]
        L.xo_unordered_order.restype = None
        L.xo_unordered_order.argtypes = [vp, ctypes.c_uint32, vp]
        L.xo_default_phased.restype = i32
        L.xo_default_phased.argtypes = [vp, vp, vp, u64, u64]
        L.xo_mac_threshold.restype = u64
        L.xo_mac_threshold.argtypes = [u64, u64, ctypes.c_double]
        L.xo_encode.restype = i32
        L.xo_encode.argtypes = [vp, vp, vp, vp, u64, u64, u64, u64, i32, ctypes.c_char_p,
                                ctypes.POINTER(vp), ctypes.POINTER(u64)]
        L.xo_encode_opt.restype = i32
        L.xo_encode_opt.argtypes = [vp, vp, vp, vp, u64, u64, u64, u64, i32, ctypes.c_char_p, i32,
                                    ctypes.POINTER(vp), ctypes.POINTER(u64)]
        L.xo_free.restype = None
        L.xo_free.argtypes = [vp]
        L.xo_open.restype = vp
        L.xo_open.argtypes = [vp, u64]
        L.xo_close.restype = None
        L.xo_close.argtypes = [vp]
        L.xo_hap_samples.restype = u64
        L.xo_hap_samples.argtypes = [vp]
        L.xo_num_blocks.restype = u64
        L.xo_num_blocks.argtypes = [vp]
        L.xo_fill_genotype_array.restype = ctypes.c_int64
        L.xo_fill_genotype_array.argtypes = [vp, vp, u64, u64, u64]
        L.xo_fill_allele_counts.restype = ctypes.c_int64
        L.xo_fill_allele_counts.argtypes = [vp, u64, u64]
        L.xo_allele_counts.restype = u64
        L.xo_allele_counts.argtypes = [vp, vp, u64]
        _lib = L
    return _lib


def wah_encode_bits(bits):
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    out = np.zeros(bits.size // 15 + 4, dtype=np.uint16)
    n = lib().xo_wah_encode_bits(bits.ctypes.data, bits.size, out.ctypes.data)
    return out[:n].copy()


def wah_decode_bits(words, n_bits):
    words = np.ascontiguousarray(words, dtype=np.uint16)
    bits = np.zeros(n_bits + 32, dtype=np.uint8)
    ones = ctypes.c_uint64(0)
    used = lib().xo_wah_decode_bits(words.ctypes.data, words.size, n_bits, bits.ctypes.data, ctypes.byref(ones))
    return bits[:n_bits].copy(), int(used), int(ones.value)


def unordered_order(keys):
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    out = np.zeros_like(keys)
    lib().xo_unordered_order(keys.ctypes.data, keys.size, out.ctypes.data)
    return out


def row_offsets(ngt):
    ngt = np.asarray(ngt, dtype=np.int64)
    off = np.zeros(ngt.size, dtype=np.uint64)
    if ngt.size > 1:
        off[1:] = np.cumsum(ngt[:-1]).astype(np.uint64)
    return off


def default_phased(gt, rec_off, ngt, n_samples):
    gt = np.ascontiguousarray(gt, dtype=np.int32)
    rec_off = np.ascontiguousarray(rec_off, dtype=np.uint64)
    ngt = np.ascontiguousarray(ngt, dtype=np.int32)
    return int(lib().xo_default_phased(gt.ctypes.data, rec_off.ctypes.data, ngt.ctypes.data, ngt.size, n_samples))


def mac_threshold(n_samples, first_ploidy, maf):
    return int(lib().xo_mac_threshold(n_samples, first_ploidy, float(maf)))


def encode(gt, rec_off, ngt, n_allele, n_samples, block_len, mac_thr, default_phased_, sample_names=None,
           wah_encode_missing=False):
    """Returns the .xsi image (bytes) the reference writer would produce."""
    gt = np.ascontiguousarray(gt, dtype=np.int32)
    rec_off = np.ascontiguousarray(rec_off, dtype=np.uint64)
    ngt = np.ascontiguousarray(ngt, dtype=np.int32)
    n_allele = np.ascontiguousarray(n_allele, dtype=np.int32)
    blob = None
    if sample_names is not None:
        blob = b"".join(s.encode() + b"\0" for s in sample_names)
    out = ctypes.c_void_p()
    n = ctypes.c_uint64()
    rc = lib().xo_encode_opt(gt.ctypes.data, rec_off.ctypes.data, ngt.ctypes.data, n_allele.ctypes.data, ngt.size,
                             n_samples, block_len, mac_thr, int(default_phased_), blob, 1 if wah_encode_missing else 0,
                             ctypes.byref(out), ctypes.byref(n))
    if rc != 0:
        raise RuntimeError("oracle encode failed rc=%d" % rc)
    data = ctypes.string_at(out.value, n.value)
    lib().xo_free(out)
    return data


class Reader:
    """Oracle restatement of Accessor::fill_genotype_array over an in-memory .xsi image."""

    def __init__(self, image):
        self._img = np.frombuffer(bytes(image), dtype=np.uint8).copy()
        self.h = lib().xo_open(self._img.ctypes.data, self._img.size)
        if not self.h:
            raise RuntimeError("oracle: cannot open image")
        self.hap_samples = int(lib().xo_hap_samples(self.h))
        self.num_blocks = int(lib().xo_num_blocks(self.h))

    def fill_genotype_array(self, n_alleles, position, out=None):
        if out is None:
            out = np.empty(max(self.hap_samples, 1), dtype=np.int32)
        n = lib().xo_fill_genotype_array(self.h, out.ctypes.data, out.size, n_alleles, position)
        if n < 0:
            raise RuntimeError("oracle fill_genotype_array rc=%d" % n)
        return out, int(n)

    def fill_allele_counts(self, n_alleles, position):
        """Accessor::fill_allele_counts: counts only (allele_counts[0] ignores missing / end-of-vector entries)."""
        n = lib().xo_fill_allele_counts(self.h, n_alleles, position)
        if n < 0:
            raise RuntimeError("oracle fill_allele_counts rc=%d" % n)
        return self.allele_counts()

    def allele_counts(self):
        buf = np.zeros(256, dtype=np.uint64)
        n = lib().xo_allele_counts(self.h, buf.ctypes.data, buf.size)
        return buf[:n].copy()

    def close(self):
        if self.h:
            lib().xo_close(self.h)
            self.h = None

    def __del__(self):
        self.close()


def bm_positions(n_allele, block_len):
    """BM = block<<15 | binary line offset for every record (reference xcf.cpp:685-704)."""
    n_allele = np.asarray(n_allele, dtype=np.int64)
    pos = np.zeros(n_allele.size, dtype=np.uint64)
    off = 0
    for r in range(n_allele.size):
        if r % block_len == 0:
            off = 0
        pos[r] = ((r // block_len) << 15) | off
        off += int(n_allele[r]) - 1
    return pos


def select_samples(row, n, num_samples, samples_to_use, n_allele):
    """fill_selected_genotypes (reference include/gt_decompressor_new.hpp:208-238): the selected samples' entries of a
    decoded row of n values (ploidy n / num_samples, 1 or 2), in the order of samples_to_use, and ac_s[alt-1] = number of
    selected entries whose allele is alt.  Returns (selected row, ac_s)."""
    pl = n // num_samples
    if pl not in (1, 2):
        raise ValueError("PLOIDY ERROR")
    sel = np.asarray(samples_to_use, dtype=np.int64)
    idx = (sel[:, None] * pl + np.arange(pl)[None, :]).reshape(-1)
    picked = np.asarray(row[:n])[idx]
    allele = (picked >> 1) - 1
    ac = np.array([int((allele == a).sum()) for a in range(1, n_allele)], dtype=np.uint32)
    return picked, ac
