"""xsqueezeit_b200 -- B200-native (sm_100a) implementation of xSqueezeIt's genotype encode/decode
hot path, behind the reference's own interfaces.

Layers (top to bottom):
  * `Compressor` / `Accessor`  : host-side mirror of the reference's NewCompressor
    (include/gt_compressor_new.hpp:166-208) and Accessor (include/accessor.hpp:31-124) --
    same argument meaning, same error behaviour, same file bytes.
  * `Context`                  : thin ctypes binding of the C ABI in include/xsi_b200.h.
  * libxsi_b200.so             : hand-written CUDA kernels (csrc/*.cuh) + host container layer.

There is no CPU fallback: importing works anywhere, but every genotype operation needs the
CUDA library and a device, and raises `XsiError` otherwise.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("XSI_B200_SO") or os.path.join(_HERE, "libxsi_b200.so")  # the override is for A/B builds of the kernels

XSI_OK = 0
_ERRORS = {-1: "XSI_E_CUDA", -2: "XSI_E_ARG", -3: "XSI_E_ALLELE", -4: "XSI_E_PLOIDY", -5: "XSI_E_UNSUPPORTED",
           -6: "XSI_E_FORMAT", -7: "XSI_E_NOMEM", -8: "XSI_E_IO", -9: "XSI_E_ZSTD"}


class XsiError(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        super().__init__("%s (%d)%s" % (_ERRORS.get(code, "XSI_E_?"), code, ": " + msg if msg else ""))


class _EncodeDesc(ctypes.Structure):
    _fields_ = [("n_records", ctypes.c_uint64), ("n_samples", ctypes.c_uint32), ("block_len", ctypes.c_uint32),
                ("mac_threshold", ctypes.c_uint64), ("default_phasing", ctypes.c_int32),
                ("gt_elem_bytes", ctypes.c_int32), ("gt_on_device", ctypes.c_int32), ("wah_encode_missing", ctypes.c_int32),
                ("gt", ctypes.c_void_p), ("n_allele", ctypes.c_void_p), ("ploidy", ctypes.c_void_p)]


_lib = None


def lib():
    """Loads libxsi_b200.so (built in-tree by xsqueezeit_b200/build.py). Fails loudly when missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise XsiError(-1, "CUDA extension %s is missing: run `python -m xsqueezeit_b200.build` "
                           "(there is no CPU fallback)" % SO_PATH)
    L = ctypes.CDLL(SO_PATH)
    vp, u64, u32, i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int32
    P = ctypes.POINTER
    L.xsi_create.restype = i32
    L.xsi_create.argtypes = [i32, P(vp)]
    L.xsi_destroy.restype = None
    L.xsi_destroy.argtypes = [vp]
    L.xsi_last_error.restype = ctypes.c_char_p
    L.xsi_last_error.argtypes = [vp]
    L.xsi_version.restype = ctypes.c_char_p
    L.xsi_stream.restype = vp
    L.xsi_stream.argtypes = [vp]
    L.xsi_kernel_launches.restype = u64
    L.xsi_kernel_launches.argtypes = [vp]
    L.xsi_sync.restype = i32
    L.xsi_sync.argtypes = [vp]
    L.xsi_profile.restype = i32
    L.xsi_profile.argtypes = [vp, i32]
    L.xsi_profile_read.restype = ctypes.c_char_p
    L.xsi_profile_read.argtypes = [vp]
    L.xsi_encode_launch.restype = i32
    L.xsi_encode_launch.argtypes = [vp, P(_EncodeDesc)]
    L.xsi_encode_collect.restype = i32
    L.xsi_encode_collect.argtypes = [vp, P(u32), P(P(vp)), P(P(u64))]
    L.xsi_encode_async.restype = i32
    L.xsi_encode_async.argtypes = [vp, i32]
    L.xsi_encode_block_sizes.restype = i32
    L.xsi_encode_block_sizes.argtypes = [vp, P(u32), P(P(u64))]
    L.xsi_encode_max_ploidy.restype = i32
    L.xsi_encode_max_ploidy.argtypes = [vp]
    L.xsi_encode_line_counts.restype = i32
    L.xsi_encode_line_counts.argtypes = [vp, P(u64), P(u64)]
    L.xsi_decode_load_blocks.restype = i32
    L.xsi_decode_load_blocks.argtypes = [vp, u32, P(vp), P(u64), u64, i32]
    L.xsi_decode_records.restype = i32
    L.xsi_decode_records.argtypes = [vp, u64, vp, vp, vp, vp, u64, i32, vp, vp, u32]
    L.xsi_decode_records_i8.restype = i32
    L.xsi_decode_records_i8.argtypes = [vp, u64, vp, vp, vp, vp, u64, i32, vp, vp, u32]
    L.xsi_decode_records_subset.restype = i32
    L.xsi_decode_records_subset.argtypes = [vp, u64, vp, vp, vp, vp, u32, vp, u64, i32, vp, vp, u32]
    L.xsi_decode_allele_counts.restype = i32
    L.xsi_decode_allele_counts.argtypes = [vp, u64, vp, vp, vp, vp, u32]
    L.xsi_host_narrow_i32_i8.restype = i32
    L.xsi_host_narrow_i32_i8.argtypes = [vp, vp, u64]
    L.xsi_host_widen_i8_i32.restype = None
    L.xsi_host_widen_i8_i32.argtypes = [vp, u64, vp, u64, vp, u64]
    L.xsi_host_threads.restype = u32
    L.xsi_host_threads.argtypes = []
    L.xsi_transport_stats.restype = None
    L.xsi_transport_stats.argtypes = [vp, P(u64), P(u64)]
    L.xsi_writer_open.restype = i32
    L.xsi_writer_open.argtypes = [ctypes.c_char_p, u32, ctypes.c_char_p, u32, u64, i32, i32, i32, P(vp)]
    L.xsi_writer_add_blocks.restype = i32
    L.xsi_writer_add_blocks.argtypes = [vp, u32, P(vp), P(u64), u64, u64]
    L.xsi_writer_close.restype = i32
    L.xsi_writer_close.argtypes = [vp, i32]
    L.xsi_writer_finalize_sharded.restype = i32
    L.xsi_writer_finalize_sharded.argtypes = [ctypes.c_char_p, u32, ctypes.c_char_p, u32, u64, i32, i32, u32, P(u64), u64, u64, u64]
    L.xsi_host_alloc.restype = i32
    L.xsi_host_alloc.argtypes = [P(vp), u64]
    L.xsi_host_free.restype = None
    L.xsi_host_free.argtypes = [vp]
    L.xsi_device_alloc.restype = i32
    L.xsi_device_alloc.argtypes = [vp, P(vp), u64]
    L.xsi_device_free.restype = None
    L.xsi_device_free.argtypes = [vp, vp]
    L.xsi_encode_launch_strided.restype = i32
    L.xsi_encode_launch_strided.argtypes = [vp, vp, u64]
    L.xsi_decode_load_blocks_lazy.restype = i32
    L.xsi_decode_load_blocks_lazy.argtypes = [vp, u32, P(vp), P(u64), u64, i32, u32]
    L.xsi_decode_extend.restype = i32
    L.xsi_decode_extend.argtypes = [vp, u32, u32]
    L.xsi_decode_lines_ready.restype = i32
    L.xsi_decode_lines_ready.argtypes = [vp, u32, P(u32)]
    L.xsi_decode_dot_products.restype = i32
    L.xsi_decode_dot_products.argtypes = [vp, u64, vp, vp, vp, vp, i32, vp, u32]
    L.xsi_decode_block_info.restype = i32
    L.xsi_decode_block_info.argtypes = [vp, u32, P(u32), P(u32)]
    L.xsi_reader_open.restype = i32
    L.xsi_reader_open.argtypes = [ctypes.c_char_p, P(vp)]
    L.xsi_reader_close.restype = None
    L.xsi_reader_close.argtypes = [vp]
    L.xsi_reader_info.restype = i32
    L.xsi_reader_info.argtypes = [vp, P(u64), P(u64), P(u32), P(u32), P(u32), P(u32), P(u64), P(u64), P(i32), P(u64), P(i32)]
    L.xsi_reader_sample_name.restype = ctypes.c_char_p
    L.xsi_reader_sample_name.argtypes = [vp, u64]
    L.xsi_reader_gt_block.restype = i32
    L.xsi_reader_gt_block.argtypes = [vp, u32, P(vp), P(u64)]
    _lib = L
    return L


def _ptr(x):
    """numpy array -> address; int -> device/host address as is; None -> NULL."""
    if x is None:
        return None
    if isinstance(x, (int, np.integer)):
        return int(x)
    return x.ctypes.data


class Context:
    """One CUDA context of the library (a stream + scratch pools) on `device`."""

    def __init__(self, device=0):
        self._L = lib()
        h = ctypes.c_void_p()
        rc = self._L.xsi_create(int(device), ctypes.byref(h))
        if rc != XSI_OK:
            raise XsiError(rc, "no usable CUDA device %d (the hot path has no CPU fallback)" % device)
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self._L.xsi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != XSI_OK:
            raise XsiError(rc, self._L.xsi_last_error(self.h).decode())

    @property
    def stream(self):
        return self._L.xsi_stream(self.h)

    @property
    def kernel_launches(self):
        return int(self._L.xsi_kernel_launches(self.h))

    @property
    def transport_stats(self):
        """(bytes moved host->device, device->host) in the int8 transport encoding of host int32 rows."""
        a, b = ctypes.c_uint64(), ctypes.c_uint64()
        self._L.xsi_transport_stats(self.h, ctypes.byref(a), ctypes.byref(b))
        return int(a.value), int(b.value)

    def sync(self):
        self._check(self._L.xsi_sync(self.h))

    def profile(self, on=True):
        self._L.xsi_profile(self.h, 1 if on else 0)

    def profile_read(self):
        """{kernel: (launches, total_ms)} since the previous read (CUDA events on the ctx stream)."""
        out = {}
        for line in self._L.xsi_profile_read(self.h).decode().splitlines():
            n, c, ms = line.split()
            out[n] = (int(c), float(ms))
        return out

    # ---- encode --------------------------------------------------------------------------
    def encode_launch(self, gt, n_allele, n_samples, block_len, mac_threshold, default_phasing, ploidy=None,
                      gt_elem_bytes=4, gt_on_device=False, wah_encode_missing=False, row_stride=0):
        """gt: numpy array (host) or an integer device address when gt_on_device.  row_stride (elements, device rows only):
        rows r * row_stride apart instead of back to back (xsi_encode_launch_strided)."""
        self._na = np.ascontiguousarray(n_allele, dtype=np.uint32)
        self._pl = None if ploidy is None else np.ascontiguousarray(ploidy, dtype=np.uint8)
        self._gt_keep = gt
        d = _EncodeDesc()
        d.n_records = self._na.size
        d.n_samples = int(n_samples)
        d.block_len = int(block_len)
        d.mac_threshold = int(mac_threshold)
        d.default_phasing = int(default_phasing)
        d.gt_elem_bytes = int(gt_elem_bytes)
        d.gt_on_device = 1 if gt_on_device else 0
        d.wah_encode_missing = 1 if wah_encode_missing else 0
        d.gt = _ptr(gt)
        d.n_allele = self._na.ctypes.data
        d.ploidy = None if self._pl is None else self._pl.ctypes.data
        if row_stride:
            self._check(self._L.xsi_encode_launch_strided(self.h, ctypes.byref(d), int(row_stride)))
        else:
            self._check(self._L.xsi_encode_launch(self.h, ctypes.byref(d)))

    def device_alloc(self, nbytes):
        """xsi_device_alloc: a device address on this context's GPU (free with device_free)."""
        p = ctypes.c_void_p()
        self._check(self._L.xsi_device_alloc(self.h, ctypes.byref(p), int(nbytes)))
        return p.value

    def device_free(self, addr):
        self._L.xsi_device_free(self.h, ctypes.c_void_p(addr))

    def encode_async(self, on=True):
        """xsi_encode_async: launches of device rows return at once (a library thread encodes on its own stream) and
        encode_collect waits; decode calls on this context may run meanwhile."""
        self._check(self._L.xsi_encode_async(self.h, 1 if on else 0))

    def encode_collect(self):
        """Returns the byte-exact GT blocks (list of bytes) of the last launch."""
        n = ctypes.c_uint32()
        blocks = ctypes.POINTER(ctypes.c_void_p)()
        sizes = ctypes.POINTER(ctypes.c_uint64)()
        self._check(self._L.xsi_encode_collect(self.h, ctypes.byref(n), ctypes.byref(blocks), ctypes.byref(sizes)))
        return [ctypes.string_at(blocks[i], sizes[i]) for i in range(n.value)]

    def encode_collect_sizes(self):
        n = ctypes.c_uint32()
        blocks = ctypes.POINTER(ctypes.c_void_p)()
        sizes = ctypes.POINTER(ctypes.c_uint64)()
        self._check(self._L.xsi_encode_collect(self.h, ctypes.byref(n), ctypes.byref(blocks), ctypes.byref(sizes)))
        return [int(sizes[i]) for i in range(n.value)]

    @property
    def encode_max_ploidy(self):
        return int(self._L.xsi_encode_max_ploidy(self.h))

    def encode_line_counts(self):
        """(binary lines, PBWT+WAH lines) of the last collected batch."""
        a, b = ctypes.c_uint64(), ctypes.c_uint64()
        self._check(self._L.xsi_encode_line_counts(self.h, ctypes.byref(a), ctypes.byref(b)))
        return int(a.value), int(b.value)

    # ---- decode --------------------------------------------------------------------------
    def decode_load_blocks(self, blocks, num_samples, aet_bytes, lazy_lines=None):
        """blocks: list of bytes-like GT block payloads (or (address, size) tuples).  lazy_lines: stop the inverse-PBWT
        chain after that many binary lines per block (xsi_decode_load_blocks_lazy); decode calls continue it on demand."""
        n = len(blocks)
        ptrs = (ctypes.c_void_p * n)()
        sizes = (ctypes.c_uint64 * n)()
        self._blk_keep = []
        for i, b in enumerate(blocks):
            if isinstance(b, tuple):
                ptrs[i], sizes[i] = b
            else:
                a = np.frombuffer(b, dtype=np.uint8)
                self._blk_keep.append(a)
                ptrs[i], sizes[i] = a.ctypes.data, a.size
        # the loaded set belongs to the context (a load replaces it): readers that cache "my block is resident" compare this
        self.load_generation = getattr(self, "load_generation", 0) + 1
        if lazy_lines is None:
            self._check(self._L.xsi_decode_load_blocks(self.h, n, ptrs, sizes, int(num_samples), int(aet_bytes)))
        else:
            self._check(self._L.xsi_decode_load_blocks_lazy(self.h, n, ptrs, sizes, int(num_samples), int(aet_bytes), int(lazy_lines)))
        self._dec_hap = 2 * int(num_samples)

    def decode_dot_products(self, block_index, line_offset, n_alleles, y):
        """xsi_decode_dot_products: float64 [n, max(n_alleles)-1], column a-1 = sum of y[sample] over the carriers of ALT a."""
        bi = np.ascontiguousarray(block_index, dtype=np.uint32)
        lo = np.ascontiguousarray(line_offset, dtype=np.uint32)
        na = np.ascontiguousarray(n_alleles, dtype=np.uint32)
        yy = np.ascontiguousarray(y, dtype=np.float64)
        stride = max(1, int(na.max()) - 1) if na.size else 1
        out = np.zeros((bi.size, stride), dtype=np.float64)
        self._check(self._L.xsi_decode_dot_products(self.h, bi.size, bi.ctypes.data, lo.ctypes.data, na.ctypes.data, yy.ctypes.data, 0,
                                                    out.ctypes.data, stride))
        return out

    def decode_lines_ready(self, block_index):
        v = ctypes.c_uint32()
        self._check(self._L.xsi_decode_lines_ready(self.h, int(block_index), ctypes.byref(v)))
        return v.value

    def decode_extend(self, block_index, line_end):
        self._check(self._L.xsi_decode_extend(self.h, int(block_index), int(line_end)))

    def decode_records(self, block_index, line_offset, n_alleles, out=None, out_stride=None, out_on_device=False,
                       want_counts=False, elem_bytes=4):
        """elem_bytes=4: int32 rows (bcf_get_genotypes values); 1: raw BCF int8 FORMAT/GT rows."""
        bi = np.ascontiguousarray(block_index, dtype=np.uint32)
        lo = np.ascontiguousarray(line_offset, dtype=np.uint32)
        na = np.ascontiguousarray(n_alleles, dtype=np.uint32)
        n = bi.size
        stride = int(out_stride or self._dec_hap)
        if out is None:
            out = np.empty((n, stride), dtype=np.int32 if elem_bytes == 4 else np.int8)
        filled = np.zeros(n, dtype=np.uint32)
        counts = None
        cs = 0
        if want_counts:
            cs = int(na.max()) if n else 2
            counts = np.zeros((n, cs), dtype=np.uint64)
        fn = self._L.xsi_decode_records if elem_bytes == 4 else self._L.xsi_decode_records_i8
        self._check(fn(self.h, n, bi.ctypes.data, lo.ctypes.data, na.ctypes.data, _ptr(out), stride,
           1 if out_on_device else 0, filled.ctypes.data, None if counts is None else counts.ctypes.data, cs))
        return out, filled, counts


    def decode_records_subset(self, block_index, line_offset, n_alleles, samples_to_use, want_ac=True, out_device=None):
        """Rows of the selected samples only, in the order given (the extractor's -s/-S, gt_decompressor_new.hpp:208-238),
        their lengths, and the selected carriers per ALT allele (ac_s).  out_device: a device address that receives the rows
        (2 * len(samples_to_use) int32 apart) instead of a host array."""
        bi = np.ascontiguousarray(block_index, dtype=np.uint32)
        lo = np.ascontiguousarray(line_offset, dtype=np.uint32)
        na = np.ascontiguousarray(n_alleles, dtype=np.uint32)
        sel = np.ascontiguousarray(samples_to_use, dtype=np.uint32)
        n = bi.size
        stride = 2 * sel.size
        out = None if out_device else np.zeros((n, stride), dtype=np.int32)
        filled = np.zeros(n, dtype=np.uint32)
        acs = max(1, int(na.max()) - 1) if n else 1
        ac = np.zeros((n, acs), dtype=np.uint32) if want_ac else None
        self._check(self._L.xsi_decode_records_subset(self.h, n, bi.ctypes.data, lo.ctypes.data, na.ctypes.data, sel.ctypes.data,
                                                      sel.size, out_device if out_device else out.ctypes.data, stride,
                                                      1 if out_device else 0, filled.ctypes.data,
                                                      None if ac is None else ac.ctypes.data, acs))
        return out, filled, ac

    def decode_allele_counts(self, block_index, line_offset, n_alleles):
        """Counts only (AccessorInternals::fill_allele_counts); returns uint64 [n, max(n_alleles)]."""
        bi = np.ascontiguousarray(block_index, dtype=np.uint32)
        lo = np.ascontiguousarray(line_offset, dtype=np.uint32)
        na = np.ascontiguousarray(n_alleles, dtype=np.uint32)
        n = bi.size
        cs = int(na.max()) if n else 2
        counts = np.zeros((n, cs), dtype=np.uint64)
        self._check(self._L.xsi_decode_allele_counts(self.h, n, bi.ctypes.data, lo.ctypes.data, na.ctypes.data,
                                                     counts.ctypes.data, cs))
        return counts


# ---- file-level parameters (host logic, not the hot path) --------------------------------------
def seek_default_phased(rows):
    """Majority phase bit of the 2nd allele over the first 3 records; 0 as soon as one of them is
    haploid; ties -> phased (reference xcf.cpp:811-836).  rows: iterable of (gt_row, ploidy)."""
    c0 = c1 = 0
    for k, (row, pl) in enumerate(rows):
        if k >= 3:
            break
        if pl == 1:
            return 0
        ph = np.asarray(row)[1::pl] & 1
        c1 += int(ph.sum())
        c0 += int(ph.size - ph.sum())
    return 0 if c0 > c1 else 1


def mac_threshold(n_samples, first_record_ploidy, maf):
    """(size_t)((double)N_HAPS * MAF) with N_HAPS from the first record's ploidy
    (reference gt_compressor_new.hpp:88,98-99)."""
    return int(float(n_samples * first_record_ploidy) * float(maf))


class Compressor:
    """Mirror of the reference NewCompressor / GtCompressorStream / XsiFactoryExt driver
    (include/gt_compressor_new.hpp:84-142,166-208; include/xsi_factory.hpp:513-606), fed with
    in-memory genotype rows instead of htslib records.  Defaults follow include/xsqueezeit.hpp:101-113."""

    def __init__(self, ctx=None, maf=0.001, reset_sort_block_length=8192, zstd_compression_on=False,
                 zstd_compression_level=7, blocks_per_batch=8, wah_encode_missing=False):
        self.ctx = ctx or Context(0)
        self.MAF = maf
        self.RESET_SORT_BLOCK_LENGTH = reset_sort_block_length
        self.zstd_compression_on = zstd_compression_on
        self.zstd_compression_level = zstd_compression_level
        self.blocks_per_batch = blocks_per_batch
        self.wah_encode_missing = wah_encode_missing  # --wah-encode-missing (xsqueezeit.hpp:58)

    def set_maf(self, maf):
        self.MAF = maf

    def set_reset_sort_block_length(self, n):
        self.RESET_SORT_BLOCK_LENGTH = n

    def set_zstd_compression_on(self, on):
        self.zstd_compression_on = on

    def set_zstd_compression_level(self, level):
        self.zstd_compression_level = level

    def compress_to_file(self, filename, gt, ngt, n_allele, n_samples, sample_names=None, gt_elem_bytes=4):
        """gt: flat numpy array of rows back to back (int32, or int8 when gt_elem_bytes == 1);
        ngt[r] = entries of row r (n_samples * ploidy); n_allele[r] = bcf1_t::n_allele."""
        ngt = np.asarray(ngt, dtype=np.int64)
        n_allele = np.ascontiguousarray(n_allele, dtype=np.uint32)
        R = ngt.size
        if n_samples <= 0 or R == 0:
            raise XsiError(-2, "no samples / no records")
        ploidy = (ngt // n_samples).astype(np.int64)
        if (ploidy > 2).any():
            raise XsiError(-4, "Ploidy higher than 2 is not yet supported")
        off = np.zeros(R + 1, dtype=np.int64)
        np.cumsum(ngt, out=off[1:])
        gt = np.ascontiguousarray(gt)

        def rows3():
            for r in range(min(3, R)):
                row = gt[off[r]:off[r + 1]]
                if gt_elem_bytes == 1:
                    row = row.astype(np.int32)
                yield row, int(ploidy[r])

        default_phased = seek_default_phased(rows3())
        thr = mac_threshold(n_samples, int(ploidy[0]), self.MAF)
        L = self.ctx._L
        blob = None if sample_names is None else b"".join(s.encode() + b"\0" for s in sample_names)
        w = ctypes.c_void_p()
        rc = L.xsi_writer_open(filename.encode(), n_samples, blob, self.RESET_SORT_BLOCK_LENGTH, thr, default_phased,
                               1 if self.zstd_compression_on else 0, self.zstd_compression_level, ctypes.byref(w))
        if rc != XSI_OK:
            raise XsiError(rc, "Failed to open file")
        max_ploidy = 0
        bl = self.RESET_SORT_BLOCK_LENGTH
        step = bl * self.blocks_per_batch
        try:
            for r0 in range(0, R, step):
                r1 = min(R, r0 + step)
                self.ctx.encode_launch(gt[off[r0]:off[r1]], n_allele[r0:r1], n_samples, bl, thr, default_phased,
                                       ploidy=ploidy[r0:r1].astype(np.uint8), gt_elem_bytes=gt_elem_bytes,
                                       wah_encode_missing=self.wah_encode_missing)
                n = ctypes.c_uint32()
                blocks = ctypes.POINTER(ctypes.c_void_p)()
                sizes = ctypes.POINTER(ctypes.c_uint64)()
                self.ctx._check(L.xsi_encode_collect(self.ctx.h, ctypes.byref(n), ctypes.byref(blocks), ctypes.byref(sizes)))
                rc = L.xsi_writer_add_blocks(w, n.value, blocks, sizes, r1 - r0, int((n_allele[r0:r1].astype(np.int64) - 1).sum()))
                if rc != XSI_OK:
                    raise XsiError(rc, "block write failed")
                max_ploidy = max(max_ploidy, self.ctx.encode_max_ploidy)
        except Exception:
            L.xsi_writer_close(w, max_ploidy)
            raise
        rc = L.xsi_writer_close(w, max_ploidy)
        if rc != XSI_OK:
            raise XsiError(rc, "finalize failed")
        return dict(default_phased=default_phased, mac_threshold=thr, max_ploidy=max_ploidy)


class Accessor:
    """Mirror of the reference Accessor (include/accessor.hpp:31-124, accessor.cpp:26-82):
    fill_genotype_array(n_alleles, position) with position = BM = block<<15 | binary line offset.
    A whole block is decoded on the GPU when first touched; `prefetch` decodes many records in
    one launch (what `xsqueezeit -x` and c_xcf_get_genotypes loops want)."""

    BM_BLOCK_BITS = 15

    def __init__(self, filename, ctx=None, blocks_resident=1):
        self.ctx = ctx or Context(0)
        self._L = self.ctx._L
        r = ctypes.c_void_p()
        rc = self._L.xsi_reader_open(filename.encode(), ctypes.byref(r))
        if rc != XSI_OK:
            raise XsiError(rc, "Failed to open / bad magic / bad version: " + filename)
        self.r = r
        ns, hs, pl, aet, nb, bl = (ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint32(), ctypes.c_uint32(),
                                   ctypes.c_uint32(), ctypes.c_uint32())
        ent, nv, z, rt, dp = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_int32(), ctypes.c_uint64(), ctypes.c_int32()
        self._L.xsi_reader_info(r, ctypes.byref(ns), ctypes.byref(hs), ctypes.byref(pl), ctypes.byref(aet), ctypes.byref(nb),
                                ctypes.byref(bl), ctypes.byref(ent), ctypes.byref(nv), ctypes.byref(z), ctypes.byref(rt),
                                ctypes.byref(dp))
        self.num_samples, self.hap_samples, self.ploidy = ns.value, hs.value, pl.value
        self.aet_bytes, self.n_blocks, self.block_len = aet.value, nb.value, bl.value
        self.xcf_entries, self.num_variants, self.zstd = ent.value, nv.value, bool(z.value)
        self.rare_threshold, self.default_phased = rt.value, dp.value
        self._loaded = None  # (first block, count, load generation of the context)
        self._counts = None
        self.blocks_resident = max(1, int(blocks_resident))  # blocks brought to the device per miss

    def get_sample_list(self):
        out = []
        for i in range(self.hap_samples // self.ploidy):
            s = self._L.xsi_reader_sample_name(self.r, i)
            if s is None:
                break
            out.append(s.decode())
        return out

    def get_number_of_samples(self):
        return len(self.get_sample_list())

    def _load(self, b0, nb=None):
        """Makes block b0 resident (with the blocks_resident - 1 that follow).  The context may be shared: another Accessor or
        a direct decode_load_blocks call replaces the loaded set, which the generation check notices."""
        if nb is None:
            nb = min(self.blocks_resident, self.n_blocks - b0)
        if (self._loaded is not None and self._loaded[2] == getattr(self.ctx, "load_generation", 0)
                and self._loaded[0] <= b0 < self._loaded[0] + self._loaded[1]):
            return
        if b0 < 0 or b0 >= self.n_blocks:
            raise XsiError(-2, "block %d out of range (%d blocks)" % (b0, self.n_blocks))
        blocks = []
        for b in range(b0, b0 + nb):
            p, s = ctypes.c_void_p(), ctypes.c_uint64()
            rc = self._L.xsi_reader_gt_block(self.r, b, ctypes.byref(p), ctypes.byref(s))
            if rc != XSI_OK:
                raise XsiError(rc, "block error")
            blocks.append((p.value, s.value))
        self.ctx.decode_load_blocks(blocks, self.num_samples, self.aet_bytes)
        self._loaded = (b0, nb, self.ctx.load_generation)

    def split_bm(self, position):
        position = int(position) & 0xFFFFFFFF
        return position >> self.BM_BLOCK_BITS, position & ((1 << self.BM_BLOCK_BITS) - 1)

    def fill_genotype_array(self, n_alleles, position, gt_arr=None):
        """Returns (gt_arr, n_filled) like AccessorInternals::fill_genotype_array."""
        b, off = self.split_bm(position)
        self._load(b)
        stride = max(int(self.hap_samples), 2 * int(self.num_samples))
        out = np.empty((1, stride), dtype=np.int32)
        out, filled, counts = self.ctx.decode_records([b - self._loaded[0]], [off], [n_alleles], out=out,
                                                      out_stride=stride, want_counts=True)
        self._counts = counts[0, :n_alleles].copy()
        n = int(filled[0])
        if gt_arr is not None:
            gt_arr[:n] = out[0, :n]
            return gt_arr, n
        return out[0], n

    def get_allele_counts(self):
        return self._counts

    def fill_allele_counts(self, n_alleles, position):
        """AccessorInternals::fill_allele_counts: counts without the genotype row; read with get_allele_counts()."""
        b, off = self.split_bm(position)
        self._load(b)
        self._counts = self.ctx.decode_allele_counts([b - self._loaded[0]], [off], [n_alleles])[0, :n_alleles].copy()

    def fill_allele_counts_batch(self, n_alleles, positions):
        """Batch form of fill_allele_counts: uint64 [n, max(n_alleles)]."""
        positions = np.asarray(positions, dtype=np.uint64)
        n_alleles = np.asarray(n_alleles, dtype=np.uint32)
        blk = ((positions & np.uint64(0xFFFFFFFF)) >> np.uint64(self.BM_BLOCK_BITS)).astype(np.int64)
        off = (positions & np.uint64((1 << self.BM_BLOCK_BITS) - 1)).astype(np.uint32)
        n = positions.size
        counts = np.zeros((n, int(n_alleles.max()) if n else 2), dtype=np.uint64)
        i = 0
        while i < n:
            self._load(int(blk[i]))
            lo, hi = self._loaded[0], self._loaded[0] + self._loaded[1]
            j = i
            while j < n and lo <= blk[j] < hi:
                j += 1
            c = self.ctx.decode_allele_counts((blk[i:j] - lo).astype(np.uint32), off[i:j], n_alleles[i:j])
            counts[i:j, :c.shape[1]] = c
            i = j
        return counts

    def fill_genotype_arrays(self, n_alleles, positions, out=None, out_on_device=False, want_counts=False, elem_bytes=4):
        """Batch form: decodes every requested record; blocks are loaded in runs.
        elem_bytes=1 yields the records' raw BCF int8 FORMAT/GT rows instead of int32."""
        positions = np.asarray(positions, dtype=np.uint64)
        n_alleles = np.asarray(n_alleles, dtype=np.uint32)
        blk = ((positions & np.uint64(0xFFFFFFFF)) >> np.uint64(self.BM_BLOCK_BITS)).astype(np.int64)
        off = (positions & np.uint64((1 << self.BM_BLOCK_BITS) - 1)).astype(np.uint32)
        stride = max(int(self.hap_samples), 2 * int(self.num_samples))
        n = positions.size
        if out is None:
            out = np.empty((n, stride), dtype=np.int32 if elem_bytes == 4 else np.int8)
        filled = np.zeros(n, dtype=np.uint32)
        counts = np.zeros((n, int(n_alleles.max()) if n else 2), dtype=np.uint64) if want_counts else None
        i = 0
        while i < n:
            self._load(int(blk[i]))
            lo, hi = self._loaded[0], self._loaded[0] + self._loaded[1]
            j = i
            while j < n and lo <= blk[j] < hi:
                j += 1
            sub_out = out[i:j] if not out_on_device else int(out) + i * stride * elem_bytes
            o, f, c = self.ctx.decode_records((blk[i:j] - lo).astype(np.uint32), off[i:j], n_alleles[i:j], out=sub_out,
                                              out_stride=stride, out_on_device=out_on_device, want_counts=want_counts,
                                              elem_bytes=elem_bytes)
            filled[i:j] = f
            if want_counts:
                counts[i:j, :c.shape[1]] = c
            i = j
        return out, filled, counts

    def close(self):
        if getattr(self, "r", None):
            self._L.xsi_reader_close(self.r)
            self.r = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def bm_positions(n_allele, block_len):
    """BM of every record: block<<15 | binary-line offset (reference xcf.cpp:685-704)."""
    n_allele = np.asarray(n_allele, dtype=np.int64)
    R = n_allele.size
    rec = np.arange(R, dtype=np.int64)
    blk = rec // block_len
    nalt = n_allele - 1
    cs = np.cumsum(nalt) - nalt
    base = cs[(blk * block_len).clip(max=R - 1)] if R else cs
    return ((blk << 15) | (cs - base)).astype(np.uint64)
