#!/usr/bin/env python
"""bench.py -- throughput of the genotype encode/decode hot path on B200, in Gigagenotypes/s.

A *step* is one pass of the hot path over one batch of synthetic HRC-shaped input
(32,488 samples = 64,976 haplotypes, PBWT blocks of 8,192 records, `--maf 0.001`):
encode the batch to byte-exact GT blocks, then decode every record of those blocks back to
int32 genotype rows.  A step moves G genotypes through the encoder and the same G through the
decoder; `value` = 2*G / t_step, and the two directions are also reported separately.

  value   inputs resident in HBM (int32 rows, the bcf_get_genotypes boundary), device pointers
          in / out through the C ABI (include/xsi_b200.h).
  e2e     the same calls with pinned HOST buffers: H2D of the int32 rows and D2H of the decoded
          int32 rows are inside the timed region.
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md "Measurement".

`--impl reference` times the unmodified reference (oracle/_ref/libxsi_ref.so, built from
/root/reference by oracle/Makefile) on all host cores on bounded samples of the same workload.

Multi-GPU: one process per GPU (torchrun), every rank encodes/decodes its own blocks (weak
scaling); the only exchange is the all-gather of per-block byte counts that builds the global
block offset table (reference xsi_factory.hpp:533,554,575).
"""
import argparse
import zlib
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np

HRC_SAMPLES = 32488
BLOCK_LEN = 8192
MAF = 0.001
METRIC = "Gigagenotypes/s compress & decompress (HRC-shape)"


# ------------------------------------------------------------------------------------------------
# synthetic workload: log-uniform allele frequency, LD by haplotype copying
# ------------------------------------------------------------------------------------------------
class HrcSynth:
    """Generates HRC-shaped int32 genotype rows chunk by chunk with torch (cuda or cpu).
    K founder haplotypes with Bernoulli(f) alleles; every sample haplotype copies a founder and
    switches to a pseudo-random founder with probability `switch` per site; per-genotype flips;
    sites whose expected carrier count is below 2*MAC threshold are placed on copiers of one founder."""

    def __init__(self, n_samples, seed, device, founders=256, switch=1e-3, flip=1e-4, maf=MAF, chrx=False):
        # chrx: SURVEY 8(d) S4 -- every 2nd sample haploid (second entry = end of vector), 5% of the records with 2-3 ALT
        # alleles, 0.5% missing alleles, 1% unphased genotypes
        self.chrx = chrx
        self.n_allele = []
        import torch
        self.t = torch
        self.S, self.H, self.K = n_samples, 2 * n_samples, founders
        self.dev = torch.device(device)
        self.g = torch.Generator(device=self.dev)
        self.g.manual_seed(seed)
        self.switch, self.flip = switch, flip
        self.rare_cut = 2 * int(float(self.H) * maf)
        self.nsw = torch.zeros(self.H, dtype=torch.int64, device=self.dev)   # switches so far, per haplotype
        self.hid = torch.arange(self.H, dtype=torch.int64, device=self.dev)
        self.seed = seed
        self.odd = (self.hid & 1).to(torch.int32)

    def chunk(self, out):
        """Fills out[rc, H] (int32, device tensor) with the next rc records."""
        t = self.t
        rc = out.shape[0]
        H, K = self.H, self.K
        u = t.rand(rc, device=self.dev, generator=self.g, dtype=t.float64)
        f = t.exp(u * (np.log(0.5) - np.log(1.0 / H)) + np.log(1.0 / H))
        fb = t.rand(rc, K, device=self.dev, generator=self.g) < f[:, None].to(t.float32)
        sw = t.rand(rc, H, device=self.dev, generator=self.g) < self.switch
        cnt = t.cumsum(sw.to(t.int32), dim=0).to(t.int64) + self.nsw[None, :]
        self.nsw = cnt[-1].clone()
        founder = ((self.hid[None, :] * 2654435761 + cnt * 40503 + self.seed * 97) >> 7) % K
        allele = t.gather(fb, 1, founder)
        allele ^= t.rand(rc, H, device=self.dev, generator=self.g) < self.flip
        target = t.round(f * H)
        rare = target < self.rare_cut
        if bool(rare.any()):
            kstar = t.randint(0, K, (rc, 1), device=self.dev, generator=self.g)
            p = (target * K / H).clamp(max=1.0).to(t.float32)
            carriers = (founder == kstar) & (t.rand(rc, H, device=self.dev, generator=self.g) < p[:, None])
            allele = t.where(rare[:, None], carriers, allele)
        # htslib encoding: (allele+1)<<1 | phased, phase bit only on the 2nd allele of a sample ("0|1")
        if not self.chrx:
            self.n_allele.append(np.full(rc, 2, np.uint32))
            out.copy_(((allele.to(t.int32) + 1) << 1) | self.odd[None, :])
            return
        a = allele.to(t.int32)
        nalt = t.where(t.rand(rc, device=self.dev, generator=self.g) < 0.05,
                       t.randint(2, 4, (rc,), device=self.dev, generator=self.g), t.ones(rc, dtype=t.int64, device=self.dev))
        alt = (t.rand(rc, H, device=self.dev, generator=self.g) * nalt[:, None].to(t.float32)).to(t.int32).clamp_(max=2) + 1
        alt = t.minimum(alt, nalt[:, None].to(t.int32))
        a = t.where(a > 0, alt, a)                                             # carriers get one of the ALT alleles
        phase = self.odd[None, :] & ~(t.rand(rc, H, device=self.dev, generator=self.g) < 0.01).to(t.int32)  # 1% unphased
        val = ((a + 1) << 1) | phase
        val = t.where(t.rand(rc, H, device=self.dev, generator=self.g) < 0.005, phase, val)                   # missing allele
        male = ((self.hid >> 1) & 1).bool() & (self.hid & 1).bool()               # 2nd entry of every 2nd sample
        val = t.where(male[None, :], t.full_like(val, -2147483647), val)          # bcf_int32_vector_end
        self.n_allele.append((nalt + 1).to(t.int32).cpu().numpy().astype(np.uint32))
        out.copy_(val)

    def fill(self, out, chunk=256):
        for r0 in range(0, out.shape[0], chunk):
            self.chunk(out[r0:r0 + chunk])


# ------------------------------------------------------------------------------------------------
# clocks sampler (NVML), during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake", 0x2: "app_clocks", 0x10: "sync_boost", 0x100: "display_clocks"}

    def __init__(self, index):
        self.samples, self.reasons, self.max = [], set(), None
        self._stop = threading.Event()
        self._th = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv:
            self._th = threading.Thread(target=self._run, daemon=True)
            self._th.start()

    def stop(self):
        self._stop.set()
        if self._th:
            self._th.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# reference arm: the unmodified reference on host cores (worker processes, one per core)
# ------------------------------------------------------------------------------------------------
def ref_worker(args):
    """Child process: builds one sample of the workload on the CPU, then on every 'go' line runs the
    reference writer (XsiFactoryExt) and the reference Accessor over it and reports the two times."""
    import torch
    torch.set_num_threads(1)
    import xsi_ref
    S, R = args.samples, args.ref_records
    # every worker maps the SAME rows (one full PBWT block by default), generated once by the parent into /dev/shm:
    # identical, deterministic work per core without a private 2 GB copy each
    g = np.memmap(args.ref_data, dtype=np.int32, mode="r", shape=(R * 2 * S,))
    ngt = np.full(R, 2 * S, np.int32)
    nal = np.full(R, 2, np.int32)
    off = (np.arange(R, dtype=np.uint64) * np.uint64(2 * S))
    thr = int(float(2 * S) * MAF)
    tmpdir = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    path = os.path.join(tmpdir, "xsi_ref_bench_%d_%d.xsi" % (os.getpid(), args.ref_worker))
    out = np.empty(2 * S, np.int32)
    rr = np.arange(R, dtype=np.uint64)
    pos = ((rr // np.uint64(BLOCK_LEN)) << np.uint64(15)) | (rr % np.uint64(BLOCK_LEN))  # bi-allelic: BM = block<<15 | line
    print("ready", flush=True)
    for line in sys.stdin:
        if line.strip() != "go":
            break
        t0 = time.perf_counter()
        xsi_ref.encode_file(path, g, off, ngt, nal, S, BLOCK_LEN, thr, 1)
        t1 = time.perf_counter()
        acc = xsi_ref.RefAccessor(path)
        chk = 0
        for r in range(R):
            _, n = acc.fill_genotype_array(2, int(pos[r]), out)
            chk += n
        acc.close()
        t2 = time.perf_counter()
        ok = bool(chk == R * 2 * S and np.array_equal(out, np.asarray(g[(R - 1) * 2 * S:])))
        print(json.dumps({"enc_s": t1 - t0, "dec_s": t2 - t1, "ok": ok, "xsi_bytes": os.path.getsize(path)}), flush=True)
    try:
        os.unlink(path)
    except OSError:
        pass


def ref_sample_file(samples, records):
    """One sample of the workload (HrcSynth, the generator of the GPU arm, seed 9000) as raw int32 rows in /dev/shm."""
    import torch
    tmpdir = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    path = os.path.join(tmpdir, "xsi_ref_rows_%d.i32" % os.getpid())
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    gen = HrcSynth(samples, 9000, dev)
    mm = np.memmap(path, dtype=np.int32, mode="w+", shape=(records, 2 * samples))
    step = 512
    buf = torch.empty((step, 2 * samples), dtype=torch.int32, device=dev)
    for r0 in range(0, records, step):
        n = min(step, records - r0)
        gen.fill(buf[:n], chunk=256 if dev == "cuda" else 64)
        mm[r0:r0 + n] = buf[:n].cpu().numpy()
    mm.flush()
    del mm
    return path


class RefPool:
    def __init__(self, samples, records, workers):
        self.samples, self.records = samples, records
        self.procs = []
        self.data = ref_sample_file(samples, records)
        for w in range(workers):
            self.procs.append(subprocess.Popen(
                [sys.executable, os.path.abspath(__file__), "--ref-worker", str(w), "--samples", str(samples),
                 "--ref-records", str(records), "--ref-data", self.data], stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, bufsize=1))
        for p in self.procs:
            line = p.stdout.readline()
            if line.strip() != "ready":
                raise RuntimeError("reference worker failed to start: %r" % line)

    def step(self):
        """All workers run encode+decode of their sample concurrently. Returns (wall_s, enc_s max, dec_s max, ok)."""
        t0 = time.perf_counter()
        for p in self.procs:
            p.stdin.write("go\n")
            p.stdin.flush()
        res = [json.loads(p.stdout.readline()) for p in self.procs]
        wall = time.perf_counter() - t0
        return wall, max(r["enc_s"] for r in res), max(r["dec_s"] for r in res), all(r["ok"] for r in res), res[0]["xsi_bytes"]

    def close(self):
        for p in self.procs:
            try:
                p.stdin.write("quit\n")
                p.stdin.flush()
                p.stdin.close()
            except Exception:
                pass
        for p in self.procs:
            p.wait(timeout=60)
        try:
            os.unlink(self.data)
        except OSError:
            pass


def getenv_flag(name):
    return os.environ.get(name, "") not in ("", "0")


def bind_to_gpu_numa_node(index):
    """Restricts this process (and the library's worker pool, created later) to the cores of the NUMA node the GPU hangs off, so
    that the pinned rings and the int8 <-> int32 conversion of a rank stay on the memory of that node.  Returns a short
    description, or None when the topology is not exposed (single node, container without sysfs).  Round 1 measured the
    host-buffer leg at 67 GB/s each way for 8 ranks together: one node's memory serving every rank."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = "/sys/bus/pci/devices/%s:%s/numa_node" % (dom[-4:].lower(), rest.lower())
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed or allowed == set(os.sched_getaffinity(0)):
            return None
        os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cores": len(allowed)}
    except Exception:
        return None


def usable_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, steps, warmup, label_impl=True):
    """Times the reference on the host: every worker owns `ref_records` HRC-shaped records."""
    import xsi_ref
    if not xsi_ref.available():
        return None
    import psutil
    S, R = args.samples, args.ref_records
    per_worker = 900e6  # the rows are shared (one mapping); a worker holds the reference's block state and one output row
    workers = int(max(1, min(usable_cores(), (psutil.virtual_memory().available * 0.6 - R * 2 * S * 4) // per_worker)))
    if args.ref_workers:
        workers = args.ref_workers
    pool = RefPool(S, R, workers)
    try:
        for _ in range(warmup):
            pool.step()
        walls, encs, decs, oks = [], [], [], []
        for _ in range(steps):
            w, e, d, ok, xb = pool.step()
            walls.append(w); encs.append(e); decs.append(d); oks.append(ok)
    finally:
        pool.close()
    G = workers * R * 2 * S
    t = float(np.sum(walls))
    return {"value": 2 * G * steps / t / 1e9, "compress": G * steps / float(np.sum(encs)) / 1e9,
            "decompress": G * steps / float(np.sum(decs)) / 1e9, "ms_per_step": t / steps * 1e3, "cores": workers,
            "ok": all(oks), "sample": "%d worker processes (one per host core) x %d HRC-shaped records (%d haplotypes, %s of %d records each; "
            "every worker maps the same rows), unmodified reference XsiFactoryExt encode to /dev/shm + Accessor decode of every record"
            % (workers, R, 2 * S, "%d full PBWT block(s)" % (R // BLOCK_LEN) if R % BLOCK_LEN == 0 else "a partial PBWT block", BLOCK_LEN),
            "genotypes_per_step": G}


# ------------------------------------------------------------------------------------------------
# BCF legs: one synthetic BCF written once with htslib, fed to both implementations
# ------------------------------------------------------------------------------------------------
BIND = os.path.join(ROOT, "bindings", "_out")
REFBIN = os.path.join(ROOT, "oracle", "_ref")


def _run_timed(argv, env=None):
    t0 = time.perf_counter()
    p = subprocess.run(argv, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    dt = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError("%s failed (%d): %s" % (" ".join(argv), p.returncode, p.stderr.decode()[-400:]))
    return dt, p.stdout.decode()


def _sha(path):
    import hashlib
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 24), b""):
            h.update(chunk)
    return h.hexdigest()


def bcf_legs(samples, records, which, threads):
    """BCF file -> .xsi + _var.bcf -> rows / BCF, whole processes timed by wall clock (start-up, CUDA init, file I/O on
    /dev/shm all inside).  which: 'b200', 'reference' or 'both'.  The reference programs are the unmodified reference
    (oracle/_ref/xsqueezeit_ref; bindings/_out/capi_decode_ref = its c_api.h loop); the B200 programs are
    bindings/_out/xsi_b200_bcf (ingest / egress around the C ABI), xsqueezeit_b200 (the reference CLI with the two
    adapters) and capi_decode_b200 (c_api.h on the GPU Accessor)."""
    need = [os.path.join(BIND, "synth_bcf")]
    if which in ("b200", "both"):
        need += [os.path.join(BIND, x) for x in ("xsi_b200_bcf", "xsqueezeit_b200", "capi_decode_b200")]
    if which in ("reference", "both"):
        need += [os.path.join(REFBIN, "xsqueezeit_ref"), os.path.join(BIND, "capi_decode_ref")]
    missing = [x for x in need if not os.path.exists(x)]
    if missing:
        return {"unavailable": "not built: " + ", ".join(os.path.relpath(x, ROOT) for x in missing)}
    import shutil
    import tempfile
    G = records * 2 * samples
    tmp = tempfile.mkdtemp(prefix="xsi_bcf_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    out = {"workload": "HRC-shaped synthetic BCF: %d samples x %d records (%d PBWT blocks), bindings/synth_bcf.cpp seed 1002 "
                       "(log-uniform allele frequency, haplotype-copying LD), bgzf-compressed, in /dev/shm" % (samples, records, -(-records // BLOCK_LEN)),
           "genotypes": G, "unit": "Ggt/s", "timing": "wall clock of whole processes", "host_cores": usable_cores()}
    try:
        src = os.path.join(tmp, "in.bcf")
        _run_timed([os.path.join(BIND, "synth_bcf"), src, "hrc", str(samples), str(records), "1002", str(threads)])
        out["bcf_bytes"] = os.path.getsize(src)
        env_nosum = dict(os.environ, XSI_CAPI_NO_CHECKSUM="1")

        def ggts(t):
            return G / t / 1e9

        ref = None
        if which in ("reference", "both"):
            rx = os.path.join(tmp, "ref", "d.xsi")
            os.makedirs(os.path.dirname(rx))
            tc, _ = _run_timed([os.path.join(REFBIN, "xsqueezeit_ref"), "-c", "-f", src, "-o", rx])
            td, so = _run_timed([os.path.join(BIND, "capi_decode_ref"), rx + "_var.bcf"], env=env_nosum)
            tdl = float(so.split()[5])
            _, so = _run_timed([os.path.join(BIND, "capi_decode_ref"), rx + "_var.bcf"])
            ref_sum = so.split()[7]
            tx, _ = _run_timed([os.path.join(REFBIN, "xsqueezeit_ref"), "-x", "-f", rx, "-o", os.path.join(tmp, "ref", "o.bcf")])
            ref = {"compress": {"ggts": ggts(tc), "seconds": tc, "program": "xsqueezeit -c (as shipped: one encode thread + one _var.bcf thread)"},
                   "decode_c_api": {"ggts": ggts(td), "seconds": td, "loop_seconds": tdl, "program": "c_xcf_get_genotypes per record (c_api.h), 1 thread"},
                   "extract_bcf": {"ggts": ggts(tx), "seconds": tx, "program": "xsqueezeit -x to bgzf-compressed BCF, 1 thread"},
                   "value": 2 * G / (tc + td) / 1e9}
            out["reference"] = ref
        if which in ("b200", "both"):
            bx = os.path.join(tmp, "b200", "d.xsi")
            os.makedirs(os.path.dirname(bx))
            tc, so = _run_timed([os.path.join(BIND, "xsi_b200_bcf"), "compress", src, bx, "--threads", str(threads), "--batch-blocks", "1"])
            td, so2 = _run_timed([os.path.join(BIND, "capi_decode_b200"), bx + "_var.bcf"], env=env_nosum)
            tdl = float(so2.split()[5])
            td2, so3 = _run_timed([os.path.join(BIND, "capi_decode_b200"), bx + "_var.bcf"])
            tx, _ = _run_timed([os.path.join(BIND, "xsi_b200_bcf"), "extract", bx, os.path.join(tmp, "b200", "o.bcf"), "--threads", str(threads)])
            ax = os.path.join(tmp, "ada", "d.xsi")
            os.makedirs(os.path.dirname(ax))
            tac, _ = _run_timed([os.path.join(BIND, "xsqueezeit_b200"), "-c", "-f", src, "-o", ax])
            out["compress"] = {"ggts": ggts(tc), "seconds": tc, "program": "xsi_b200_bcf compress: one pass, %d BGZF threads, raw int8 FORMAT/GT rows, encode thread" % threads}
            tok = so2.split()
            out["decode_c_api"] = {"ggts": ggts(td), "seconds": td, "loop_seconds": tdl,
                                   "setup_seconds": float(tok[9]) if len(tok) > 11 else None, "teardown_seconds": float(tok[11]) if len(tok) > 11 else None,
                                   "seconds_second_run_with_checksum": td2,
                                   "program": "c_xcf_get_genotypes per record (c_api.h) on AccessorInternalsB200 (decode-ahead window), 1 thread"}
            out["extract_bcf"] = {"ggts": ggts(tx), "seconds": tx, "program": "xsi_b200_bcf extract: int8 rows spliced into the records, %d BGZF threads" % threads}
            out["compress_reference_cli_with_adapter"] = {"ggts": ggts(tac), "seconds": tac, "program": "xsqueezeit -c built with bindings/gt_block_b200.hpp"}
            out["value"] = 2 * G / (tc + td) / 1e9
            out["value_is"] = "2 * genotypes / (compress seconds + C API decode seconds)"
            ident = {"adapter_cli_xsi_equals_ingest_xsi": _sha(ax) == _sha(bx)}
            if ref is not None:
                ident.update({"xsi": _sha(bx) == _sha(rx), "var_bcf": _sha(bx + "_var.bcf") == _sha(rx + "_var.bcf"),
                              "extract_bcf": _sha(os.path.join(tmp, "b200", "o.bcf")) == _sha(os.path.join(tmp, "ref", "o.bcf")),
                              "c_api_checksum": so3.split()[7] == ref_sum})
            out["identical_to_reference"] = ident
            if ref is not None:
                out["speedup"] = {"compress": ref["compress"]["seconds"] / tc, "decode_c_api": ref["decode_c_api"]["seconds"] / td,
                                  "extract_bcf": ref["extract_bcf"]["seconds"] / tx}
        elif ref is not None:
            out["value"] = ref["value"]
    except Exception as ex:
        out["error"] = repr(ex)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


# ------------------------------------------------------------------------------------------------
# the other shapes of BASELINE.json (configs[1], [3], [4]) as short sub-runs of this script
# ------------------------------------------------------------------------------------------------
SHAPES = {
    # key: (BASELINE.json config, argv)
    "s2_1kgp3": ("configs[1]: 1KGP3-shaped, 5,008 haplotypes x 1.8 M bi-allelic records (220 PBWT blocks)",
                 ["--samples", "2504", "--blocks", "220"]),
    "s4_chrx": ("configs[3]: chrX-shaped, 2,504 samples (every 2nd haploid), 200 k records, 5% multi-allelic, 0.5% missing, 1% unphased (24 PBWT blocks)",
                ["--samples", "2504", "--blocks", "24", "--shape", "chrx"]),
    "s5_biobank": ("configs[4]: biobank-shaped, 1,000,000 haplotypes (uint32 indices, grid-cooperative PBWT); 2 of the 25 PBWT blocks per GPU and step",
                   ["--samples", "500000", "--blocks", "2"]),
}


def shape_subrun(key, device_index):
    desc, argv = SHAPES[key]
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    env["CUDA_VISIBLE_DEVICES"] = vis.split(",")[device_index] if vis else str(device_index)
    try:
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--sub", "--steps", "3", "--warmup", "2"] + argv,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=600)
        if p.returncode != 0:
            return {"config": desc, "error": p.stderr.decode()[-300:]}
        d = json.loads(p.stdout.decode().strip().splitlines()[-1])
    except Exception as ex:  # a shape that fails must not take the headline line with it
        return {"config": desc, "error": repr(ex)}
    return {"config": desc, "workload": d["config"]["workload"], "blocks_per_step": d["config"]["blocks_per_gpu_per_step"],
            "genotypes_per_step": d["config"]["genotypes_per_gpu_per_step"], "value": d["value"], "unit": "Ggt/s",
            "compress_ggts": d["compress_ggts"], "decompress_ggts": d["decompress_ggts"], "ms_per_step": d["ms_per_step"],
            "verified": d["verified"], "gpu_launches": d["gpu_launches"], "steps": d["steps"],
            "xsi_payload_bytes_per_step": d["stats"]["xsi_payload_bytes_per_step"],
            "kernels": {k: {"ms_per_step": v["ms_per_step"], "share": v["share"]} for k, v in d["kernels"].items() if v["share"] and v["share"] > 0.02},
            "roofline_kernels": [{"kernel": r["kernel"], "frac": r["frac"], "achieved": r["achieved"], "unit": r["unit"]} for r in d["roofline_kernels"]]}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def algorithmic_bytes(kernel, G, L_wah, L, WS, payload, elem=4):
    """Bytes a kernel must move per step (DESIGN.md 'Kernels'): G genotypes, L binary lines of WS words."""
    row = WS * 4
    return {
        "scan_rows": elem * G + L * row,                 # int32 rows in, one bit-row per binary line out
        "pbwt_permute": 2 * L_wah * row,              # bit-row in, permuted bit-row out (a[] stays in smem)
        "wah_encode_rows": L_wah * row + payload,
        "pack_wah": 2 * payload,
        "sparse_emit": (L - L_wah) * row,
        "wah_expand": payload + L_wah * row,
        "pbwt_unpermute": 2 * L_wah * row,
        "compose_simple": elem * G + L * row,
        "compose_records": elem * G + L * row,           # bit-rows / index lists in, int32 rows out
    }.get(kernel)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--samples", type=int, default=HRC_SAMPLES)
    ap.add_argument("--blocks", type=int, default=32, help="PBWT blocks per GPU per step (resident leg)")
    ap.add_argument("--e2e-blocks", type=int, default=8, help="PBWT blocks per GPU per step (host-buffer leg)")
    ap.add_argument("--e2e-workers", type=int, default=4, help="host threads (one xsi_ctx each) of the host-buffer leg")
    ap.add_argument("--block-len", type=int, default=BLOCK_LEN)
    ap.add_argument("--elem", type=int, default=4, choices=[1, 4],
                    help="bytes per genotype of the resident rows: 4 = int32 (the metric's boundary type), 1 = raw BCF int8")
    ap.add_argument("--ref-records", type=int, default=BLOCK_LEN, help="records per reference worker and step (default: one full PBWT block)")
    ap.add_argument("--ref-data", default="", help=argparse.SUPPRESS)
    ap.add_argument("--ref-workers", type=int, default=0)
    ap.add_argument("--ref-worker", type=int, default=-1, help=argparse.SUPPRESS)
    ap.add_argument("--resident-contexts", type=int, default=3,
                    help="also time the resident leg from this many host threads (one xsi_ctx each, blocks split between them)")
    ap.add_argument("--shape", default="hrc", choices=["hrc", "chrx"],
                    help="chrx: mixed ploidy, multi-allelic, missing and unphased genotypes (SURVEY 8(d) S4); resident one-context leg only")
    ap.add_argument("--bcf-records", type=int, default=4 * BLOCK_LEN, help="records of the synthetic BCF of the e2e_bcf legs (0: skip them)")
    ap.add_argument("--no-shapes", action="store_true", help="skip the other BASELINE.json shapes (1KGP3, chrX, biobank) of the default run")
    ap.add_argument("--sub", action="store_true", help=argparse.SUPPRESS)  # a shape sub-run: resident one-context leg only
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--profile-only", action="store_true", help="one resident step, no e2e / cpu legs (for ncu)")
    args = ap.parse_args()

    if args.ref_worker >= 0:
        return ref_worker(args)
    if args.sub:
        args.resident_contexts, args.no_e2e, args.no_cpu_baseline, args.bcf_records, args.no_shapes = 0, True, True, 0, True

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    steps, warmup = args.steps, max(args.warmup, 0)
    workload = "%s synthetic: %d haplotypes, PBWT blocks of %d records, maf %.3g" % (
        ("HRC-shaped" if args.samples == HRC_SAMPLES else "1KGP3-shaped" if args.samples == 2504 else "biobank-shaped" if args.samples >= 400000
         else "custom-width") if args.shape == "hrc" else "chrX-shaped (mixed ploidy, multi-allelic, missing, unphased)", 2 * args.samples, args.block_len, MAF)

    # the configuration both arms are run on (identical keys and values in both JSON lines; what a run measured on top
    # of it -- payload bytes, line counts, contexts -- is reported under "stats")
    G_cfg = args.blocks * args.block_len * 2 * args.samples
    config = {"workload": workload, "blocks_per_gpu_per_step": args.blocks, "records_per_gpu_per_step": args.blocks * args.block_len,
              "genotypes_per_gpu_per_step": G_cfg,
              "input": "%s rows resident in HBM (%.1f GB, > L2; no flush needed)" % ("int32" if args.elem == 4 else "int8", G_cfg * args.elem / 1e9),
              "parallelism": "blocks sharded over %d GPU(s)" % world}

    # ---------------- reference arm ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        r = run_reference(args, steps, warmup)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libxsi_ref.so not built"}))
            return
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "Ggt/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": config, "stats": {"sample_records_per_core": args.ref_records, "genotypes_per_step": r["genotypes_per_step"]},
                "compress_ggts": r["compress"], "decompress_ggts": r["decompress"], "verified": r["ok"],
                "cpu_baseline": {"value": r["value"], "unit": "Ggt/s", "cores": r["cores"], "kind": "reference",
                                 "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "Ggt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        if args.bcf_records > 0 and world == 1:
            line["e2e_bcf"] = bcf_legs(args.samples, args.bcf_records, "reference", min(16, usable_cores()))
        print(json.dumps(line))
        return

    # ---------------- B200 arm ----------------
    # the BCF legs run whole programs (own CUDA contexts): before this process creates its own and fills the device
    e2e_bcf = None
    if rank == 0 and world == 1 and args.bcf_records > 0 and not args.profile_only and args.shape == "hrc" and not args.sub:
        e2e_bcf = bcf_legs(args.samples, args.bcf_records, "both" if not args.no_cpu_baseline else "b200", min(16, usable_cores()))
    import torch
    import xsqueezeit_b200 as xb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries the one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 and not getenv_flag("XSI_BENCH_NO_NUMA") else None
    if numa is not None and "XSI_HOST_THREADS" not in os.environ:
        # the ranks whose GPUs share this node share its cores
        per_node = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world)) // max(1, len([n for n in os.listdir("/sys/devices/system/node") if n.startswith("node")])))
        os.environ["XSI_HOST_THREADS"] = str(max(2, numa["cores"] // per_node))
    if world > 1 and "XSI_HOST_THREADS" not in os.environ:
        # the ranks of one box share its cores: size each rank's conversion pool accordingly (read once by the library)
        os.environ["XSI_HOST_THREADS"] = str(max(2, len(os.sched_getaffinity(0)) // int(os.environ.get("LOCAL_WORLD_SIZE", world))))
    ctx = xb.Context(local_rank)
    ctx.profile(True)
    L = ctx._L
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    S, H, BL = args.samples, 2 * args.samples, args.block_len
    thr = xb.mac_threshold(S, 2, MAF)
    AET = 2 if S <= 65535 else 4  # header.aet_bytes: uint16 indices up to 65,535 samples (xsi_factory.hpp:425)
    if args.profile_only:
        steps, warmup = 1, 1

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- data: B blocks resident in HBM ----
    B = args.blocks
    R = B * BL
    G = R * H
    EL = args.elem
    rdt = torch.int32 if EL == 4 else torch.int8
    gt = torch.empty((R, H), dtype=rdt, device=dev)
    synth_gen = HrcSynth(S, 1002 + 131 * rank, dev, chrx=args.shape == "chrx")
    synth_gen.fill(gt)
    dec = torch.empty((R, H), dtype=rdt, device=dev)
    nal = np.concatenate(synth_gen.n_allele).astype(np.uint32)
    if args.shape == "chrx":
        args.resident_contexts, args.no_e2e, args.no_cpu_baseline = 0, True, True
    pos = xb.bm_positions(nal, BL)
    blk = (pos >> np.uint64(15)).astype(np.uint32)
    off = (pos & np.uint64(0x7FFF)).astype(np.uint32)
    torch.cuda.synchronize(dev)
    sizes_dev = torch.zeros(B, dtype=torch.int64, device=dev)
    gathered = torch.zeros(B * world, dtype=torch.int64, device=dev) if dist is not None else None

    def ev():
        return torch.cuda.Event(enable_timing=True)

    host_ms = {"encode_launch": 0.0, "encode_collect": 0.0, "decode_load_blocks": 0.0, "decode_records": 0.0}

    def encode_step(gt_ptr, on_device, n_rec, elem=4):
        t0 = time.perf_counter()
        ctx.encode_launch(gt_ptr, nal[:n_rec], S, BL, thr, 1, gt_on_device=on_device, gt_elem_bytes=elem)
        t1 = time.perf_counter()
        n = ctypes.c_uint32()
        bp = ctypes.POINTER(ctypes.c_void_p)()
        sz = ctypes.POINTER(ctypes.c_uint64)()
        ctx._check(L.xsi_encode_collect(ctx.h, ctypes.byref(n), ctypes.byref(bp), ctypes.byref(sz)))
        host_ms["encode_launch"] += (t1 - t0) * 1e3
        host_ms["encode_collect"] += (time.perf_counter() - t1) * 1e3
        blocks = [(bp[i], sz[i]) for i in range(n.value)]
        if dist is not None:  # the one exchange step: per-block byte counts -> global offset table
            sizes_dev[:n.value].copy_(torch.tensor([b[1] for b in blocks], dtype=torch.int64), non_blocking=False)
            dist.all_gather_into_tensor(gathered[:n.value * world], sizes_dev[:n.value])
            _ = torch.cumsum(gathered, 0)
        return blocks

    def decode_step(blocks, out_ptr, on_device, n_rec, elem=4):
        t0 = time.perf_counter()
        ctx.decode_load_blocks(blocks, S, AET)
        t1 = time.perf_counter()
        fn = L.xsi_decode_records if elem == 4 else L.xsi_decode_records_i8
        ctx._check(fn(ctx.h, n_rec, blk[:n_rec].ctypes.data, off[:n_rec].ctypes.data,
                      nal[:n_rec].ctypes.data, out_ptr, H, 1 if on_device else 0, None, None, 0))
        ctx.sync()
        host_ms["decode_load_blocks"] += (t1 - t0) * 1e3
        host_ms["decode_records"] += (time.perf_counter() - t1) * 1e3

    crcs = []

    def run_leg(n_rec, gt_ptr, out_ptr, on_device, nsteps, nwarm, sampler=None, elem=4):
        for _ in range(nwarm):
            decode_step(encode_step(gt_ptr, on_device, n_rec, elem), out_ptr, on_device, n_rec, elem)
        ctx.profile_read()
        for k in host_ms:
            host_ms[k] = 0.0
        launches0 = ctx.kernel_launches
        barrier()
        if sampler:
            sampler.start()
        e0, e1, e2 = [], [], []
        payload = 0
        for _ in range(nsteps):
            a, b, c = ev(), ev(), ev()
            a.record(stream)
            blocks = encode_step(gt_ptr, on_device, n_rec, elem)
            b.record(stream)
            decode_step(blocks, out_ptr, on_device, n_rec, elem)
            c.record(stream)
            e0.append(a); e1.append(b); e2.append(c)
            payload = sum(s for _, s in blocks)
            lines = ctx.encode_line_counts()
        barrier()
        clocks = sampler.stop() if sampler else None
        crcs[:] = [zlib.crc32(ctypes.string_at(p_, s_)) for p_, s_ in blocks]  # outside the timed region
        t_enc = sum(a.elapsed_time(b) for a, b in zip(e0, e1)) / 1e3
        t_dec = sum(b.elapsed_time(c) for b, c in zip(e1, e2)) / 1e3
        t_all = e0[0].elapsed_time(e2[-1]) / 1e3
        return dict(t_enc=t_enc, t_dec=t_dec, t_all=t_all, payload=payload, clocks=clocks, blocks=blocks,
                    launches=ctx.kernel_launches - launches0, prof=ctx.profile_read(), lines=lines,
                    host_ms={k: v / nsteps for k, v in host_ms.items()})

    def run_pipelined(nsteps, nwarm):
        """ONE context, ONE host thread, software-pipelined through xsi_encode_async: in every step the encode of the batch is
        handed to the library (its own thread and stream) and, while it runs, the blocks of the previous step are decoded on the
        context stream; xsi_encode_collect then closes the step.  A step still encodes one batch and decodes one batch."""
        def collect():
            n = ctypes.c_uint32()
            bp = ctypes.POINTER(ctypes.c_void_p)()
            sz = ctypes.POINTER(ctypes.c_uint64)()
            ctx._check(L.xsi_encode_collect(ctx.h, ctypes.byref(n), ctypes.byref(bp), ctypes.byref(sz)))
            return [(bp[i], sz[i]) for i in range(n.value)]

        # measured (profiles/r02r_*): launch first 651-672 Ggt/s, load first 628 -- the scan then waits 2 ms for the load's host work
        launch_first = os.environ.get("XSI_BENCH_PIPE_ORDER", "launch_first") == "launch_first"
        ctx.encode_async(True)
        try:
            ctx.encode_launch(gt.data_ptr(), nal[:R], S, BL, thr, 1, gt_on_device=True, gt_elem_bytes=EL)
            prev = collect()
            l0 = 0
            for it in range(nwarm + nsteps):
                if it == nwarm:
                    ctx.profile_read()
                    l0 = ctx.kernel_launches
                    barrier()
                    a = ev()
                    a.record(stream)
                if launch_first:
                    ctx.encode_launch(gt.data_ptr(), nal[:R], S, BL, thr, 1, gt_on_device=True, gt_elem_bytes=EL)
                    decode_step(prev, dec.data_ptr(), True, R, EL)
                else:
                    # A/B (XSI_BENCH_PIPE_ORDER=load_first): the SM-bound half of the decode (expand + inverse PBWT of batch i) enqueued
                    # BEFORE the encode of batch i+1, so that it does not fight the scan for SMs; the scan follows, and the compose kernels
                    # (which the library orders behind the scan) run beside the cluster kernel (device timeline: profiles/r02r_*)
                    ctx.decode_load_blocks(prev, S, AET)
                    ctx.encode_launch(gt.data_ptr(), nal[:R], S, BL, thr, 1, gt_on_device=True, gt_elem_bytes=EL)
                    fn = L.xsi_decode_records if EL == 4 else L.xsi_decode_records_i8
                    ctx._check(fn(ctx.h, R, blk[:R].ctypes.data, off[:R].ctypes.data, nal[:R].ctypes.data, dec.data_ptr(), H, 1, None, None, 0))
                    ctx.sync()
                prev = collect()
            z = ev()
            z.record(stream)
            barrier()
            ctx.profile_read()  # (XSI_TIMELINE: the kernels of this leg, both streams, on one time axis)
            return a.elapsed_time(z) / 1e3, ctx.kernel_launches - l0, [zlib.crc32(ctypes.string_at(p_, s_)) for p_, s_ in prev]
        finally:
            ctx.encode_async(False)

    def maxr(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank) if rank == 0 else None
    res = run_leg(R, gt.data_ptr(), dec.data_ptr(), True, steps, warmup, sampler, elem=EL)
    verified = all(bool(torch.equal(gt[r0:r0 + BL], dec[r0:r0 + BL])) for r0 in range(0, R, BL))  # block-wise: no batch-sized temporary
    t_all, t_enc, t_dec = maxr(res["t_all"]), maxr(res["t_enc"]), maxr(res["t_dec"])
    value = 2.0 * G * world * steps / t_all / 1e9

    # ---- the same context and host thread, encode and decode software-pipelined inside the library ----
    pipelined = None
    if not args.profile_only and not getenv_flag("XSI_BENCH_NO_PIPELINE"):
        dec.zero_()
        tp, lp, crc_p = run_pipelined(steps, max(1, warmup))
        tp = maxr(tp)
        okp = all(bool(torch.equal(gt[r0:r0 + BL], dec[r0:r0 + BL])) for r0 in range(0, R, BL)) and crc_p == crcs
        verified = verified and okp
        pipelined = {"value": 2.0 * G * world * steps / tp / 1e9, "unit": "Ggt/s", "ms_per_step": tp / steps * 1e3, "contexts": 1,
                     "host_threads": 1, "gpu_launches": lp, "verified": okp,
                     "how": "xsi_encode_async: batch i+1 encodes on the library's thread and stream while batch i is decoded by the caller"}

    # ---- N GPUs: ONE ordered file from the blocks the ranks encoded (outside the timed region) ----
    # every rank writes its first blocks at the offsets of the all-gathered table (xsqueezeit_b200/sharded.py, finalised by
    # xsi_writer_finalize_sharded); rank 0 also writes the same blocks through the single writer and compares the bytes
    sharded_file = None
    if dist is not None:
        from xsqueezeit_b200 import sharded
        nbw = min(2, B)
        mine = [ctypes.string_at(p_, s_) for p_, s_ in res["blocks"][:nbw]]
        tmpd = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
        tag = os.environ.get("MASTER_PORT", "0")
        path_sh = os.path.join(tmpd, "xsi_bench_sharded_%s.xsi" % tag)
        sharded.write_sharded(path_sh, rank, world, dist, dev, mine, nbw * world, S, None, BL, thr, 1, nbw * BL, nbw * BL, 2)
        allb = [None] * world if rank == 0 else None
        dist.gather_object(mine, allb, dst=0)
        if rank == 0:
            path_1w = os.path.join(tmpd, "xsi_bench_single_%s.xsi" % tag)
            w = ctypes.c_void_p()
            ok = L.xsi_writer_open(path_1w.encode(), S, None, BL, thr, 1, 0, 7, ctypes.byref(w)) == 0
            for blks in allb:
                for b_ in blks:
                    buf_ = (ctypes.c_uint8 * len(b_)).from_buffer_copy(b_)
                    ptr_ = (ctypes.c_void_p * 1)(ctypes.addressof(buf_))
                    sz_ = (ctypes.c_uint64 * 1)(len(b_))
                    ok = ok and L.xsi_writer_add_blocks(w, 1, ptr_, sz_, BL, BL) == 0
            ok = ok and L.xsi_writer_close(w, 2) == 0
            same = ok and open(path_sh, "rb").read() == open(path_1w, "rb").read()
            sharded_file = {"ranks": world, "blocks": nbw * world, "bytes": os.path.getsize(path_sh),
                            "equals_single_writer_file": bool(same)}
            for f_ in (path_sh, path_1w):
                try:
                    os.unlink(f_)
                except OSError:
                    pass
        dist.barrier()

    # ---- roofline of the dominant kernel (per-kernel CUDA-event times of the timed region) ----
    host_phases = {k: {"calls": n, "ms_per_step": ms / steps} for k, (n, ms) in res["prof"].items() if k.startswith("host:")}
    prof = {k: v for k, v in res["prof"].items() if not k.startswith("host:")}
    WS = ((H + 31) // 32 + 3) // 4 * 4
    roof = None
    roof_all = []
    kernels = {}
    if prof:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        top = max(prof, key=lambda k: prof[k][1])
        total_ms = sum(v[1] for v in prof.values())
        L_lines, L_wah = res["lines"]
        for k, (n, ms) in prof.items():
            kernels[k] = {"launches": n, "ms_per_step": ms / steps, "share": ms / total_ms if total_ms else None}
        traffic = {}
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if tj.get("blocks") == B and tj.get("elem") == EL and S == HRC_SAMPLES and BL == BLOCK_LEN:
                traffic = tj.get("bytes_per_launch", {})
        except Exception:
            pass

        def roof_of(k):
            ab = algorithmic_bytes(k, G, L_wah, L_lines, WS, res["payload"], EL)
            if not ab:
                return None
            n, ms = prof[k]
            ach = ab * steps / (ms / 1e3) / 1e9
            return {"kernel": k, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic.get(k), "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650",
                    "algorithmic_bytes_per_step": ab, "launches_per_step": n / steps, "ms_per_launch": ms / n}

        roof = roof_of(top)
        # every kernel with more than 5% of the step, so that the HBM-bound ones are judged next to the dominant one
        roof_all = [r for r in (roof_of(k) for k in sorted(prof, key=lambda k: -prof[k][1]) if prof[k][1] > 0.05 * total_ms) if r]

    # ---- resident leg from several contexts: the same batch, blocks split over W host threads with one xsi_ctx
    #      (stream + pools) each, so that one context's PBWT chain (SM-bound, 4 SMs per block) runs beside another's
    #      HBM-bound scan / compose kernels ----
    resident_mt = None
    if args.resident_contexts > 1 and not args.profile_only:
        W = min(args.resident_contexts, B)
        ctxs = [xb.Context(local_rank) for _ in range(W)]
        row_bytes = H * EL
        fn = L.xsi_decode_records if EL == 4 else L.xsi_decode_records_i8
        errs = []
        dec.zero_()
        torch.cuda.synchronize(dev)
        crc_single = list(crcs)
        crc_mt = [None] * B
        last_blocks = [None] * B

        def work(w, nrounds):
            c = ctxs[w]
            b0, b1 = w * B // W, (w + 1) * B // W
            r0, nr = b0 * BL, (b1 - b0) * BL
            try:
                for _ in range(nrounds):
                    c.encode_launch(gt.data_ptr() + r0 * row_bytes, nal[:nr], S, BL, thr, 1, gt_on_device=True, gt_elem_bytes=EL)
                    n = ctypes.c_uint32()
                    bp = ctypes.POINTER(ctypes.c_void_p)()
                    sz = ctypes.POINTER(ctypes.c_uint64)()
                    c._check(L.xsi_encode_collect(c.h, ctypes.byref(n), ctypes.byref(bp), ctypes.byref(sz)))
                    for i in range(n.value):
                        last_blocks[b0 + i] = (bp[i], sz[i])  # hashed after the timed region (valid until the next launch)
                    c.decode_load_blocks([(bp[i], sz[i]) for i in range(n.value)], S, AET)
                    c._check(fn(c.h, nr, blk[:nr].ctypes.data, off[:nr].ctypes.data, nal[:nr].ctypes.data,
                                dec.data_ptr() + r0 * row_bytes, H, 1, None, None, 0))
                    c.sync()
            except Exception as ex:
                errs.append(repr(ex))


        def run_mt(nrounds):
            ts = [threading.Thread(target=work, args=(w, nrounds)) for w in range(W)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
            if errs:
                raise SystemExit("bench.py: resident worker failed: " + errs[0])

        run_mt(max(1, warmup))
        s0 = torch.cuda.ExternalStream(ctxs[0].stream, device=dev)
        l0 = sum(c.kernel_launches for c in ctxs)
        barrier()
        a, z = ev(), ev()
        a.record(s0)
        run_mt(steps)
        z.record(s0)
        barrier()
        tm = maxr(a.elapsed_time(z) / 1e3)
        okm = all(bool(torch.equal(gt[r0:r0 + BL], dec[r0:r0 + BL])) for r0 in range(0, R, BL))
        verified = verified and okm
        crc_mt = [zlib.crc32(ctypes.string_at(p_, s_)) for p_, s_ in last_blocks]
        enc_same = crc_mt == crc_single
        bad_blocks = [r0 // BL for r0 in range(0, R, BL) if not bool(torch.equal(gt[r0:r0 + BL], dec[r0:r0 + BL]))]
        verified = verified and enc_same
        resident_mt = {"value": 2.0 * G * world * steps / tm / 1e9, "unit": "Ggt/s", "contexts": W, "host_threads": W,
                       "encoded_blocks_equal_single_context": enc_same, "blocks_decoded_wrong": bad_blocks,
                       "blocks_encoded_differently": [i for i in range(B) if crc_mt[i] != crc_single[i]],
                       "ms_per_step": tm / steps * 1e3, "gpu_launches": sum(c.kernel_launches for c in ctxs) - l0, "verified": okm}
        for c in ctxs:
            c.close()

    # ---- e2e: pinned host buffers through the same calls ----
    # `e2e` keeps the reference's own boundary types (int32 rows: bcf_get_genotypes in, fill_genotype_array out);
    # `e2e_bcf_int8` moves the records' raw BCF FORMAT/GT bytes instead (gt_elem_bytes = 1 in, xsi_decode_records_i8
    # out: SURVEY 8(f).1), a quarter of the PCIe traffic for the same genotypes.
    e2e = None
    e2e_i8 = None
    if not args.no_e2e and not args.profile_only:
        import psutil
        Be = max(1, min(args.e2e_blocks, B))
        while Be > 1 and 2 * Be * BL * H * 4 > 0.5 * psutil.virtual_memory().available / max(1, world):
            Be //= 2
        Re = Be * BL
        ns = max(1, min(steps, 3))

        # Host-buffer legs.  `serial` is one context doing encode(batch) then decode(batch): PCIe carries one
        # direction at a time.  `e2e` is the same calls from `--e2e-workers` host threads, each with its own
        # xsi_ctx (stream + pools) taking whole blocks round-robin, so that one block's upload overlaps another's
        # download and the kernels of a third: PCIe runs full duplex.  All work of all steps is inside the timed
        # region (CUDA events on worker 0's stream around start/join of the threads).
        def host_leg_mt(dtype, elem, label, serial, narrow=True):
            nonlocal verified
            import threading
            os.environ["XSI_HOST_NARROW"] = "1" if narrow else "0"  # read by the library at every call
            W = max(1, min(args.e2e_workers, Be))
            h_in = torch.empty((Re, H), dtype=dtype, pin_memory=True)
            h_out = torch.empty((Re, H), dtype=dtype, pin_memory=True)
            for r0 in range(0, Re, BL):
                h_in[r0:r0 + BL].copy_(gt[r0:r0 + BL].to(dtype))
            torch.cuda.synchronize(dev)
            r2 = run_leg(Re, h_in.data_ptr(), h_out.data_ptr(), False, ns, 1, elem=elem) if serial else None
            ok_serial = bool(torch.equal(h_in, h_out)) if serial else True
            h_out.zero_()
            ctxs = [xb.Context(local_rank) for _ in range(W)]
            fn = L.xsi_decode_records if elem == 4 else L.xsi_decode_records_i8
            row_bytes = H * elem
            payload = [0] * W
            errs = []

            def work(w, nrounds):
                c = ctxs[w]
                try:
                    for _ in range(nrounds):
                        payload[w] = 0
                        for b in range(w, Be, W):
                            r0 = b * BL
                            c.encode_launch(h_in.data_ptr() + r0 * row_bytes, nal[:BL], S, BL, thr, 1, gt_on_device=False, gt_elem_bytes=elem)
                            n = ctypes.c_uint32()
                            bp = ctypes.POINTER(ctypes.c_void_p)()
                            sz = ctypes.POINTER(ctypes.c_uint64)()
                            c._check(L.xsi_encode_collect(c.h, ctypes.byref(n), ctypes.byref(bp), ctypes.byref(sz)))
                            blocks = [(bp[i], sz[i]) for i in range(n.value)]
                            payload[w] += sum(x[1] for x in blocks)
                            c.decode_load_blocks(blocks, S, AET)
                            c._check(fn(c.h, BL, blk[:BL].ctypes.data, off[:BL].ctypes.data, nal[:BL].ctypes.data,
                                        h_out.data_ptr() + r0 * row_bytes, H, 0, None, None, 0))
                            c.sync()
                except Exception as ex:  # surfaced after join
                    errs.append(repr(ex))

            def run(nrounds):
                ts = [threading.Thread(target=work, args=(w, nrounds)) for w in range(W)]
                for t in ts:
                    t.start()
                for t in ts:
                    t.join()
                if errs:
                    raise SystemExit("bench.py: e2e worker failed: " + errs[0])

            run(1)  # warm-up: pools of every context sized
            s0 = torch.cuda.ExternalStream(ctxs[0].stream, device=dev)
            launches0 = sum(c.kernel_launches for c in ctxs)
            tr0 = [c.transport_stats for c in ctxs]
            barrier()
            a, z = ev(), ev()
            a.record(s0)
            run(ns)
            z.record(s0)
            barrier()
            t2 = maxr(a.elapsed_time(z) / 1e3)
            ok2 = bool(torch.equal(h_in, h_out)) and ok_serial
            verified = verified and ok2
            pl = sum(payload)
            # bytes that crossed the bus: rows the library moved in the int8 transport encoding count 1 byte per genotype
            tr1 = [c.transport_stats for c in ctxs]
            nar_h2d = sum(b_[0] - a_[0] for a_, b_ in zip(tr0, tr1)) // ns
            nar_d2h = sum(b_[1] - a_[1] for a_, b_ in zip(tr0, tr1)) // ns
            bus_h2d = Re * H * elem + pl - (nar_h2d * 3 if elem == 4 else 0)
            bus_d2h = Re * H * elem + pl - (nar_d2h * 3 if elem == 4 else 0)
            out = {"value": 2.0 * Re * H * world * ns / t2 / 1e9, "unit": "Ggt/s",
                   "h2d_bytes_per_step": bus_h2d, "d2h_bytes_per_step": bus_d2h,
                   "host_buffer_bytes_per_step_each_way": Re * H * elem,
                   "blocks_per_step": Be, "ms_per_step": t2 / ns * 1e3, "steps": ns, "host_buffers": label,
                   "host_threads": W, "contexts": W, "gpu_launches": sum(c.kernel_launches for c in ctxs) - launches0,
                   "pcie_gbs_each_way": (bus_h2d + bus_d2h) / 2 * ns / t2 / 1e9, "verified": ok2}
            if elem == 4:
                out["transport"] = ("int32 rows cross PCIe in their BCF int8 encoding, converted on %d host threads beside the DMA (csrc/host_narrow.cpp)"
                                    % L.xsi_host_threads()) if nar_h2d else "int32 over PCIe (XSI_HOST_NARROW=0)"
            if r2 is not None:
                ts_ = maxr(r2["t_all"])
                out["serial"] = {"value": 2.0 * Re * H * world * ns / ts_ / 1e9,
                                 "compress_ggts": Re * H * world * ns / maxr(r2["t_enc"]) / 1e9,
                                 "decompress_ggts": Re * H * world * ns / maxr(r2["t_dec"]) / 1e9,
                                 "ms_per_step": ts_ / ns * 1e3, "contexts": 1}
            for c in ctxs:
                c.close()
            del h_in, h_out
            return out

        # the resident leg's output buffer is no longer needed: make room for the worker contexts' pools
        del dec
        torch.cuda.empty_cache()
        e2e = host_leg_mt(torch.int32, 4, "pinned int32 rows in and out (bcf_get_genotypes / fill_genotype_array types)", True)
        e2e["int32_over_pcie"] = host_leg_mt(torch.int32, 4, "pinned int32 rows in and out, moved as int32 (XSI_HOST_NARROW=0)", False, narrow=False)
        os.environ["XSI_HOST_NARROW"] = "1"
        e2e_i8 = host_leg_mt(torch.int8, 1, "pinned int8 rows in and out (raw BCF FORMAT/GT payload)", True)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.profile_only:
        r = run_reference(args, 2, 1)
        if r is not None:
            cpu = {"value": r["value"], "unit": "Ggt/s", "cores": r["cores"], "kind": "reference", "sample": r["sample"],
                   "compress_ggts": r["compress"], "decompress_ggts": r["decompress"], "verified": r["ok"]}

    # ---- the other shapes: short sub-runs once this process has let go of its device memory ----
    shapes = None
    if not args.no_shapes and not args.profile_only and args.shape == "hrc" and S == HRC_SAMPLES:
        del gt
        if "dec" in dir():
            del dec
        ctx.close()
        ctx = None
        torch.cuda.empty_cache()
        if world == 1:
            shapes = {k: shape_subrun(k, local_rank) for k in SHAPES}
        else:
            # N GPUs: the biobank decode sweep of configs[4] -- every rank runs its own blocks, the rates add up (weak scaling)
            mine = shape_subrun("s5_biobank", local_rank)
            allr = [None] * world
            dist.all_gather_object(allr, mine)
            if rank == 0:
                ok = [r for r in allr if "value" in r]
                shapes = {"s5_biobank": dict(allr[0], per_rank_values=[r.get("value") for r in allr],
                                             value=sum(r["value"] for r in ok), compress_ggts=sum(r["compress_ggts"] for r in ok),
                                             decompress_ggts=sum(r["decompress_ggts"] for r in ok), n_gpus=world,
                                             verified=all(r.get("verified") for r in allr))}

    if rank == 0:
        # `value`: the better of the two resident legs (same batch, same calls, both verified); the per-kernel numbers
        # (`kernels`, `roofline`) always come from the one-context leg, where kernels do not overlap
        use_mt = resident_mt is not None and resident_mt["verified"] and resident_mt["value"] > value
        best, best_ms, mode = value, t_all / steps * 1e3, "one context, calls in sequence"
        if pipelined is not None and pipelined["verified"] and pipelined["value"] > best:
            best, best_ms, mode = pipelined["value"], pipelined["ms_per_step"], "one context, one host thread, xsi_encode_async pipeline"
        if use_mt and resident_mt["value"] > best:
            best, best_ms, mode = resident_mt["value"], resident_mt["ms_per_step"], "%d contexts on %d host threads" % (resident_mt["contexts"], resident_mt["contexts"])
        use_mt = use_mt and mode.endswith("host threads")
        line = {"metric": METRIC, "value": best, "unit": "Ggt/s", "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": best_ms, "value_mode": mode,
                "value_one_context": value, "ms_per_step_one_context": t_all / steps * 1e3, "one_context_pipelined": pipelined,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int32" if EL == 4 else "int8", "data": "synthetic",
                "config": config,
                "stats": {"xsi_payload_bytes_per_step": res["payload"], "binary_lines": res["lines"][0], "wah_lines": res["lines"][1],
                          "host_threads_per_gpu": resident_mt["contexts"] if use_mt else 1,
                          "contexts_per_gpu": resident_mt["contexts"] if use_mt else 1},
                "compress_ggts": G * world * steps / t_enc / 1e9, "decompress_ggts": G * world * steps / t_dec / 1e9,
                "verified": bool(verified and (sharded_file is None or sharded_file["equals_single_writer_file"])), "roofline": roof, "roofline_kernels": roof_all, "kernels": kernels, "call_wall_ms_per_step": res["host_ms"], "host_phases": host_phases, "resident_multi_context": resident_mt, "cpu_baseline": cpu, "e2e": e2e, "e2e_bcf_int8": e2e_i8, "e2e_bcf": e2e_bcf, "shapes": shapes, "sharded_file": sharded_file, "numa_binding": numa,
                "gpu_launches": resident_mt["gpu_launches"] if use_mt else (pipelined["gpu_launches"] if mode.startswith("one context, one host") else res["launches"]),
                "clocks": res["clocks"]}
        # the whole path against the HBM roofline: the boundary rows alone (EL bytes per genotype read by the encode, written by the
        # decode) over the step of `value`, per GPU -- the per-kernel fractions above say where the rest of the time goes
        try:
            hbm = float(roof["peak"]) if roof else None
            if hbm:
                gbs = 2.0 * G * EL / (best_ms * 1e-3) / 1e9
                line["whole_path"] = {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                                      "bytes_per_step_per_gpu": 2 * G * EL, "what": "boundary rows only: %d B per genotype in, %d B out" % (EL, EL)}
        except Exception:  # never lose the line over a derived number
            pass
        print(json.dumps(line))
    if ctx is not None:
        ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
