"""Synthetic genotype matrices for tests (numpy, seeded).  htslib GT encoding:
value = (allele+1)<<1 | phased ; missing = 0|phased ; bcf_int32_missing = INT32_MIN ;
end-of-vector = INT32_MIN+1 (htslib/htslib/vcf.h:892-898,1324,1329)."""
import numpy as np

EOV = np.int32(-2147483647)
I32_MISSING = np.int32(-2147483648)


def ld_matrix(n_records, n_samples, seed, n_founders=64, switch=2e-3, flip=1e-4, max_alt=1,
              multi_frac=0.0, fmin=None):
    """Haplotype-copying model: returns allele matrix [n_records, 2*n_samples] (int8, 0..max_alt) and n_allele."""
    rng = np.random.default_rng(seed)
    H = 2 * n_samples
    fmin = fmin or 1.0 / H
    f = np.exp(rng.uniform(np.log(fmin), np.log(0.5), size=n_records))
    founders = (rng.random((n_records, n_founders)) < f[:, None]).astype(np.int8)
    # copy path: piecewise constant founder index per haplotype
    path = np.empty((n_records, H), dtype=np.int16)
    cur = rng.integers(0, n_founders, size=H)
    for r in range(n_records):
        sw = rng.random(H) < switch
        if sw.any():
            cur = np.where(sw, rng.integers(0, n_founders, size=H), cur)
        path[r] = cur
    alleles = np.take_along_axis(founders, path.astype(np.int64), axis=1)
    alleles ^= (rng.random((n_records, H)) < flip).astype(np.int8)
    n_allele = np.full(n_records, 2, dtype=np.int32)
    if max_alt > 1 and multi_frac > 0:
        multi = rng.random(n_records) < multi_frac
        for r in np.nonzero(multi)[0]:
            k = int(rng.integers(2, max_alt + 1))
            n_allele[r] = k + 1
            ones = np.nonzero(alleles[r])[0]
            if ones.size:
                alleles[r, ones] = rng.integers(1, k + 1, size=ones.size).astype(np.int8)
    return alleles, n_allele


def encode_gt(alleles, phased=1):
    return ((alleles.astype(np.int32) + 1) << 1) | np.int32(phased)


def make_dataset(n_records, n_samples, seed, max_alt=1, multi_frac=0.0, missing=0.0, unphased=0.0,
                 haploid_samples=0.0, phased=1, n_founders=64, fmin=None):
    """Returns dict(gt flat int32, ngt, n_allele, n_samples). Every record is diploid-shaped (ngt = 2*S);
    `haploid_samples` fraction of samples get an end-of-vector second allele (chrX-like males)."""
    rng = np.random.default_rng(seed + 7919)
    alleles, n_allele = ld_matrix(n_records, n_samples, seed, n_founders=n_founders, max_alt=max_alt,
                                  multi_frac=multi_frac, fmin=fmin)
    gt = encode_gt(alleles, phased)
    H = 2 * n_samples
    if unphased > 0:
        m = rng.random((n_records, H)) < unphased
        m[:, 0::2] = False
        gt = np.where(m, gt ^ 1, gt)
    if missing > 0:
        m = rng.random((n_records, H)) < missing
        gt = np.where(m, np.int32(phased) * (np.arange(H) & 1).astype(np.int32)[None, :], gt)
    if haploid_samples > 0:
        males = rng.random(n_samples) < haploid_samples
        cols = np.nonzero(males)[0] * 2 + 1
        gt[:, cols] = EOV
    ngt = np.full(n_records, H, dtype=np.int32)
    return dict(gt=np.ascontiguousarray(gt.reshape(-1), dtype=np.int32), ngt=ngt, n_allele=n_allele,
                n_samples=n_samples)


def haploid_multiallelic_block(seed=77, ns=120, nrec=120):
    """All-haploid records and multi-allelic records in the same PBWT block: the shape the reference writes but cannot read
    back (gt_block.hpp:219-224,639-642 vs accessor_internals_new.hpp:116,165,265-271)."""
    rng = np.random.default_rng(seed)
    rows, ngt, nal = [], [], []
    for r in range(nrec):
        p = 1 if r % 7 == 3 else 2
        k = 3 if r % 5 == 1 else 2
        al = (rng.random(ns * p) < 0.3).astype(np.int8)
        if k == 3:
            al = np.where(al > 0, rng.integers(1, 3, size=al.size), 0).astype(np.int8)
        rows.append(encode_gt(al, 1 if p == 2 else 0))
        ngt.append(ns * p)
        nal.append(k)
    return dict(gt=np.concatenate(rows).astype(np.int32), ngt=np.array(ngt, np.int32), n_allele=np.array(nal, np.int32), n_samples=ns)
