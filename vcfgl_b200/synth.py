"""Synthetic "msprime-shaped" genotype inputs.

tskit VCF dialect: one contig "1" of length L, REF=0 ALT=1, phased a|b GT,
samples tsk_0.., biallelic sites whose derived-allele count k follows the
neutral site-frequency spectrum P(k) ~ 1/k on [1, 2S-1]; the k carrier
haplotypes are chosen uniformly without replacement.  The same generator
produces (a) VCF text for the reference CPU binary and (b) the packed-nibble
genotype matrix the C-ABI consumes (include/vgl.h), from one seed.
"""
from __future__ import annotations

import numpy as np

GT_MISSING = 0xF


def sfs_genotypes(n_sites: int, n_samples: int, seed: int, missing_rate: float = 0.0) -> np.ndarray:
    """-> int8 [n_sites, 2*n_samples] binary haplotype alleles (0/1, -1 missing)."""
    rng = np.random.default_rng(seed)
    H = 2 * n_samples
    ks = np.arange(1, H)
    p = 1.0 / ks
    p /= p.sum()
    k = rng.choice(ks, size=n_sites, p=p)
    # k smallest of H iid uniforms per site == uniform k-subset
    u = rng.random((n_sites, H))
    thresh = np.sort(u, axis=1)[np.arange(n_sites), k - 1]
    hap = (u <= thresh[:, None]).astype(np.int8)
    if missing_rate > 0:
        miss = rng.random((n_sites, n_samples)) < missing_rate
        hap = hap.reshape(n_sites, n_samples, 2)
        hap[miss] = -1
        hap = hap.reshape(n_sites, H)
    return hap


def pack_gt(hap_acgt: np.ndarray) -> np.ndarray:
    """int8 [n_sites, 2S] ACGT ints (-1 missing) -> uint8 [n_sites, S]:
    low nibble = first haplotype, high nibble = second, 0xF = missing."""
    h = hap_acgt.astype(np.int16)
    h = np.where(h < 0, GT_MISSING, h).astype(np.uint8)
    return (h[:, 0::2] | (h[:, 1::2] << 4)).astype(np.uint8)


def positions(n_sites: int, length: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed + 7)
    if n_sites > length:
        raise ValueError("more sites than positions")
    if n_sites * 4 > length:
        return np.sort(rng.choice(length, size=n_sites, replace=False)) + 1
    pos = np.unique(rng.integers(1, length + 1, size=int(n_sites * 1.2) + 16))
    while len(pos) < n_sites:
        pos = np.unique(np.concatenate([pos, rng.integers(1, length + 1, size=n_sites)]))
    return np.sort(rng.choice(pos, size=n_sites, replace=False))


def vcf_header(n_samples: int, length: int, contig: str = "1") -> bytes:
    return ("##fileformat=VCFv4.2\n##source=vcfgl_b200.synth\n"
            "##FILTER=<ID=PASS,Description=\"All filters passed\">\n"
            "##contig=<ID=%s,length=%d>\n"
            "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n"
            "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t%s\n"
            % (contig, length, "\t".join("tsk_%d" % i for i in range(n_samples)))).encode()


def vcf_body(hap: np.ndarray, pos: np.ndarray, contig: str = "1", ref: str = "0", alt: str = "1") -> bytes:
    """the record lines of a tskit-style VCF: one 4-byte column per sample, "a|b\t" (".|.\t" when missing)"""
    n_sites, H = hap.shape
    S = H // 2
    ch = np.where(hap < 0, ord("."), hap + ord("0")).astype(np.uint8)
    body = np.empty((n_sites, S, 4), np.uint8)
    body[:, :, 0] = ch[:, 0::2]
    body[:, :, 1] = ord("|")
    body[:, :, 2] = ch[:, 1::2]
    body[:, :, 3] = ord("\t")
    body = body.reshape(n_sites, S * 4)
    body[:, -1] = ord("\n")
    out = []
    for i in range(n_sites):
        out.append(("%s\t%d\t.\t%s\t%s\t.\tPASS\t.\tGT\t" % (contig, pos[i], ref, alt)).encode())
        out.append(body[i].tobytes())
    return b"".join(out)


def write_vcf(path: str, hap: np.ndarray, pos: np.ndarray, length: int, contig: str = "1",
              ref: str = "0", alt: str = "1") -> None:
    """Write binary haplotypes as a tskit-style VCF (for the reference CPU binary)."""
    with open(path, "wb") as fh:
        fh.write(vcf_header(hap.shape[1] // 2, length, contig))
        fh.write(vcf_body(hap, pos, contig, ref, alt))
