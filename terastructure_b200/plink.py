"""PLINK .bed ingest that keeps genotypes 2-bit packed (reference: SNP::read_bed,
src/snp.cc:95-253, which unpacks to one byte per genotype)."""
import os

import numpy as np

# PLINK 2-bit code -> the reference's y (snp.cc:203-216): 00->0, 01->3 (missing), 10->1, 11->2
CODE_TO_Y = np.array([0, 3, 1, 2], dtype=np.uint8)
Y_TO_CODE = np.array([0, 2, 3, 1], dtype=np.uint8)


def _count_lines(path):
    with open(path, "rb") as f:
        return sum(1 for _ in f)


def read_bed(path, n, l):
    """Return SNP-major packed rows [l, ceil(n/4)] (uint8).  Mirrors the reference's checks:
    .bim/.fam line counts must match -l/-n (snp.cc:103-139), magic 6C 1B, mode 01 only."""
    prefix = path[:-4]
    lb = _count_lines(prefix + ".bim")
    if lb != l:
        raise ValueError("-l input doesn't match SNPs in bim file")
    nf = _count_lines(prefix + ".fam")
    if nf != n:
        raise ValueError("-n input doesn't match individuals in fam file")
    bps = (n + 3) // 4
    with open(path, "rb") as f:
        hdr = f.read(3)
        if hdr[:2] != b"\x6c\x1b":
            raise ValueError(f"{path} magic number incorrect")
        if hdr[2] == 0:
            raise ValueError("individual major mode not supported yet!")
        if hdr[2] != 1:
            raise ValueError(f"mode problem in {path}")
        rows = np.fromfile(f, dtype=np.uint8, count=l * bps)
    if rows.size != l * bps:
        raise ValueError(f"{path}: short file")
    return rows.reshape(l, bps)


def unpack(rows, n):
    """[l, bytes] packed -> y[l, n] in {0,1,2,3=missing} (what the reference holds in RAM)."""
    l = rows.shape[0]
    codes = np.empty((l, rows.shape[1] * 4), dtype=np.uint8)
    for j in range(4):
        codes[:, j::4] = (rows >> (2 * j)) & 3
    return CODE_TO_Y[codes[:, :n]]


def pack(y):
    """y[l, n] in {0,1,2,3} -> packed rows [l, ceil(n/4)]."""
    l, n = y.shape
    bps = (n + 3) // 4
    codes = np.zeros((l, bps * 4), dtype=np.uint8)
    codes[:, :n] = Y_TO_CODE[y]
    rows = np.zeros((l, bps), dtype=np.uint8)
    for j in range(4):
        rows |= codes[:, j::4] << (2 * j)
    return rows


def write_bed(prefix, rows, n):
    """Write prefix.bed/.bim/.fam (the reference only line-counts .bim/.fam)."""
    l = rows.shape[0]
    with open(prefix + ".bed", "wb") as f:
        f.write(b"\x6c\x1b\x01")
        f.write(np.ascontiguousarray(rows[:, :(n + 3) // 4]).tobytes())
    with open(prefix + ".bim", "w") as f:
        f.write("".join(f"1\tsnp{i}\t0\t{i + 1}\tA\tC\n" for i in range(l)))
    with open(prefix + ".fam", "w") as f:
        f.write("".join(f"f{i} i{i} 0 0 0 -9\n" for i in range(n)))
    return os.path.abspath(prefix + ".bed")
