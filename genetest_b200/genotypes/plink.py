"""PLINK binary (.bed/.bim/.fam) reader and writer.

Stands in for geneparse's ``parsers["plink"]`` / pyplink on the hot path
(``genetest/analysis.py:670``; the reference's tests write their fixtures with
pyplink, ``genetest/tests/test_linear.py:85-119``).  On top of the duck-typed
reader API it exposes the raw SNP-major 2-bit rows (:meth:`PlinkReader.packed`)
so that the engine can stream them to the GPU without decoding on the host.

.bed coding (SNP-major, magic ``6C 1B 01``), two bits per sample, low bits
first: ``00`` homozygous A1, ``01`` missing, ``10`` heterozygous, ``11``
homozygous A2.  The additive value is the number of A1 (``.bim`` column 5)
alleles: A1 is ``coded``, A2 is ``reference``.
"""

import os

import numpy as np
import pandas as pd

from .core import Genotypes, GenotypesReader, Variant

__all__ = ["PlinkReader", "PlinkWriter", "pack_rows", "unpack_rows",
           "BED_MAGIC"]

BED_MAGIC = bytes([0x6C, 0x1B, 0x01])

# additive value -> 2-bit code
_CODE = {2: 0b00, 1: 0b10, 0: 0b11, -1: 0b01}
_DECODE = np.array([2.0, np.nan, 1.0, 0.0])


def pack_rows(geno):
    """``geno``: int array [n_snps, n_samples] with 0/1/2 (count of A1) and -1
    (or NaN) for missing -> uint8 [n_snps, ceil(n_samples/4)]."""
    g = np.asarray(geno)
    if g.dtype.kind == "f":
        g = np.where(np.isnan(g), -1, g).astype(np.int8)
    g = np.atleast_2d(g)
    n_snps, n = g.shape
    codes = np.full(g.shape, 0b01, dtype=np.uint8)
    codes[g == 2] = 0b00
    codes[g == 1] = 0b10
    codes[g == 0] = 0b11
    pad = (-n) % 4
    if pad:
        codes = np.concatenate(
            [codes, np.zeros((n_snps, pad), dtype=np.uint8)], axis=1)
    codes = codes.reshape(n_snps, -1, 4)
    return (codes[:, :, 0] | (codes[:, :, 1] << 2) | (codes[:, :, 2] << 4) |
            (codes[:, :, 3] << 6)).astype(np.uint8)


def unpack_rows(packed, n_samples):
    """uint8 [n_snps, bps] -> float64 [n_snps, n_samples], NaN = missing."""
    p = np.atleast_2d(np.asarray(packed, dtype=np.uint8))
    codes = np.stack([(p >> s) & 3 for s in (0, 2, 4, 6)], axis=2)
    codes = codes.reshape(p.shape[0], -1)[:, :n_samples]
    return _DECODE[codes]


class PlinkReader(GenotypesReader):
    def __init__(self, prefix):
        self.prefix = prefix
        # .fam / .bim: six whitespace-separated fields per line.  One split of the whole file
        # and strided slices of the token list (0.16 s for 262k markers + 100k samples; the
        # pandas C parser with string columns took 0.5 s).
        def six_columns(path):
            with open(path) as f:
                text = f.read()
            tok = text.split()
            n_lines = sum(1 for line in text.splitlines() if line and not line.isspace())
            if len(tok) != 6 * n_lines:
                raise ValueError("{}: every line needs six whitespace-separated "
                                 "fields".format(path))
            return [tok[j::6] for j in range(6)]

        if os.path.getsize(prefix + ".fam"):
            cols = six_columns(prefix + ".fam")
            self._fam = [list(r) for r in zip(*cols)]
        else:
            self._fam = []
        fids = [r[0] for r in self._fam]
        iids = [r[1] for r in self._fam]
        # geneparse's plink reader uses the IID, falling back to FID_IID when
        # the IIDs are not unique
        if len(set(iids)) == len(iids):
            self._samples = iids
        else:
            self._samples = ["{}_{}".format(a, b) for a, b in zip(fids, iids)]
        # [(name, chrom, pos, a1, a2)]
        if os.path.getsize(prefix + ".bim"):
            cols = six_columns(prefix + ".bim")
            self._bim = list(zip(cols[1], cols[0], [int(x) for x in cols[3]],
                                 cols[4], cols[5]))
        else:
            self._bim = []
        self._index_cache = None
        self.n_samples = len(self._samples)
        self.n_variants = len(self._bim)
        self.row_bytes = (self.n_samples + 3) // 4
        with open(prefix + ".bed", "rb") as f:
            magic = f.read(3)
        if magic != BED_MAGIC:
            raise ValueError("{}.bed: not a SNP-major PLINK binary file"
                             .format(prefix))
        expected = 3 + self.row_bytes * self.n_variants
        if os.path.getsize(prefix + ".bed") != expected:
            raise ValueError("{}.bed: size does not match .bim/.fam"
                             .format(prefix))
        if self.n_variants:
            self._bed = np.memmap(prefix + ".bed", dtype=np.uint8, mode="r",
                                  offset=3,
                                  shape=(self.n_variants, self.row_bytes))
        else:
            self._bed = np.zeros((0, self.row_bytes), dtype=np.uint8)

    # -- duck-typed reader API -------------------------------------------
    def close(self):
        self._bed = None

    def get_samples(self):
        return list(self._samples)

    def get_number_samples(self):
        return self.n_samples

    def get_number_variants(self):
        return self.n_variants

    def _make(self, i, row=None):
        name, chrom, pos, a1, a2 = self._bim[i]
        if row is None:
            row = unpack_rows(self._bed[i:i + 1], self.n_samples)[0]
        return Genotypes(Variant(name, chrom, pos, (a1, a2)), row,
                         reference=a2, coded=a1, multiallelic=False)

    def iter_genotypes(self):
        step = 256
        for s in range(0, self.n_variants, step):
            block = unpack_rows(self._bed[s:s + step], self.n_samples)
            for j in range(block.shape[0]):
                yield self._make(s + j, block[j])

    def iter_variants(self):
        for name, chrom, pos, a1, a2 in self._bim:
            yield Variant(name, chrom, pos, (a1, a2))

    @property
    def _index(self):
        if self._index_cache is None:
            index = {}
            for i, b in enumerate(self._bim):
                index.setdefault(b[0], []).append(i)
            self._index_cache = index
        return self._index_cache

    def get_variant_by_name(self, name):
        return [self._make(i) for i in self._index.get(name, [])]

    # -- fast path ---------------------------------------------------------
    def packed(self):
        """The .bed body: uint8 [n_variants, ceil(n_samples/4)] (memmap)."""
        return self._bed

    def variant_table(self):
        """[(name, chrom, pos, a1, a2)] in file order."""
        return self._bim


class PlinkWriter(object):
    """Minimal SNP-major .bed/.bim/.fam writer (tests, benchmarks)."""

    def __init__(self, prefix):
        self.prefix = prefix
        self._bed = open(prefix + ".bed", "wb")
        self._bed.write(BED_MAGIC)
        self._n = None

    def write_genotypes(self, genotypes):
        row = pack_rows(np.asarray(genotypes)[None, :])
        if self._n is None:
            self._n = len(genotypes)
        elif self._n != len(genotypes):
            raise ValueError("inconsistent number of samples")
        self._bed.write(row.tobytes())

    def write_packed(self, packed):
        self._bed.write(np.ascontiguousarray(packed, dtype=np.uint8).tobytes())

    def write_bim(self, rows):
        with open(self.prefix + ".bim", "w") as f:
            for chrom, name, pos, a1, a2 in rows:
                print(chrom, name, 0, pos, a1, a2, sep="\t", file=f)

    def write_fam(self, samples):
        with open(self.prefix + ".fam", "w") as f:
            for s in samples:
                print(s, s, 0, 0, 0, -9, file=f)

    def close(self):
        self._bed.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
