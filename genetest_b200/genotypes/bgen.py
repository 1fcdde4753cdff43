"""BGEN v1.2 / v1.3 files (layout 2; layout 1 as well) as a genotype
container.

The reference reads imputed data through geneparse's BGEN reader (third
party, not installed here); this is a native reader of the published format
with the same duck API.  Supported: unphased diploid bi-allelic variants,
uncompressed or zlib-compressed probability blocks, 8 / 16 / 32-bit (or any
bit width) probabilities, embedded or external (``.sample``) identifiers.
Not supported (``ValueError``): zstd compression, phased data, other
ploidies, multi-allelic variants.

As for :class:`~genetest_b200.genotypes.impute2.Impute2Reader` the semantics
are stated here: the dosage is the expected count of the second allele,
``P(AB) + 2 P(BB)`` (``coded = alleles[1]``, ``reference = alleles[0]``); a
sample flagged missing in the file, or whose best-guess probability is below
``probability_threshold`` (default 0.9), is NaN.
"""

import struct
import zlib

import numpy as np

from .core import Genotypes, GenotypesReader, Variant

__all__ = ["BgenReader"]


def _unpack_bits(buf, count, bits):
    """``count`` little-endian unsigned integers of ``bits`` bits each."""
    if bits == 8:
        return np.frombuffer(buf, dtype=np.uint8, count=count).astype(np.float64)
    if bits == 16:
        return np.frombuffer(buf, dtype="<u2", count=count).astype(np.float64)
    if bits == 32:
        return np.frombuffer(buf, dtype="<u4", count=count).astype(np.float64)
    raw = np.unpackbits(np.frombuffer(buf, dtype=np.uint8), bitorder="little")
    raw = raw[:count * bits].reshape(count, bits).astype(np.float64)
    return raw @ (2.0 ** np.arange(bits))


class BgenReader(GenotypesReader):
    def __init__(self, filename, sample_filename=None,
                 probability_threshold=0.9):
        self.filename = filename
        self.prob_t = float(probability_threshold)
        self._f = open(filename, "rb")
        offset, = struct.unpack("<I", self._f.read(4))
        header_len, self.n_variants, self.n_samples = struct.unpack(
            "<III", self._f.read(12))
        if self._f.read(4) not in (b"bgen", b"\0\0\0\0"):
            raise ValueError("{}: not a BGEN file".format(filename))
        self._f.read(header_len - 20)                  # free data
        flags, = struct.unpack("<I", self._f.read(4))
        self.compression = flags & 3
        self.layout = (flags >> 2) & 15
        if self.layout not in (1, 2):
            raise ValueError("{}: unsupported BGEN layout {}".format(
                filename, self.layout))
        if self.compression == 2:
            raise ValueError("{}: zstd-compressed BGEN blocks are not "
                             "supported".format(filename))
        samples = None
        if flags >> 31:                                # sample identifier block
            _block_len, n = struct.unpack("<II", self._f.read(8))
            samples = []
            for _ in range(n):
                ln, = struct.unpack("<H", self._f.read(2))
                samples.append(self._f.read(ln).decode())
        if sample_filename is not None:
            with open(sample_filename) as f:
                f.readline(), f.readline()
                rows = [line.split() for line in f if line.strip()]
            samples = [r[1] for r in rows]
            if len(set(samples)) != len(samples):
                samples = ["{}_{}".format(r[0], r[1]) for r in rows]
        if samples is None:
            raise ValueError("{}: no sample identifiers (give a .sample "
                             "file)".format(filename))
        if len(samples) != self.n_samples:
            raise ValueError("{}: {} identifiers for {} samples".format(
                filename, len(samples), self.n_samples))
        self._samples = samples
        self._first = offset + 4

    # -- duck-typed reader API ---------------------------------------------
    def close(self):
        if self._f is not None:
            self._f.close()
            self._f = None

    def get_samples(self):
        return list(self._samples)

    def get_number_samples(self):
        return self.n_samples

    def get_number_variants(self):
        return self.n_variants

    def _string(self, fmt):
        ln, = struct.unpack(fmt, self._f.read(struct.calcsize(fmt)))
        return self._f.read(ln).decode()

    def _read_variant(self, want_dosage=True):
        f = self._f
        if self.layout == 1:
            n, = struct.unpack("<I", f.read(4))
        else:
            n = self.n_samples
        _vid, rsid, chrom = (self._string("<H"), self._string("<H"),
                             self._string("<H"))
        pos, = struct.unpack("<I", f.read(4))
        k = 2
        if self.layout == 2:
            k, = struct.unpack("<H", f.read(2))
        alleles = [self._string("<I") for _ in range(k)]
        variant = Variant(rsid or _vid, chrom, pos, tuple(alleles))
        if self.layout == 1:
            if self.compression:
                clen, = struct.unpack("<I", f.read(4))
                data = zlib.decompress(f.read(clen))
            else:
                data = f.read(6 * n)
            if not want_dosage:
                return variant, None
            probs = np.frombuffer(data, dtype="<u2", count=3 * n) \
                .astype(np.float64).reshape(n, 3) / 32768.0
            missing = probs.sum(axis=1) == 0.0
        else:
            total, = struct.unpack("<I", f.read(4))
            if self.compression:
                dlen, = struct.unpack("<I", f.read(4))
                blob = f.read(total - 4)
                if not want_dosage:
                    return variant, None
                data = zlib.decompress(blob)
                if len(data) != dlen:
                    raise ValueError("{}: corrupt block of {}".format(
                        self.filename, variant.name))
            else:
                data = f.read(total)
                if not want_dosage:
                    return variant, None
            n2, k2, pmin, pmax = struct.unpack("<IHBB", data[:8])
            if k != 2 or k2 != 2 or pmin != 2 or pmax != 2:
                raise ValueError("{}: {} is not a diploid bi-allelic variant"
                                 .format(self.filename, variant.name))
            ploidy = np.frombuffer(data, dtype=np.uint8, count=n2, offset=8)
            missing = (ploidy >> 7).astype(bool)
            phased, bits = data[8 + n2], data[9 + n2]
            if phased:
                raise ValueError("{}: phased data are not supported".format(
                    self.filename))
            vals = _unpack_bits(data[10 + n2:], 2 * n2, bits) \
                .reshape(n2, 2) / float(2 ** bits - 1)
            probs = np.column_stack((vals, 1.0 - vals.sum(axis=1)))
        dosage = probs[:, 1] + 2.0 * probs[:, 2]
        dosage[missing | (probs.max(axis=1) < self.prob_t)] = np.nan
        return variant, dosage

    def _make(self, variant, dosage):
        return Genotypes(variant, dosage, reference=variant.alleles[0],
                         coded=variant.alleles[1], multiallelic=False)

    def iter_genotypes(self):
        self._f.seek(self._first)
        for _ in range(self.n_variants):
            yield self._make(*self._read_variant())

    def iter_variants(self):
        self._f.seek(self._first)
        for _ in range(self.n_variants):
            yield self._read_variant(want_dosage=False)[0]

    def get_variant_by_name(self, name):
        out = []
        self._f.seek(self._first)
        for _ in range(self.n_variants):
            start = self._f.tell()
            variant, _ = self._read_variant(want_dosage=False)
            if variant.name == name:
                end = self._f.tell()
                self._f.seek(start)
                out.append(self._make(*self._read_variant()))
                self._f.seek(end)
        return out
