"""IMPUTE2 ``.impute2`` / ``.gen`` text files as a genotype container.

The reference reads imputed data through geneparse's ``Impute2Reader``
(third party, not installed here); this is a native reader with the same duck
API.  One marker per line::

    chrom name pos a1 a2  p(a1a1) p(a1a2) p(a2a2)  [three probabilities per sample ...]

and the usual two-header-line ``.sample`` file.  What a marker yields is
stated here rather than inherited, because the upstream semantics cannot be
checked in this image:

* the dosage is the expected count of ``a2``: ``p(a1a2) + 2 p(a2a2)``
  (``coded = a2``, ``reference = a1``);
* a sample whose best-guess probability is below ``probability_threshold``
  (default 0.9), or whose three probabilities are all 0, is missing (NaN).

Real-valued dosages go through the engine's FP64 dosage kernels
(``gt_run_dosage_host``).
"""

import gzip

import numpy as np

from .core import Genotypes, GenotypesReader, Variant

__all__ = ["Impute2Reader"]


def _open(path):
    return gzip.open(path, "rt") if str(path).endswith(".gz") else open(path)


class Impute2Reader(GenotypesReader):
    def __init__(self, filename, sample_filename, probability_threshold=0.9):
        self.filename = filename
        self.prob_t = float(probability_threshold)
        with _open(sample_filename) as f:
            header = f.readline().split()
            f.readline()                                # the column-type line
            rows = [line.split() for line in f if line.strip()]
        if len(header) < 2 or header[:2] != ["ID_1", "ID_2"]:
            raise ValueError("{}: not an IMPUTE2 sample file".format(
                sample_filename))
        ids = [r[1] for r in rows]
        if len(set(ids)) != len(ids):                   # fall back to ID_1_ID_2
            ids = ["{}_{}".format(r[0], r[1]) for r in rows]
        self._samples = ids
        self._n_variants = None

    def get_samples(self):
        return list(self._samples)

    def get_number_samples(self):
        return len(self._samples)

    def get_number_variants(self):
        if self._n_variants is None:
            with _open(self.filename) as f:
                self._n_variants = sum(1 for line in f if line.strip())
        return self._n_variants

    def _parse(self, line):
        fields = line.split(None, 5)
        if len(fields) < 6:
            raise ValueError("{}: malformed line".format(self.filename))
        chrom, name, pos, a1, a2, rest = fields
        probs = np.array(rest.split(), dtype=np.float64)
        n = len(self._samples)
        if probs.shape[0] != 3 * n:
            raise ValueError("{}: {} has {} probabilities for {} samples"
                             .format(self.filename, name, probs.shape[0], n))
        probs = probs.reshape(n, 3)
        dosage = probs[:, 1] + 2.0 * probs[:, 2]
        best = probs.max(axis=1)
        dosage[(best < self.prob_t) | (probs.sum(axis=1) == 0.0)] = np.nan
        return Genotypes(Variant(name, chrom, int(pos), (a1, a2)), dosage,
                         reference=a1, coded=a2, multiallelic=False)

    def iter_genotypes(self):
        with _open(self.filename) as f:
            for line in f:
                if line.strip():
                    yield self._parse(line)

    def iter_variants(self):
        with _open(self.filename) as f:
            for line in f:
                if line.strip():
                    chrom, name, pos, a1, a2 = line.split(None, 5)[:5]
                    yield Variant(name, chrom, int(pos), (a1, a2))

    def get_variant_by_name(self, name):
        out = []
        with _open(self.filename) as f:
            for line in f:
                head = line.split(None, 2)
                if len(head) > 1 and head[1] == name:
                    out.append(self._parse(line))
        return out
