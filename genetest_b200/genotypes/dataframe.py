"""In-memory genotype container backed by a pandas DataFrame.

Equivalent of geneparse's ``parsers["dataframe"]`` as used by the reference's
tests (``genetest/tests/test_logistic.py:88-106``): rows are samples, columns
markers, values are real-valued dosages of ``a1`` (NaN = missing);
``map_info`` (index = marker) has ``chrom``, ``pos``, ``a1``, ``a2``.
"""

import numpy as np

from .core import Genotypes, GenotypesReader, Variant

__all__ = ["DataFrameReader"]


class DataFrameReader(GenotypesReader):
    def __init__(self, dataframe, map_info):
        self.df = dataframe
        self.map_info = map_info

    def get_samples(self):
        return list(self.df.index)

    def get_number_variants(self):
        return self.df.shape[1]

    def _make(self, name):
        info = self.map_info.loc[name]
        return Genotypes(
            Variant(name, info.chrom, info.pos, (info.a1, info.a2)),
            self.df[name].values.astype(float), reference=info.a2,
            coded=info.a1, multiallelic=False)

    def iter_genotypes(self):
        for name in self.df.columns:
            yield self._make(name)

    def get_variant_by_name(self, name):
        if name not in self.df.columns:
            return []
        return [self._make(name)]

    def dosage_matrix(self):
        """float64 [n_variants, n_samples] (markers in column order)."""
        return np.ascontiguousarray(self.df.values.T, dtype=np.float64)
