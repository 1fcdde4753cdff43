"""Allele-frequency helpers (``genetest/statistics/descriptive.py:19-93``).

These host versions serve the variant predicates and the single-variant
path; inside a GWAS batch the engine computes the same quantity on the
device from exact genotype counts.
"""

import numpy as np

__all__ = ["get_maf", "get_sex_maf"]


def _orient(freq, minor, major):
    """Swap the alleles when the coded allele is the common one."""
    if freq > 0.5:
        return 1 - freq, major, minor, True
    return freq, minor, major, False


def get_maf(genotypes, minor, major):
    """Minor allele frequency of additive genotypes (NaN = missing).

    Returns ``(maf, minor, major, flipped)``; all-missing gives NaN and no
    flip.
    """
    g = np.asarray(genotypes, dtype=float)
    ok = ~np.isnan(g)
    if not ok.any():
        return np.nan, minor, major, False
    return _orient(g[ok].mean() / 2, minor, major)


def get_sex_maf(genotypes, sexes, minor, major):
    """MAF on a sex chromosome: males (``sexes == 1``) carry one copy and are
    coded 0/2, females (``0``) two.  Samples with a missing genotype or sex
    are left out."""
    g = np.asarray(genotypes, dtype=float)
    s = np.asarray(sexes, dtype=float)
    ok = ~(np.isnan(g) | np.isnan(s))
    if not ok.any():
        return np.nan, minor, major, False
    g, s = g[ok], s[ok]
    n_male = s.sum()
    n_female = s.shape[0] - n_male
    freq = (g.sum() - 0.5 * np.dot(g, s)) / (n_male + 2 * n_female)
    return _orient(freq, minor, major)
