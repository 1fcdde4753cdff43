"""Duck-typed genotype objects the analysis layer consumes.

The reference takes these from geneparse (``geneparse.core``: ``Variant``,
``Genotypes``; call sites ``genetest/analysis.py:116,138-139,192-193``):
``obj.genotypes`` (float ndarray, NaN = missing, in ``get_samples()`` order),
``obj.variant.name/.chrom/.pos``, ``obj.coded``, ``obj.reference``.  geneparse
is not installed in this image, so the same small value classes live here.
"""

__all__ = ["Variant", "Genotypes", "GenotypesReader"]


class Variant(object):
    __slots__ = ("name", "chrom", "pos", "alleles")

    def __init__(self, name, chrom, pos, alleles):
        self.name = name
        self.chrom = chrom
        self.pos = pos
        self.alleles = tuple(alleles)

    def __repr__(self):
        return "<Variant {} chr{}:{}_{}>".format(
            self.name, self.chrom, self.pos, "/".join(self.alleles))


class Genotypes(object):
    """One marker: additive dosage of the ``coded`` allele per sample."""
    __slots__ = ("variant", "genotypes", "reference", "coded", "multiallelic")

    def __init__(self, variant, genotypes, reference, coded,
                 multiallelic=False):
        self.variant = variant
        self.genotypes = genotypes
        self.reference = reference
        self.coded = coded
        self.multiallelic = multiallelic


class GenotypesReader(object):
    """Interface of a genotype container (geneparse ``GenotypesReader``)."""

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        pass

    def get_samples(self):
        raise NotImplementedError()

    def iter_genotypes(self):
        raise NotImplementedError()

    def get_variant_by_name(self, name):
        raise NotImplementedError()

    def get_number_samples(self):
        return len(self.get_samples())

    def get_number_variants(self):
        raise NotImplementedError()
