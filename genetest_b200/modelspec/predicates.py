"""Variant predicates: callables deciding whether a marker is analysed
(``genetest/modelspec/predicates.py``).  They may raise ``StopIteration`` to
stop the scan."""

from ..statistics.descriptive import get_maf

__all__ = ["MAFFilter", "NameFilter", "VariantPredicate"]


class VariantPredicate(object):
    def __call__(self, snp):
        raise NotImplementedError()


class MAFFilter(VariantPredicate):
    """Keep markers whose MAF (all genotyped samples) is >= ``maf``."""

    def __init__(self, maf):
        self.maf = maf

    def __call__(self, snp):
        return get_maf(snp.genotypes, None, None)[0] >= self.maf


class NameFilter(VariantPredicate):
    """Keep markers whose name is in ``extract``."""

    def __init__(self, extract):
        self.extract = extract

    def __call__(self, snp):
        return snp.variant.name in self.extract
