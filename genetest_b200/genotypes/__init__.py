"""Genotype containers (geneparse-compatible duck types)."""

from .core import Genotypes, GenotypesReader, Variant
from .dataframe import DataFrameReader
from .plink import PlinkReader, PlinkWriter, pack_rows, unpack_rows

# same registry shape as ``geneparse.parsers``
parsers = {"plink": PlinkReader, "dataframe": DataFrameReader}

__all__ = ["Genotypes", "GenotypesReader", "Variant", "DataFrameReader",
           "PlinkReader", "PlinkWriter", "pack_rows", "unpack_rows",
           "parsers"]
