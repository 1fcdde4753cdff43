"""Genotype containers (geneparse-compatible duck types)."""

from .bgen import BgenReader
from .core import Genotypes, GenotypesReader, Variant
from .dataframe import DataFrameReader
from .impute2 import Impute2Reader
from .plink import PlinkReader, PlinkWriter, pack_rows, unpack_rows

# same registry shape as ``geneparse.parsers``
parsers = {"plink": PlinkReader, "dataframe": DataFrameReader,
           "impute2": Impute2Reader, "bgen": BgenReader}

__all__ = ["Genotypes", "GenotypesReader", "Variant", "DataFrameReader",
           "Impute2Reader", "BgenReader",
           "PlinkReader", "PlinkWriter", "pack_rows", "unpack_rows",
           "parsers"]
