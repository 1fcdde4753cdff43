"""Phenotype containers (``genetest/phenotypes/__init__.py``)."""

from .core import PhenotypesContainer
from .dataframe import DataFramePhenotypes
from .dummy import _DummyPhenotypes
from .text import TextPhenotypes

format_map = {"text": TextPhenotypes, "dataframe": DataFramePhenotypes}

__all__ = ["PhenotypesContainer", "DataFramePhenotypes", "TextPhenotypes",
           "_DummyPhenotypes", "format_map"]
