"""Delimited text file of phenotypes (``genetest/phenotypes/text.py``)."""

import pandas as pd

from .dataframe import DataFramePhenotypes

__all__ = ["TextPhenotypes"]


class TextPhenotypes(DataFramePhenotypes):
    def __init__(self, filename, sample_column="sample", field_separator="\t",
                 missing_values=None, repeated_measurements=False,
                 keep_sample_column=False, sex_column=None):
        df = pd.read_csv(filename, sep=field_separator,
                         na_values=missing_values)
        df = df.set_index(sample_column, drop=not keep_sample_column)
        df.index = df.index.astype(str)
        super().__init__(df, repeated_measurements=repeated_measurements,
                         sex_column=sex_column)
