"""Delimited text file of phenotypes (``genetest/phenotypes/text.py:21-120``):
same constructor arguments, sample ids kept as strings, duplicated samples
only accepted with ``repeated_measurements=True``."""

import pandas as pd

from .dataframe import DataFramePhenotypes

__all__ = ["TextPhenotypes"]


class TextPhenotypes(DataFramePhenotypes):
    def __init__(self, filename, sample_column="sample", field_separator="\t",
                 missing_values=None, repeated_measurements=False,
                 keep_sample_column=False, sex_column=None):
        df = pd.read_csv(filename, sep=field_separator,
                         na_values=missing_values,
                         dtype={sample_column: str})
        df = df.set_index(sample_column, drop=not keep_sample_column)
        if not repeated_measurements and not df.index.is_unique:
            raise ValueError("Index has duplicate keys: {}".format(
                sorted(set(df.index[df.index.duplicated()]))))
        df.index.name = "sample_id"
        super().__init__(df, repeated_measurements=repeated_measurements,
                         sex_column=sex_column)

    def __repr__(self):
        return "TextPhenotypes({:,d} samples, {:,d} variables)".format(
            self.get_nb_samples(), self.get_nb_variables())

    def merge(self, other):
        """Join the variables of another container (on the sample ids)."""
        self._phenotypes = self._phenotypes.join(other._phenotypes)
