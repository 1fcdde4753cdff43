"""Phenotype container over an existing DataFrame
(``genetest/phenotypes/dataframe.py``)."""

from .core import PhenotypesContainer

__all__ = ["DataFramePhenotypes"]


class DataFramePhenotypes(PhenotypesContainer):
    def __init__(self, dataframe, keep_sample_column=False,
                 repeated_measurements=False, sex_column=None):
        """``dataframe`` is indexed by sample id."""
        self._phenotypes = dataframe
        self._repeated = repeated_measurements
        self._sex_column = sex_column
        if not repeated_measurements and dataframe.index.duplicated().any():
            raise ValueError("duplicated samples (use repeated_measurements)")

    def close(self):
        pass

    def get_phenotypes(self, li=None):
        if li is None:
            return self._phenotypes.copy()
        for c in li:
            if c not in self._phenotypes.columns:
                raise KeyError(c)
        return self._phenotypes.loc[:, list(li)].copy()

    def get_nb_samples(self):
        return self._phenotypes.index.unique().shape[0]

    def get_nb_variables(self):
        return self._phenotypes.shape[1]

    def is_repeated(self):
        return self._repeated

    def keep_samples(self, keep):
        self._phenotypes = self._phenotypes[
            self._phenotypes.index.isin(set(keep))]
