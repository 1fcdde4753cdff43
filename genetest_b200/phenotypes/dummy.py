"""In-memory phenotype container used by tests
(``genetest/phenotypes/dummy.py``): assign ``data`` freely."""

import numpy as np
import pandas as pd

from .core import PhenotypesContainer

__all__ = ["_DummyPhenotypes"]


class _DummyPhenotypes(PhenotypesContainer):
    def __init__(self):
        self.data = pd.DataFrame(
            {"y": [0, 1, 0, np.nan, 0], "V1": [1, 2, 3, 4, 5],
             "V2": [0, np.nan, 1, 0, 1], "V3": [-2, 1, np.nan, 0, -1],
             "V4": [0.1, 0.2, 0.3, 0.4, 0.5], "V5": [0, 0, 0, 1, 1]},
            index=["s{}".format(i) for i in range(1, 6)])
        self._repeated = False

    def close(self):
        pass

    def get_phenotypes(self, li=None):
        cols = list(self.data.columns if li is None else li)
        for c in cols:
            if c not in self.data.columns:
                raise KeyError(c)
        return self.data.loc[:, cols]

    def get_nb_samples(self):
        return self.data.shape[0]

    def get_nb_variables(self):
        return self.data.shape[1]

    def is_repeated(self):
        return self._repeated

    def keep_samples(self, keep):
        self.data = self.data.loc[keep, :]
