"""Phenotype containers (``genetest/phenotypes/core.py``): sample-indexed
tables the model specification pulls its variables from."""

__all__ = ["PhenotypesContainer"]


class PhenotypesContainer(object):
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        raise NotImplementedError()

    def get_phenotypes(self, li=None):
        """DataFrame of the requested variables (all when ``li`` is None),
        indexed by sample id."""
        raise NotImplementedError()

    def get_nb_samples(self):
        raise NotImplementedError()

    def get_nb_variables(self):
        raise NotImplementedError()

    def is_repeated(self):
        """True when samples have several rows (repeated measurements)."""
        raise NotImplementedError()

    def keep_samples(self, keep):
        raise NotImplementedError()

    def get_sex(self):
        """Series of sex per sample (0 = female, 1 = male, NaN unknown)."""
        column = getattr(self, "_sex_column", None)
        if column is None:
            raise ValueError("No sex column was specified in Phenotype "
                             "container")
        sex = self._phenotypes.loc[:, column]
        if self._repeated:
            groups = sex.groupby(sex.index)
            bad = sorted(s for s, v in groups if len(v.unique()) != 1)
            if bad:
                raise ValueError("Samples with duplicated different sex: {}"
                                 .format(",".join(bad)))
            sex = groups.first()
        if set(sex.dropna().unique()) - {0.0, 1.0}:
            raise ValueError("Sex should only contain 0 and 1 (and maybe NaN)")
        return sex
