"""Plugin base class and the one recoverable error of the statistics layer.

Mirrors the contract of ``genetest/statistics/core.py:36-61``: a model object
exposes ``fit(y, X) -> dict``; a statistical problem with one marker raises
:class:`StatsError`, whose ``str()`` is the reason written to
``<prefix>_failed_snps.txt`` (``genetest/analysis.py:177-181``).
"""

__all__ = ["StatsModels", "StatsError", "BatchedGWAS"]


class StatsError(Exception):
    """A statistical problem with one test (not a program error)."""

    def __init__(self, msg):
        super().__init__(str(msg))
        self.message = str(msg)

    def __str__(self):
        return self.message


class StatsModels(object):
    """Interface of a statistical test."""

    def fit(self, y, X):
        """Fit the model.

        Args:
            y (pandas.DataFrame): endogenous variable(s).
            X (pandas.DataFrame): exogenous variables (design matrix).

        Returns:
            dict: ``{"MODEL": {...}, <column>: {...}, ...}``.
        """
        raise NotImplementedError()

    def __repr__(self):
        return self.__class__.__name__


class BatchedGWAS(object):
    """Capability marker: the test can run a whole batch of markers on the
    B200 engine.  ``analysis._execute_gwas`` looks for ``gwas_model`` (the
    engine's model name) and ``gwas_options()`` (engine thresholds)."""

    gwas_model = None

    def gwas_options(self):
        return {}
