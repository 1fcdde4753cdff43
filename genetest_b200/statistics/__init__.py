"""Statistical tests available to a model specification
(``genetest/statistics/__init__.py:26-40``).  ``mixedlm`` is not on the B200 hot
path: it is fitted on the host, one model at a time (see ``models/mixedlm.py``
and ``analysis.execute_mixedlm_two_step``)."""

from .core import StatsError, StatsModels
from .models.linear import StatsLinear
from .models.logistic import StatsLogistic
from .models.survival import StatsCoxPH
from .models.mixedlm import StatsMixedLM

__all__ = ["available_models", "model_map", "StatsError", "StatsModels"]

available_models = {
    "linear": "Linear regression (ordinary least squares).",
    "logistic": "Logistic regression (GLM with binomial distribution).",
    "coxph": "Cox's proportional hazard regression (survival analysis).",
    "mixedlm": "Linear mixed effect model (random intercept).",
}

model_map = {
    "linear": StatsLinear,
    "logistic": StatsLogistic,
    "coxph": StatsCoxPH,
    "mixedlm": StatsMixedLM,
}
