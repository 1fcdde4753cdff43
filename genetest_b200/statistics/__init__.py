"""Statistical tests available to a model specification
(``genetest/statistics/__init__.py:26-40``).  ``mixedlm`` is outside the
B200 hot path and is not provided."""

from .core import StatsError, StatsModels
from .models.linear import StatsLinear
from .models.logistic import StatsLogistic
from .models.survival import StatsCoxPH

__all__ = ["available_models", "model_map", "StatsError", "StatsModels"]

available_models = {
    "linear": "Linear regression (ordinary least squares).",
    "logistic": "Logistic regression (GLM with binomial distribution).",
    "coxph": "Cox's proportional hazard regression (survival analysis).",
}

model_map = {
    "linear": StatsLinear,
    "logistic": StatsLogistic,
    "coxph": StatsCoxPH,
}
