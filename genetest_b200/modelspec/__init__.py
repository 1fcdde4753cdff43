"""Model specification API (``genetest/modelspec/__init__.py``)."""

from . import core
from .core import ModelSpec
from ..statistics import model_map
from .grammar import parse_formula, parse_modelspec

__all__ = [
    "ModelSpec", "parse_modelspec", "parse_formula", "result", "phenotypes",
    "genotypes", "factor", "log10", "ln", "pow", "interaction",
    "gwas_interaction", "SNPs", "_reset", "PheWAS", "modelspec_from_formula",
]

result = core.result
phenotypes = core.phenotypes
genotypes = core.genotypes
factor = core.factor
log10 = core.log10
ln = core.ln
pow = core.pow
interaction = core.interaction
gwas_interaction = core.gwas_interaction
SNPs = core.SNPs
PheWAS = core.PheWAS
_reset = core._reset


def modelspec_from_formula(formula, test, test_kwargs):
    """(ModelSpec, subgroups) for a formula and a test name or class."""
    parsed = dict(parse_formula(formula))
    kwargs = test_kwargs or {}
    if callable(test):
        parsed["test"] = lambda: test(**kwargs)
    else:
        parsed["test"] = lambda: model_map[test](**kwargs)
    conditions = parsed.pop("conditions")
    subgroups = None
    if conditions is not None:
        parsed["stratify_by"] = [c["name"] for c in conditions]
        subgroups = [c["level"] for c in conditions]
    return ModelSpec(**parsed), subgroups
