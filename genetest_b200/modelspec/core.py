"""Model specification: variables, transformations and the design matrix.

Same public surface as ``genetest/modelspec/core.py`` (the formula / ModelSpec
API is kept as is, it is how users describe an analysis):

* ``phenotypes.<name>`` / ``genotypes.<name>`` hand out :class:`EntityIdentifier`
  placeholders (reference ``core.py:166-187``);
* ``factor``, ``log10``, ``ln``, ``pow``, ``interaction``, ``gwas_interaction``
  register transformations and return the identifier of their result
  (``core.py:190-231``, handlers ``:553-651``);
* :class:`ModelSpec` builds the analysis data frame
  (``create_data_matrix``, ``core.py:343-479``);
* ``result[...]`` builds accessors into a result dictionary (``core.py:47-85``).

The registries are process-wide, as in the reference; ``_reset()`` clears them.
"""

import functools
import itertools
import operator
import re
import uuid

import numpy as np
import pandas as pd

from ..statistics import model_map
from ..statistics.descriptive import get_maf

SNPs = "SNPs"
model = "MODEL"

_PHENO = "PHENOTYPES"
_GENO = "GENOTYPES"


class PheWAS(object):
    """Outcome marker for phenome-wide scans (``li`` = outcome entities)."""
    def __init__(self, li=None):
        self.li = li


# --------------------------------------------------------------------------
# identifiers and arithmetic expressions
# --------------------------------------------------------------------------
class _Arithmetic(object):
    """``a * b``, ``a + b`` ... on identifiers build an :class:`Expression`."""

    def _node(self):
        return self

    def _binary(self, op, other):
        return Expression((op, self._node(), other))

    def __mul__(self, other):
        return self._binary(operator.mul, other)

    def __truediv__(self, other):
        return self._binary(operator.truediv, other)

    def __add__(self, other):
        return self._binary(operator.add, other)

    def __sub__(self, other):
        return self._binary(operator.sub, other)

    __rmul__ = __mul__
    __radd__ = __add__
    __rtruediv__ = __truediv__
    __rsub__ = __sub__
    __div__ = __truediv__
    __rdiv__ = __truediv__


class EntityIdentifier(_Arithmetic):
    """Placeholder for a variable of the model.  ``id`` is a UUID4 unless one
    is given; ``bind`` attaches values so expressions can be evaluated."""

    def __init__(self, id=None):
        self.id = id if id else str(uuid.uuid4())
        self.values = None

    def bind(self, values):
        self.values = values

    def __repr__(self):
        return self.id.split("-")[0]


class Expression(_Arithmetic):
    def __init__(self, expression):
        self.expression = expression

    def _node(self):
        return self.expression

    def eval(self):
        return self._evaluate(self.expression)

    @classmethod
    def _evaluate(cls, node):
        op, left, right = node

        def value(x):
            if isinstance(x, tuple):
                return cls._evaluate(x)
            if isinstance(x, Expression):
                return cls._evaluate(x.expression)
            return x.values if hasattr(x, "values") else x

        return op(value(left), value(right))


# --------------------------------------------------------------------------
# registries
# --------------------------------------------------------------------------
class DependencyManager(object):
    """Remembers which variables of a source were asked for."""
    dependencies = {}

    def __init__(self, source):
        self.source = source

    def __getitem__(self, key):
        registry = DependencyManager.dependencies
        entity = registry.get((self.source, key))
        if entity is None:
            entity = registry[(self.source, key)] = EntityIdentifier()
        return entity

    __getattr__ = __getitem__


class TransformationManager(object):
    """Remembers which data manipulations the model needs, in order."""
    transformations = []

    def __init__(self, action):
        self.action = action

    def __call__(self, source, *params, name=None):
        if name is not None:
            target_id = "TRANSFORM:{}".format(name)
        elif isinstance(source, Expression):
            raise ValueError("The name parameter is mandatory for "
                             "expression-based transformations.")
        else:
            target_id = "TRANSFORM:{}:{}".format(self.action, source.id)
        for _action, _source, target, _params in \
                TransformationManager.transformations:
            if target.id == target_id:
                return target          # same transformation asked twice
        target = EntityIdentifier(target_id)
        TransformationManager.transformations.append(
            (self.action, source, target, params))
        return target


_HANDLERS = {}


def _handler(action):
    def register(f):
        _HANDLERS[action] = f
        return f
    return register


def _columns_of(data, entities):
    """For every entity, its columns in ``data`` (a factor has one per level)
    and the level suffix of each."""
    columns = [[c for c in data.columns if c.startswith(e.id)]
               for e in entities]
    levels = [[c[len(e.id) + 1:] for c in cols]
              for cols, e in zip(columns, entities)]
    return columns, levels


def _level_key(level_names):
    return re.sub(":{2,}", "", ":".join(level_names).strip(":"))


@_handler("LOG10")
def _log10(data, entity):
    return np.log10(data[entity.id])


@_handler("LN")
def _ln(data, entity):
    return np.log(data[entity.id])


@_handler("POW")
def _pow(data, entity, power):
    return np.power(data[entity.id], power)


@_handler("ENCODE_FACTOR")
def _encode_factor(data, entity):
    """0/1 indicator per level; the first sorted level is the reference,
    missing values stay missing."""
    values = data[entity.id]
    missing = values.isnull()
    if hasattr(values, "cat"):
        levels = sorted(values.cat.categories)
    else:
        levels = sorted(values[~missing].unique())
    out = {}
    for level in levels[1:]:
        col = (values == level).astype(float)
        col[missing] = np.nan
        out["level.{}".format(level)] = col
    return out


@_handler("INTERACTION")
def _interaction(data, *entities):
    columns, levels = _columns_of(data, entities)
    out = {}
    for cols, names in zip(itertools.product(*columns),
                           itertools.product(*levels)):
        product = functools.reduce(np.multiply, (data[c] for c in cols))
        key = _level_key(names)
        if key == "":
            return product             # no factor involved: a single column
        out[key] = product
    return out


@_handler("GWAS_INTERACTION")
def _gwas_interaction(data, *entities):
    """Only names the columns to multiply with the SNPs column later."""
    columns, levels = _columns_of(data, entities)
    return {_level_key(names): cols
            for cols, names in zip(itertools.product(*columns),
                                   itertools.product(*levels))}


# --------------------------------------------------------------------------
# result accessors
# --------------------------------------------------------------------------
class Result(object):
    """Path into a results dictionary, e.g. ``result["SNPs"]["coef"]``."""

    def __init__(self, entity):
        self.path = [entity]

    def __getitem__(self, key):
        self.path.append(key)
        return self

    def get(self, results):
        cur = results
        for field in self.path:
            if not isinstance(field, EntityIdentifier):
                cur = cur[field]
            elif field.id in results:
                cur = cur[field.id]
            else:
                # a factor: gather "<id>:<level>" entries by level
                per_level = {}
                for key in results:
                    prefix, _, level = key.rpartition(":")
                    if prefix == field.id:
                        per_level[level] = cur[key]
                cur = per_level
        return cur


class ResultMetaclass(object):
    def __getitem__(self, entity):
        return Result(entity)


# --------------------------------------------------------------------------
# the model specification
# --------------------------------------------------------------------------
def genotype_to_df(g, samples, as_string=False):
    """One marker as a one-column frame indexed by sample
    (``geneparse.utils.genotype_to_df``)."""
    name = g.variant.name if g.variant.name else "genotypes"
    df = pd.DataFrame(g.genotypes, index=samples, columns=[name])
    if as_string:
        df["alleles"] = None
        hard = df[name].round()
        df.loc[hard == 0, "alleles"] = "{0}/{0}".format(g.reference)
        df.loc[hard == 2, "alleles"] = "{0}/{0}".format(g.coded)
        df.loc[hard == 1, "alleles"] = "{0}/{1}".format(g.reference, g.coded)
        df = df[["alleles"]]
        df.columns = [name]
    return df


class ModelSpec(object):
    def __init__(self, outcome, predictors, test, no_intercept=False,
                 stratify_by=None):
        """
        Args:
            outcome: an EntityIdentifier, or a dict of labelled outcomes
                (``{"tte": ..., "event": ...}``), or a PheWAS marker.
            predictors (list): EntityIdentifiers and/or ``SNPs``.
            test: name of a statistical test, or a callable returning an
                instance of one.
            no_intercept (bool): do not add the column of ones.
            stratify_by (list): EntityIdentifiers to stratify on.
        """
        if isinstance(outcome, EntityIdentifier):
            outcome = {"y": outcome}
        self.outcome = outcome
        try:
            iter(predictors)
        except TypeError:
            raise TypeError("'predictors' argument needs to be an iterable.")
        self.predictors = predictors
        self.test = self._resolve_test(test)
        self.no_intercept = no_intercept
        if stratify_by is not None and not hasattr(stratify_by, "__iter__"):
            stratify_by = list(stratify_by)
        self.stratify_by = stratify_by
        # filled while the data matrix is built
        self.variant_metadata = {}
        self.has_gwas_interaction = False

    @staticmethod
    def _resolve_test(test):
        """A zero-argument factory of test instances."""
        try:
            return model_map[test]
        except (KeyError, TypeError):
            pass
        if callable(test):
            return test
        raise ValueError("{} is not a valid statistical model.".format(test))

    @property
    def dependencies(self):
        return DependencyManager.dependencies

    @property
    def transformations(self):
        return TransformationManager.transformations

    def get_tested_variants(self):
        if SNPs in self.predictors:
            return SNPs
        return [(entity, key[1]) for key, entity in self.dependencies.items()
                if key[0] == _GENO]

    def get_translations(self):
        """entity id -> variable name."""
        return {entity.id: key[1]
                for key, entity in self.dependencies.items()}

    # -- data matrix --------------------------------------------------------
    def create_data_matrix(self, phenotypes, genotypes):
        """Outcome(s), predictors (transformed) and the intercept, one row
        per phenotype sample.  In a GWAS the SNPs column is added later, per
        marker."""
        if isinstance(self.outcome, PheWAS) and self.outcome.li is None:
            df = phenotypes.get_phenotypes()
            self.outcome.li = []
            ids = []
            for col in df.columns:
                entity = self.dependencies.get((_PHENO, col))
                if entity is None:
                    entity = self.dependencies[(_PHENO, col)] = \
                        EntityIdentifier()
                if entity not in self.predictors:
                    self.outcome.li.append(entity)
                ids.append(entity.id)
            df.columns = ids
        else:
            names = [key[1] for key in self.dependencies if key[0] == _PHENO]
            df = phenotypes.get_phenotypes(names)
            df.columns = [self.dependencies[(_PHENO, nm)].id for nm in names]

        markers = self.get_tested_variants()
        if markers is not SNPs:
            for entity, marker in markers:
                df = self._join_marker(df, entity, marker, phenotypes,
                                       genotypes)

        df = self._apply_transformations(df)
        keep = self._filter_columns(df)
        if not self.no_intercept:
            keep.append("intercept")
            df["intercept"] = 1
        return df[keep]

    def _join_marker(self, df, entity, marker, phenotypes, genotypes):
        """Explicitly named variant: becomes a regular column, oriented on
        its minor allele (reference ``core.py:392-468``)."""
        try:
            found = genotypes.get_variant_by_name(marker)
        except (KeyError, ValueError):
            found = []
        if len(found) == 0:
            raise ValueError("Could not find '{}' in genotypes container."
                             .format(marker))
        if len(found) != 1:
            raise ValueError("{}: invalid marker name".format(marker))
        g = found[0]
        stratifier = bool(self.stratify_by) and entity in self.stratify_by
        cur = genotype_to_df(g, genotypes.get_samples(), as_string=stratifier)
        cur.columns = [entity.id]
        df = df.join(cur, how="inner")
        if df.shape[0] == 0:
            raise ValueError(
                "No sample left after joining. Perhaps the sample IDs in the "
                "genotypes and phenotypes containers are different.")
        if stratifier:
            return df
        values = df[entity.id]
        if phenotypes.is_repeated():
            values = values.loc[~values.index.duplicated(keep="first")]
        maf, minor, major, flip = get_maf(values, g.coded, g.reference)
        if flip:
            df[entity.id] = 2 - df[entity.id]
        self.variant_metadata[entity.id] = {
            "name": marker, "chrom": str(g.variant.chrom),
            "pos": g.variant.pos, "minor": minor, "major": major, "maf": maf,
            "coded": minor,
        }
        return df

    def _apply_transformations(self, df):
        for action, source, target, params in self.transformations:
            res = _HANDLERS[action](df, source, *params)
            if action == "GWAS_INTERACTION":
                # nothing to compute yet: {output column: columns to multiply
                # with the SNPs column}
                self.has_gwas_interaction = True
                self.gwas_interaction = {}
                for key, cols in res.items():
                    out = target.id if key == "" else "{}:{}".format(
                        target.id, key)
                    self.gwas_interaction[out] = cols
            elif isinstance(res, dict):
                for key, col in res.items():
                    df["{}:{}".format(target.id, key)] = col
            else:
                df[target.id] = res
        return df

    def _filter_columns(self, df):
        if isinstance(self.outcome, PheWAS):
            keep = [v.id for v in self.outcome.li if v not in self.predictors]
        else:
            keep = [v.id for v in self.outcome.values()]
        for pred in self.predictors:
            if pred is SNPs:
                continue
            if not isinstance(pred, EntityIdentifier):
                raise ValueError(
                    "Predictors are expected to be entity identifiers (and "
                    "'{}' is of type {}).".format(pred, type(pred)))
            # a factor owns one column per level
            keep.extend(c for c in df.columns if c.startswith(pred.id))
        if self.stratify_by is not None:
            for col in self.stratify_by:
                if not isinstance(col, EntityIdentifier):
                    raise ValueError(
                        "Statification variables are expected to be entity "
                        "identifiers (and '{}' is of type {})."
                        .format(col, type(col)))
                keep.append(col.id)
        return keep


def _reset():
    TransformationManager.transformations = []
    DependencyManager.dependencies = {}


result = ResultMetaclass()

phenotypes = DependencyManager(_PHENO)
genotypes = DependencyManager(_GENO)

factor = TransformationManager("ENCODE_FACTOR")
log10 = TransformationManager("LOG10")
ln = TransformationManager("LN")
pow = TransformationManager("POW")
interaction = TransformationManager("INTERACTION")
gwas_interaction = TransformationManager("GWAS_INTERACTION")
