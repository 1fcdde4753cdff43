"""Run a statistical analysis (``genetest/analysis.py``).

``execute`` keeps the reference's entry point and call protocol
(``analysis.py:230-325``); the GWAS branch (``_execute_gwas``, reference
``analysis.py:593-728`` + ``_gwas_worker`` ``:68-205``) no longer spawns worker
processes that refit statsmodels one marker at a time: markers are streamed
in batches to the B200 engine (``genetest_b200.engine``), which applies the
complete-case mask, MAF, flip, interaction columns and the fit for every
marker of the batch, and comes back with fixed-size result blocks that are
turned into the reference's per-marker dictionaries for the subscribers.

Under ``torchrun`` (one process per GPU) every rank analyses a contiguous
range of the markers and rank 0 gathers the result blocks (NCCL) and drives
the subscribers -- there is no other communication.
"""

import itertools
import os
import logging

import numpy as np
import pandas as pd

from . import sharding
from . import subscribers as subscribers_module
from .engine import get_engine
from .modelspec import PheWAS, SNPs, modelspec_from_formula
from .statistics.core import StatsError
from .statistics.descriptive import get_maf, get_sex_maf
from .statistics.models._single import failure_reason
from .statistics.models.survival import cox_param_dict

logger = logging.getLogger(__name__)

PARAM_FIELDS = ("coef", "std_err", "lower_ci", "upper_ci", "t_value",
                "p_value")

#: markers per engine call (bounds host staging memory)
# markers per call of the host pipeline: two waves of the tcgen05 contraction
# (148 CTAs x 512 SNP rows); the engine splits it into half-wave stages
BATCH_SNPS = 151552


def _missing(y, X):
    """Boolean vector of the rows without any missing value."""
    return ~(y.isnull().any(axis=1) | X.isnull().any(axis=1))


def _generate_sample_order(x_samples, geno_samples):
    """Positions (in the genotype container's sample order) of the samples
    that are rows of X, and their labels (``analysis.py:43-65``)."""
    in_x = set(x_samples)
    geno_order = np.array(
        [i for i, s in enumerate(geno_samples) if s in in_x], dtype=int)
    x_order = np.asarray(geno_samples, dtype=object)[geno_order]
    x_index = pd.Index(x_samples)
    if x_index.duplicated().any():
        # repeated measurements (mixedlm): one genotype per row of X, in the
        # order X.loc[x_order] lists them (analysis.py:54-63)
        counts = x_index.value_counts()
        geno_order = np.array(
            [i for i, s in zip(geno_order, x_order) for _ in range(counts[s])],
            dtype=int)
    return geno_order, x_order


def execute_formula(phenotypes, genotypes, formula, test, test_kwargs=None,
                    subscribers=None, variant_predicates=None,
                    output_prefix=None, maf_t=None, cpus=None,
                    sexual_chromosome=False):
    modelspec, subgroups = modelspec_from_formula(formula, test, test_kwargs)
    execute(phenotypes, genotypes, modelspec, subscribers, variant_predicates,
            output_prefix, subgroups, maf_t, cpus, sexual_chromosome)


def execute_mixedlm_two_step(phenotypes, genotypes, formula, p_threshold,
                             output_prefix, test_kwargs=None,
                             variant_predicates=None, maf_t=None, cpus=None,
                             sexual_chromosome=False):
    """The "optimized" MixedLM analysis of the reference's command line
    (``genetest/scripts/cli.py:183-310``), as a library call:

    1. fit the mixed model once WITHOUT the SNPs (host) and keep the predicted
       random effect of every sample;
    2. run a plain linear GWAS ``_random_effects ~ SNPs`` over all the markers
       -- the B200 hot path -- into ``<prefix>.mixedlm_approximation.txt``;
    3. refit the real mixed model (host, one marker at a time) only for the
       markers whose approximate p-value is below ``p_threshold``, into
       ``<prefix>.txt``.

    Returns the set of marker names that went through step 3."""
    from . import modelspec as spec
    from .modelspec import result as analysis_results
    from .modelspec.predicates import NameFilter
    from .phenotypes import DataFramePhenotypes
    from .subscribers import GWASWriter, ResultsMemory, RowWriter

    modelspec, subgroups = modelspec_from_formula(formula, "mixedlm", test_kwargs)
    if subgroups is not None:
        raise ValueError("Subgroups are not available for MixedLM optimization")
    if "SNPs" in modelspec.predictors:
        del modelspec.predictors[modelspec.predictors.index("SNPs")]

    logger.info("Computing the random effects")
    memory = ResultsMemory()
    execute(phenotypes, genotypes, modelspec, subscribers=[memory],
            output_prefix=output_prefix, subgroups=subgroups, maf_t=maf_t,
            cpus=cpus, sexual_chromosome=sexual_chromosome)
    random_effects = memory.results.pop()["MODEL"]["random_effects"]
    group_col = modelspec.get_translations()[modelspec.outcome["groups"].id]

    old = phenotypes.get_phenotypes()
    assert "_random_effects" not in old.columns
    duplicated = old.index.duplicated(keep="first")
    new = pd.merge(
        left=old.loc[~duplicated, :].copy(),
        right=pd.DataFrame({"_random_effects": random_effects}),
        left_on=group_col, right_index=True)

    spec._reset()
    new_phenotypes = DataFramePhenotypes(new)
    assert not new_phenotypes.is_repeated()
    approx_spec = spec.ModelSpec(outcome=spec.phenotypes._random_effects,
                                 predictors=[spec.SNPs], test="linear")
    approximation_fn = output_prefix + ".mixedlm_approximation.txt"
    r = analysis_results
    writer = RowWriter(
        filename=approximation_fn,
        columns=[("snp", r["SNPs"]["name"]), ("chr", r["SNPs"]["chrom"]),
                 ("pos", r["SNPs"]["pos"]), ("major", r["SNPs"]["major"]),
                 ("minor", r["SNPs"]["minor"]), ("maf", r["SNPs"]["maf"]),
                 ("n", r["MODEL"]["nobs"]), ("p", r["SNPs"]["p_value"]),
                 ("ll", r["MODEL"]["log_likelihood"]),
                 ("adj_r2", r["MODEL"]["r_squared_adj"]),
                 ("approximation", "Yes")],
        header=True, append=False)
    logger.info("Executing the optimized MixedLM")
    execute(new_phenotypes, genotypes, approx_spec, subscribers=[writer],
            output_prefix=output_prefix + ".mixedlm_approximation",
            maf_t=maf_t, cpus=cpus, variant_predicates=variant_predicates,
            sexual_chromosome=sexual_chromosome)
    if not os.path.isfile(approximation_fn):
        raise RuntimeError("{}: no such file".format(approximation_fn))
    approximation = pd.read_csv(approximation_fn, sep="\t")
    markers = set(approximation.loc[approximation.p < p_threshold, "snp"].values)
    if len(markers) == 0:
        logger.info("No marker had a p value < than {}".format(p_threshold))
        return markers

    spec._reset()
    logger.info("Executing MixedLM on {:,d} markers".format(len(markers)))
    execute_formula(phenotypes, genotypes, formula, "mixedlm",
                    test_kwargs=test_kwargs,
                    subscribers=[GWASWriter(output_prefix + ".txt", "mixedlm")],
                    variant_predicates=[NameFilter(extract=markers)],
                    output_prefix=output_prefix, maf_t=maf_t, cpus=cpus,
                    sexual_chromosome=sexual_chromosome)
    return markers


def execute(phenotypes, genotypes, modelspec, subscribers=None,
            variant_predicates=None, output_prefix=None, subgroups=None,
            maf_t=None, cpus=None, sexual_chromosome=False):
    """Execute an analysis.

    Args:
        phenotypes: the phenotypes container.
        genotypes: the genotypes container (geneparse-like reader).
        modelspec (ModelSpec): the model specification.
        subscribers (list): result sinks (default: ``Print()``).
        variant_predicates (list): callables ``f(snp) -> bool``.
        output_prefix (str): prefix of the failed / skipped marker files.
        subgroups (list): levels of the stratification variables.
        maf_t (float): MAF threshold.
        cpus (int): kept for compatibility; the GPU engine ignores it.
        sexual_chromosome (bool): sex-aware MAF.
    """
    # one process per GPU under torchrun: the ranks share the markers, the
    # result sinks live on rank 0 only (every rank would otherwise re-open and
    # re-write the same output files)
    sharding.ensure_initialized()
    if subscribers is None:
        subscribers = [subscribers_module.Print()]
    if not sharding.is_primary():
        subscribers = []
    for subscriber in subscribers:
        subscriber.init(modelspec)
    if variant_predicates is None:
        variant_predicates = []

    if isinstance(modelspec.outcome, PheWAS):
        return _execute_phewas(phenotypes, genotypes, modelspec, subscribers,
                               variant_predicates, output_prefix, subgroups)

    data = modelspec.create_data_matrix(phenotypes, genotypes)
    data = data.dropna()            # missing outcome or covariable

    y_cols = tuple(modelspec.outcome.keys())
    y = data[[modelspec.outcome[col].id for col in y_cols]]
    X = data.drop(y.columns, axis=1)
    y.columns = y_cols

    bad_cols = _get_uninformative_factors(X, modelspec)
    if len(bad_cols):
        logger.info("After removing missing values, dropping ({}) factor "
                    "levels that have no variation.".format(len(bad_cols)))
        X = X.drop(bad_cols, axis=1)

    messages = {"skipped": [], "failed": []}

    sample_sex = None
    if sexual_chromosome:
        logger.info("Analysis is performed on a sexual chromosome")
        sample_sex = phenotypes.get_sex()

    if modelspec.stratify_by is not None:
        _execute_stratified(genotypes, modelspec, subscribers, y, X,
                            variant_predicates, subgroups, messages, maf_t,
                            cpus, sample_sex)
    elif SNPs in modelspec.predictors:
        _execute_gwas(genotypes, modelspec, subscribers, y, X,
                      variant_predicates, messages, maf_t, cpus, sample_sex)
    elif sharding.is_primary():
        # a single fit: nothing to share out
        _execute_simple(modelspec, subscribers, y, X, variant_predicates,
                        messages)

    for subscriber in subscribers:
        subscriber.close()

    if not sharding.is_primary():
        return
    prefix = (output_prefix + "_") if output_prefix is not None else ""
    if messages["failed"]:
        with open(prefix + "failed_snps.txt", "w") as f:
            for name, reason in messages["failed"]:
                f.write("{}\t{}\n".format(name, reason))
    if messages["skipped"]:
        with open(prefix + "not_analyzed_snps.txt", "w") as f:
            for name in messages["skipped"]:
                f.write("{}\n".format(name))


def _get_uninformative_factors(df, modelspec):
    """Factor-level columns that are all zero (no sample has the level)."""
    factor_ids = [t[2].id for t in modelspec.transformations
                  if t[0] == "ENCODE_FACTOR"]
    zero = df.columns[df.sum() == 0]
    return [c for c in zero if any(c.startswith(f) for f in factor_ids)]


def _execute_simple(modelspec, subscribers, y, X, variant_predicates,
                    messages):
    """A single fit (no marker loop)."""
    if len(variant_predicates) != 0:
        logger.warning("Variant predicates are only used for GWAS analyses.")
    test = modelspec.test()
    logger.info("Executing {} - Design matrix has shape: {}".format(
        test, X.shape))
    results = test.fit(y, X)
    for entity in results:
        if entity in modelspec.variant_metadata:
            results[entity].update(modelspec.variant_metadata[entity])
    for subscriber in subscribers:
        try:
            subscriber.handle(results)
        except KeyError as e:
            return subscribers_module.subscriber_error(e.args[0])


def _execute_stratified(genotypes, modelspec, subscribers, y, X,
                        variant_predicates, subgroups, messages, maf_t=None,
                        cpus=None, sample_sex=None):
    """One analysis per combination of the stratification levels."""
    assert len(modelspec.stratify_by) == len(subgroups)
    var_levels = []
    for var, subgroup in zip(modelspec.stratify_by, subgroups):
        if subgroup is None:
            var_levels.append(X[var.id].dropna().unique())
        elif hasattr(subgroup, "__iter__") and type(subgroup) is not str:
            var_levels.append(subgroup)
        else:
            var_levels.append([subgroup])

    gwas_mode = SNPs in modelspec.predictors
    translations = modelspec.get_translations()
    strat_ids = [v.id for v in modelspec.stratify_by]

    for levels in itertools.product(*var_levels):
        current = list(zip(modelspec.stratify_by, levels))
        idx = np.logical_and.reduce(
            [X[var.id] == level for var, level in current])
        subset_info = {translations[var.id]: level for var, level in current}
        for sub in subscribers:
            sub._update_current_subset(subset_info)

        this_x = X.loc[idx, :]
        bad_cols = _get_uninformative_factors(this_x, modelspec)
        this_x = this_x.drop(bad_cols, axis=1).drop(strat_ids, axis=1)
        logger.info("Analysing subgroup {}".format(";".join(
            "{}:{}".format(*kv) for kv in sorted(subset_info.items()))))

        if this_x.shape[0] == 0:
            observed = {translations[var.id]: list(X[var.id].dropna().unique())
                        for var in modelspec.stratify_by}
            raise ValueError(
                "No samples left in subgroup analysis ({}). Are all requested "
                "levels valid?\nThe observed levels are: {}".format(
                    subset_info, observed))
        if this_x.shape[1] == 0:
            raise ValueError(
                "No columns left in subgroup analysis ({}). Maybe there are "
                "no samples with non-null values in the requested subgroup."
                .format(subset_info))
        if len(bad_cols):
            logger.info("After stratification, dropping factor levels ({}) "
                        "that have no variation ".format(len(bad_cols)))

        if gwas_mode:
            _execute_gwas(genotypes, modelspec, subscribers, y.loc[idx, :],
                          this_x, variant_predicates, messages, maf_t, cpus,
                          sample_sex)
        else:
            _execute_simple(modelspec, subscribers, y.loc[idx, :], this_x,
                            variant_predicates, messages)


# --------------------------------------------------------------------------
# the GWAS branch
# --------------------------------------------------------------------------
class ResultBlock(object):
    """Result blocks of a batch of markers plus their metadata; turns rows
    into the reference's result dictionaries on demand."""

    def __init__(self, model, columns, res, meta, maf_t):
        self.model = model
        self.columns = columns          # design columns in fit() order
        self.res = res                  # engine.Results
        self.meta = meta                # [(name, chrom, pos, coded, reference)]
        self.maf_t = maf_t

    def __len__(self):
        return len(self.meta)

    def ok(self, i):
        return self.res.status[i] == 0

    @property
    def ok_mask(self):
        return (self.res.status == 0).astype(np.uint8)

    def column(self, path, subset_info=None):
        """Values of one result path for every marker of the block: a float
        or integer ndarray, a list of strings, or one string.  ``path`` is a
        ``modelspec.result[...]`` path such as ``["SNPs", "coef"]``.  Raises
        KeyError for a field the test does not produce (as a dict lookup
        would)."""
        from .statistics.models.survival import COX_FIELDS
        entity, field = [getattr(x, "id", x) for x in path]
        res = self.res
        if entity == "MODEL":
            if field == "nobs":
                return (res.nobs.astype(np.int64) if self.model == "coxph"
                        else res.nobs.astype(np.float64))
            if field == "log_likelihood":
                return res.llf
            if field == "r_squared_adj" and self.model == "linear":
                return res.stat
            if field == "nevents" and self.model != "linear":
                return res.stat
            if field == "subset_info" and subset_info is not None:
                return subset_info
            raise KeyError(field)
        if entity == "SNPs" and field in ("name", "chrom", "pos", "major",
                                          "minor", "maf"):
            flip = res.flip.astype(bool)
            if field == "name":
                return [str(m[0]) for m in self.meta]
            if field == "chrom":
                return [str(m[1]) for m in self.meta]
            if field == "pos":
                return [str(m[2]) for m in self.meta]
            if field == "maf":
                return res.maf
            coded = [m[3] for m in self.meta]
            ref = [m[4] for m in self.meta]
            if field == "minor":
                return [str(r if f else c) for c, r, f in zip(coded, ref, flip)]
            return [str(c if f else r) for c, r, f in zip(coded, ref, flip)]
        if entity not in self.columns:
            raise KeyError(entity)
        vals = res.param(self.columns.index(entity))
        if self.model == "coxph":
            if field not in COX_FIELDS:
                raise KeyError(field)
            coef, se, lo, hi, z, p = vals
            return {"coef": coef, "std_err": se, "hr": np.exp(coef),
                    "hr_lower_ci": np.exp(lo), "hr_upper_ci": np.exp(hi),
                    "z_value": z, "p_value": p}[field]
        if field not in PARAM_FIELDS:
            raise KeyError(field)
        return vals[PARAM_FIELDS.index(field)]

    def reason(self, i):
        return failure_reason(self.res, i, self.maf_t)

    def to_dict(self, i):
        """The dictionary ``Stats*.fit`` + ``_gwas_worker`` build for one
        marker (``linear.py:64-89`` / ``logistic.py:49-72`` /
        ``survival.py:100-120``, then ``analysis.py:191-195``)."""
        res = self.res
        name, chrom, pos, coded, reference = self.meta[i]
        if self.model == "linear":
            out = {"MODEL": {"r_squared_adj": float(res.stat[i]),
                             "log_likelihood": float(res.llf[i]),
                             "nobs": float(res.nobs[i])}}
        elif self.model == "logistic":
            out = {"MODEL": {"log_likelihood": float(res.llf[i]),
                             "nobs": float(res.nobs[i]),
                             "nevents": res.stat[i]}}
        else:
            out = {"MODEL": {"log_likelihood": float(res.llf[i]),
                             "nobs": int(res.nobs[i]),
                             "nevents": res.stat[i]}}
        for j, col in enumerate(self.columns):
            vals = res.param(j)[:, i]
            if self.model == "coxph":
                out[col] = cox_param_dict(vals)
            else:
                out[col] = {f: float(v) for f, v in zip(PARAM_FIELDS, vals)}
        flip = bool(res.flip[i])
        out["SNPs"].update({
            "chrom": str(chrom), "pos": pos,
            "major": coded if flip else reference,
            "minor": reference if flip else coded, "name": name,
        })
        out["SNPs"]["maf"] = float(res.maf[i])
        return out


def _takes_blocks(subscriber):
    """Whole result blocks go to a sink only when its ``handle`` -- the
    reference's one extension point (subscribers.py:49) -- is a stock
    implementation of this package: a user subclass that overrides ``handle``
    keeps receiving one dictionary per marker."""
    cls = type(subscriber)
    if not hasattr(cls, "handle_block"):
        return False
    for base in cls.__mro__:
        if "handle" in base.__dict__:
            return base.__module__ == subscribers_module.__name__
    return False


def _dispatch_block(block, subscribers, messages):
    """Failures -> ``messages``; successes -> the subscribers: whole blocks
    for sinks that define ``handle_block``, one dictionary per marker for the
    others (the reference's protocol, ``analysis.py:659-664``)."""
    for i in np.flatnonzero(block.res.status != 0):
        reason = block.reason(i)
        name = block.meta[i][0]
        logger.warning("{}: {}".format(name, reason))
        if name:
            messages["failed"].append((name, reason))
    dict_sinks = []
    for subscriber in subscribers:
        if _takes_blocks(subscriber):
            try:
                subscriber.handle_block(block)
            except KeyError as e:
                subscribers_module.subscriber_error(e.args[0])
        else:
            dict_sinks.append(subscriber)
    if not dict_sinks:
        return
    for i in np.flatnonzero(block.res.status == 0):
        results = block.to_dict(i)
        for subscriber in dict_sinks:
            try:
                subscriber.handle(results)
            except KeyError as e:
                subscribers_module.subscriber_error(e.args[0])


def _execute_phewas(phenotypes, genotypes, modelspec, subscribers,
                    variant_predicates, output_prefix, subgroups):
    """Phenome-wide scan (``analysis.py:328-452``): the same predictors against
    every phenotype that is not one of them, one fit each (a batch of one on
    the engine for the GPU models) -- not a per-SNP loop, so no batching; the
    reference's worker processes are replaced by a plain loop."""
    if subgroups:
        raise NotImplementedError(
            "Stratified analyses for pheWAS are not supported yet.")
    if variant_predicates:
        raise ValueError(
            "Variant predicates can only be used in the context of GWAS "
            "analyses.")
    data = modelspec.create_data_matrix(phenotypes, genotypes)
    predictors = [i.id for i in modelspec.predictors]
    if "intercept" in data.columns:
        predictors.append("intercept")
    X = data[predictors]
    fit = modelspec.test().fit
    translations = modelspec.get_translations()
    for entity in modelspec.outcome.li:
        if entity in modelspec.predictors:
            continue
        y_data = data[[entity.id]]
        not_missing = _missing(y_data, X)
        try:
            results = fit(y_data[not_missing], X[not_missing])
        except Exception as e:           # the reference skips the phenotype
            logger.debug("Exception raised during fitting: {}".format(e))
            continue
        results["outcome"] = translations[entity.id]
        for subscriber in subscribers:
            try:
                subscriber.handle(results)
            except KeyError as e:
                subscribers_module.subscriber_error(e.args[0])
    for subscriber in subscribers:
        subscriber.close()
    logger.info("PheWAS complete.")


def _host_fit_loop(genotypes, test, subscribers, y, X, variant_predicates,
                   messages, maf_t, gwas_interaction, sample_sex):
    """Plugin path for user-supplied tests without a batched kernel: the
    reference's loop body, one ``fit`` call per marker, in this process."""
    samples = genotypes.get_samples()
    geno_index, sample_order = _generate_sample_order(X.index, samples)
    if geno_index.shape[0] == 0:
        raise RuntimeError("Genotype and phenotype data have "
                           "non-overlapping indices.")
    X = X.copy()
    unique_samples = ~X.index.duplicated(keep="first")
    has_duplicates = not unique_samples.all()
    for snp in genotypes.iter_genotypes():
        try:
            if not all([f(snp) for f in variant_predicates]):
                messages["skipped"].append(snp.variant.name)
                continue
        except StopIteration:
            break
        X["SNPs"] = np.nan
        if gwas_interaction:
            X = X.drop(list(gwas_interaction.keys()), axis=1, errors="ignore")
        X.loc[sample_order, "SNPs"] = snp.genotypes[geno_index]
        not_missing = _missing(y, X)
        if np.sum(not_missing) == 0:
            messages["failed"].append((snp.variant.name,
                                       "All genotypes are missing"))
            continue
        # repeated measurements: every sample counts once in the MAF
        # (analysis.py:74-79, 127-129)
        for_maf = not_missing
        if has_duplicates:
            for_maf = not_missing & unique_samples
        if sample_sex is None:
            maf, minor, major, flip = get_maf(
                X.loc[for_maf, "SNPs"], snp.coded, snp.reference)
        else:
            ids = X.index[for_maf]
            maf, minor, major, flip = get_sex_maf(
                X.loc[ids, "SNPs"], sample_sex.reindex(ids), snp.coded,
                snp.reference)
        if maf_t is not None and maf < maf_t:
            messages["failed"].append(
                (snp.variant.name, "MAF: {} < {}".format(maf, maf_t)))
            continue
        if flip:
            X["SNPs"] = 2 - X["SNPs"]
        if gwas_interaction:
            for key, cols in gwas_interaction.items():
                col = X["SNPs"].copy()
                for c in cols:
                    col = col * X[c]
                X[key] = col
        try:
            results = test.fit(y[not_missing], X[not_missing])
        except StatsError as e:
            logger.warning("{}: {}".format(snp.variant.name, e))
            if snp.variant.name:
                messages["failed"].append((snp.variant.name, str(e)))
            continue
        results["SNPs"].update({
            "chrom": str(snp.variant.chrom), "pos": snp.variant.pos,
            "major": major, "minor": minor, "name": snp.variant.name,
        })
        results["SNPs"]["maf"] = maf
        for subscriber in subscribers:
            try:
                subscriber.handle(results)
            except KeyError as e:
                subscribers_module.subscriber_error(e.args[0])


def _integer_valued(block):
    ok = ~np.isnan(block)
    v = block[ok]
    return bool(np.all((v == 0) | (v == 1) | (v == 2)))


def _execute_gwas(genotypes, modelspec, subscribers, y, X, variant_predicates,
                  messages, maf_t=None, cpus=None, sample_sex=None):
    gwas_interaction = None
    if modelspec.has_gwas_interaction:
        gwas_interaction = modelspec.gwas_interaction
        for subscriber in subscribers:
            subscriber._update_gwas_interaction(
                list(sorted(gwas_interaction.keys())))

    test = modelspec.test()
    model = getattr(test, "gwas_model", None)
    if model is None:
        logger.info("{} has no batched kernel: fitting on the host, one "
                    "marker at a time".format(test))
        return _host_fit_loop(genotypes, test, subscribers, y, X,
                              variant_predicates, messages, maf_t,
                              gwas_interaction, sample_sex)

    # ---- sample alignment (analysis.py:43-65, 97-116) ----
    samples = list(genotypes.get_samples())
    if X.index.duplicated().any():
        raise NotImplementedError(
            "repeated measurements (duplicated samples) are the input of "
            "mixedlm; the batched {} kernels fit one row per sample".format(model))
    geno_index, sample_order = _generate_sample_order(X.index, samples)
    if geno_index.shape[0] == 0:
        logger.critical("Genotype and phenotype data have non-overlapping "
                        "indices.")
        raise RuntimeError("Exception raised in worker processes")
    # design rows in genotype-file order; phenotype-only samples would be
    # all-NaN for every marker and never enter a fit
    Xd = X.loc[sample_order, :]
    yd = y.loc[sample_order, :]
    n_file = len(samples)

    columns = list(Xd.columns)
    W = None
    inter_keys = []
    if gwas_interaction:
        inter_keys = list(gwas_interaction.keys())
        W = np.column_stack([
            np.prod(Xd[list(cols)].values.astype(float), axis=1)
            for cols in gwas_interaction.values()])

    options = dict(test.gwas_options())
    strata = None
    if model == "coxph":
        if "strata" in yd.columns:                  # survival.py:54-56
            strata = yd["strata"].values
        if "intercept" in columns:
            columns.remove("intercept")
        C = Xd[columns].values.astype(float)
        y_main = yd["tte"].values.astype(float)
        event = yd["event"].values.astype(float)
    else:
        C = Xd[columns].values.astype(float)
        y_main = yd.iloc[:, 0].values.astype(float)
        event = None
    fit_columns = columns + ["SNPs"] + inter_keys

    rank, world = sharding.rank_world()
    eng = get_engine(sharding.local_device())
    # the stock NameFilter / MAFFilter are evaluated on the packed rows (the
    # names against the .bim table, the MAF on the device); a subclass or any
    # other callable gets the decoded markers, one at a time, like the reference
    stock = all(type(f) in _stock_filter_types() for f in variant_predicates)
    packed_path = hasattr(genotypes, "packed") and stock
    if packed_path:
        # raw .bed rows go to the GPU as they are: the engine aligns them
        eng.set_design(model, y_main, C, W=W, event=event,
                       sample_pos=geno_index.astype(np.int32), n_file=n_file,
                       maf_t=maf_t, strata=strata, **options)
    else:
        # decoded genotypes are aligned to the design rows on the host
        eng.set_design(model, y_main, C, W=W, event=event, maf_t=maf_t,
                       strata=strata, **options)
    if sample_sex is not None:
        # analysis.py:141-148: get_sex_maf on the complete cases
        eng.set_sample_sex(
            sample_sex.reindex(Xd.index).values.astype(float))

    def consume(res, meta):
        _dispatch_block(ResultBlock(model, fit_columns, res, meta, maf_t),
                        subscribers, messages)

    from concurrent.futures import ThreadPoolExecutor
    if packed_path:
        # Batches of BATCH_SNPS markers are dealt to the ranks in turn: round i
        # holds batches i * world .. i * world + world - 1, so that the blocks
        # gathered after every round are in marker order and rank 0's sinks
        # write while the GPUs work on the next round.  The sinks of batch i run
        # on one helper thread (rows stay in marker order) -- both sides spend
        # their time outside the GIL.
        plan = _PackedPlan(genotypes, variant_predicates, messages, rank, world)
        n_rounds = -(-len(plan.batches) // world) if plan.batches else 0
        with ThreadPoolExecutor(max_workers=1) as pool:
            pending = []
            for i in range(n_rounds):
                j = i * world + rank
                mine = plan.run(eng, j) if j < len(plan.batches) else None
                if world == 1:
                    blocks = [mine]
                else:
                    sizes = [len(plan.batches[i * world + r])
                             if i * world + r < len(plan.batches) else 0
                             for r in range(world)]
                    blocks = sharding.gather_round(mine, sizes, eng.n_fields,
                                                   eng.n_params)
                if rank != 0:
                    continue
                for r, blk in enumerate(blocks):
                    if blk is None:
                        continue
                    res, meta = plan.finish(i * world + r, blk)
                    if len(meta):
                        pending.append(pool.submit(consume, res, meta))
                while len(pending) > 1:
                    pending.pop(0).result()
            for f in pending:
                f.result()
        logger.info("Analysis completed")
        return

    blocks = _decoded_blocks(genotypes, eng, variant_predicates, messages,
                             geno_index, rank, world)
    if world == 1:
        with ThreadPoolExecutor(max_workers=1) as pool:
            pending = []
            for res, meta in blocks:
                if res is None:
                    continue
                pending.append(pool.submit(consume, res, meta))
                while len(pending) > 1:
                    pending.pop(0).result()
            for f in pending:
                f.result()
    else:
        # decoded sources: every rank finishes its range, then ONE gather of the
        # fixed-size result blocks to rank 0
        from .engine import Results, NUM_IFIELDS
        parts = [(r, m) for r, m in blocks if r is not None]
        meta = [x for _r, m in parts for x in m]
        if parts:
            res = Results(np.concatenate([r.out for r, _m in parts], axis=1),
                          np.concatenate([r.iout for r, _m in parts], axis=1),
                          eng.n_params)
        else:
            res = Results(np.zeros((eng.n_fields, 0)),
                          np.zeros((NUM_IFIELDS, 0), dtype=np.int32),
                          eng.n_params)
        res, meta = sharding.gather_results(res, meta)
        if sharding.is_primary() and len(meta):
            consume(res, meta)
    logger.info("Analysis completed")


def _stock_filter_types():
    from .modelspec.predicates import MAFFilter, NameFilter
    return (NameFilter, MAFFilter)


class _PackedPlan(object):
    """Fast path: the reader exposes SNP-major 2-bit rows (PLINK .bed).
    ``NameFilter`` is applied to the .bim table before anything is read;
    ``MAFFilter`` (predicates.py:33-57: MAF over ALL samples of the file) is
    evaluated on the device next to the fit, and the markers it rejects are
    dropped from the block and listed as skipped.  Every rank builds the same
    plan (the batches and their metadata follow from the .bim table alone), so
    nothing but the fixed-size result blocks ever travels between ranks."""

    def __init__(self, reader, predicates, messages, rank, world=1):
        from .modelspec.predicates import MAFFilter, NameFilter
        self.table = reader.variant_table()
        self.packed = reader.packed()
        self.messages = messages
        keep = np.ones(len(self.table), dtype=bool)
        self.maf_min = None
        for f in predicates:
            if type(f) is NameFilter:
                keep &= np.array([t[0] in f.extract for t in self.table], dtype=bool)
            elif type(f) is MAFFilter:
                self.maf_min = f.maf if self.maf_min is None else max(self.maf_min, f.maf)
        if rank == 0:
            messages["skipped"].extend(
                t[0] for t, k in zip(self.table, keep) if not k)
        idx = np.flatnonzero(keep)
        # several ranks: at least two rounds of smaller batches, so that every GPU works
        bs = BATCH_SNPS if world == 1 else \
            max(18944, min(BATCH_SNPS, -(-len(idx) // (2 * world))))
        self.batches = [idx[s:s + bs] for s in range(0, len(idx), bs)]

    def run(self, eng, j):
        """Batch ``j`` on this rank's GPU -> (Results, file MAF or None)."""
        sel = self.batches[j]
        if sel[-1] - sel[0] + 1 == len(sel):
            rows = self.packed[sel[0]:sel[-1] + 1]         # contiguous: read in place
        else:
            rows = self.packed[sel]
        file_maf = None
        if self.maf_min is not None:
            file_maf = np.empty(len(sel), dtype=np.float64)
        return eng.run_packed_host(rows, file_maf=file_maf), file_maf

    def finish(self, j, block):
        """(Results, file MAF) of batch ``j`` -> (Results, metadata) after the
        MAF filter (rank 0)."""
        res, file_maf = block
        meta = [self.table[i] for i in self.batches[j]]
        if self.maf_min is not None:
            ok = file_maf >= self.maf_min                  # NaN (nothing genotyped) -> skipped
            if not ok.all():
                self.messages["skipped"].extend(m[0] for m, k in zip(meta, ok) if not k)
                from .engine import Results
                res = Results(np.ascontiguousarray(res.out[:, ok]),
                              np.ascontiguousarray(res.iout[:, ok]), res.n_params)
                meta = [m for m, k in zip(meta, ok) if k]
        return res, meta


def _decoded_blocks(reader, eng, predicates, messages, geno_index, rank,
                    world):
    """Generic path: any geneparse-like reader, decoded on the host.
    Integer-valued batches are re-packed to 2 bits, anything else goes to the
    engine as FP64 dosages aligned to the design rows."""
    n_file = len(reader.get_samples())
    rows, meta = [], []

    def flush():
        block = np.ascontiguousarray(np.vstack(rows)[:, geno_index])
        if _integer_valued(block):
            from .genotypes.plink import pack_rows
            res = eng.run_packed_host(pack_rows(block))
        else:
            res = eng.run_dosage_host(block)
        out = (res, list(meta))
        del rows[:], meta[:]
        return out

    count = -1
    total = None
    if world > 1:
        total = reader.get_number_variants()
    lo, hi = (0, None) if world == 1 else sharding.shard_range(total, world,
                                                              rank)
    any_block = False
    for snp in reader.iter_genotypes():
        count += 1
        try:
            if not all([f(snp) for f in predicates]):
                if rank == 0:
                    messages["skipped"].append(snp.variant.name)
                continue
        except StopIteration:
            break
        if count < lo or (hi is not None and count >= hi):
            continue
        g = np.asarray(snp.genotypes, dtype=float)
        if g.shape[0] != n_file:
            raise ValueError("{}: {} genotypes for {} samples".format(
                snp.variant.name, g.shape[0], n_file))
        rows.append(g)
        meta.append((snp.variant.name, snp.variant.chrom, snp.variant.pos,
                     snp.coded, snp.reference))
        if len(rows) >= 8192:
            any_block = True
            yield flush()
    if rows:
        any_block = True
        yield flush()
    if not any_block:
        yield None, []
