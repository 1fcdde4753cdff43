"""Result sinks (``genetest/subscribers.py``).

The analysis calls ``init(modelspec)`` once, ``_update_gwas_interaction`` /
``_update_current_subset`` before results when they apply, ``handle(results)``
once per successful test from the host thread that owns the gathered result
blocks, and ``close()`` at the end.  Subscribers that define
``handle_block(block)`` receive whole result blocks (see
:class:`genetest_b200.analysis.ResultBlock`) instead of one dict per marker.
"""

import json
import logging
import sys

from .modelspec import result as analysis_results

__all__ = ["Subscriber", "Print", "ResultsMemory", "RowWriter", "GWASWriter",
           "subscriber_error"]

logger = logging.getLogger(__name__)


class Subscriber(object):
    """Base class of result sinks."""

    def init(self, modelspec):
        """(Re)initialise with the model specification of the analysis."""
        self.modelspec = modelspec
        self._reset_subset()

    def close(self):
        pass

    def _update_current_subset(self, info):
        """Called before the results of a subgroup: ``info`` maps the
        stratification variables to their current level."""
        self.subset_info = info
        self.subset_info_str = ";".join(
            "{}:{}".format(k, v) for k, v in sorted(info.items()))

    def _reset_subset(self):
        self.subset_info = None

    def _update_gwas_interaction(self, columns):
        """Called with the sorted SNP x covariate interaction columns."""
        pass

    def handle(self, results):
        raise NotImplementedError()

    @staticmethod
    def _apply_translation(translation, results):
        return {translation.get(k, k): v for k, v in results.items()}


class ResultsMemory(Subscriber):
    """Keeps every (translated) result dictionary in ``results``."""

    def __init__(self):
        self.results = []

    def handle(self, results):
        self.results.append(self._apply_translation(
            self.modelspec.get_translations(), results))

    def _get_gwas_results(self):
        return {r["SNPs"]["name"]: r for r in self.results}


class Print(Subscriber):
    """Dumps every result as JSON on stdout."""

    def __init__(self, raw=False):
        self.raw = raw

    def handle(self, results):
        if self.subset_info:
            results["MODEL"]["subset_info"] = self.subset_info
        if not self.raw:
            results = self._apply_translation(
                self.modelspec.get_translations(), results)
        print(json.dumps(results, indent=2, default=_jsonable))


def _jsonable(o):
    try:
        return o.item()
    except AttributeError:
        return str(o)


class RowWriter(Subscriber):
    """One delimited row per result; ``columns`` is a list of
    ``(header, Result accessor or literal string)``."""

    def __init__(self, filename=None, columns=None, header=False, sep="\t",
                 append=False):
        self.header = header
        self.sep = sep
        self.filename = filename
        self._set_columns(columns)
        self._f = open(filename, "a" if append else "w") if filename else None
        if self.header:
            self.print_header()

    def _set_columns(self, columns):
        self.columns = columns

    def _emit(self, line):
        if self._f is not None:
            self._f.write(line + "\n")
        else:
            print(line)

    def print_header(self):
        self._emit(self.sep.join(c[0] for c in self.columns))

    def close(self):
        if self._f:
            self._f.close()

    def handle(self, results):
        cells = [field if isinstance(field, str) else str(field.get(results))
                 for _name, field in self.columns]
        self._emit(self.sep.join(cells))
        if self._f is not None:
            self._f.flush()

    def handle_block(self, block):
        """Streaming sink: all successful markers of a result block at once,
        formatted by the native row formatter (``gt_format_rows``) -- the same
        bytes ``handle`` would write one dictionary at a time."""
        from .engine import format_rows
        info = getattr(self, "subset_info_str", None) \
            if self.subset_info is not None else None
        cols = [field if isinstance(field, str)
                else block.column(field.path, info)
                for _name, field in self.columns]
        data = format_rows(cols, keep=block.ok_mask, sep=self.sep)
        if self._f is not None and hasattr(self._f, "buffer"):
            # the rows are ASCII already: straight into the file's byte buffer
            self._f.flush()
            self._f.buffer.write(data)
            self._f.buffer.flush()
        elif self._f is not None:
            self._f.write(data.decode())
            self._f.flush()
        else:
            print(data.decode(), end="")


_STAT_COLUMNS = {
    "linear": ([("coef", "coef"), ("se", "std_err"), ("lower", "lower_ci"),
                ("upper", "upper_ci"), ("t", "t_value"), ("p", "p_value")],
               ("adj_r2", "r_squared_adj")),
    "logistic": ([("coef", "coef"), ("se", "std_err"), ("lower", "lower_ci"),
                  ("upper", "upper_ci"), ("t", "t_value"), ("p", "p_value")],
                 ("n_events", "nevents")),
    "coxph": ([("coef", "coef"), ("se", "std_err"), ("hr", "hr"),
               ("hr_lower", "hr_lower_ci"), ("hr_upper", "hr_upper_ci"),
               ("z", "z_value"), ("p", "p_value")],
              ("n_events", "nevents")),
    "mixedlm": ([("coef", "coef"), ("se", "std_err"), ("lower", "lower_ci"),
                 ("upper", "upper_ci"), ("z", "z_value"), ("p", "p_value")],
                ("n_groups", "n_groups")),
}


class GWASWriter(RowWriter):
    """Tab-delimited GWAS table: ``snp chr pos major minor maf n ll`` + the
    model column + the statistics of the SNPs column (or of every SNP x
    covariate interaction column, ``<column>:<stat>``); a ``subgroup`` column
    is added for stratified analyses."""

    def __init__(self, filename, test, sep="\t"):
        self._inter_already_updated = False
        self._final_columns = None
        r = analysis_results
        self._common_cols = [
            ("snp", r["SNPs"]["name"]), ("chr", r["SNPs"]["chrom"]),
            ("pos", r["SNPs"]["pos"]), ("major", r["SNPs"]["major"]),
            ("minor", r["SNPs"]["minor"]), ("maf", r["SNPs"]["maf"]),
            ("n", r["MODEL"]["nobs"]), ("ll", r["MODEL"]["log_likelihood"]),
        ]
        self._specific_cols = []
        self._specific_model_cols = []
        if test in _STAT_COLUMNS:
            stats, (model_header, model_field) = _STAT_COLUMNS[test]
            self._specific_cols = [(h, ("SNPs", f)) for h, f in stats]
            self._specific_model_cols = [
                (model_header, r["MODEL"][model_field])]
        else:
            logger.warning("{}: invalid test: only common columns will be "
                           "written to file.".format(test))
        super().__init__(filename=filename, columns=None, header=True,
                         append=False, sep=sep)

    @property
    def columns(self):
        if self._final_columns is None:
            stats = [(header, analysis_results[column][field])
                     for header, (column, field) in self._specific_cols]
            self._final_columns = (self._common_cols +
                                   self._specific_model_cols + stats)
        return self._final_columns

    def _set_columns(self, columns):
        pass

    def _reset_output_header(self):
        logger.debug("Re-writing the output header")
        self._final_columns = None
        self._f.seek(0)
        self.print_header()

    def _update_current_subset(self, info):
        if self.subset_info is None:
            # first subgroup: grow the table by one column
            self._common_cols.append(
                ("subgroup", analysis_results["MODEL"]["subset_info"]))
            self._reset_output_header()
        super()._update_current_subset(info=info)

    def _update_gwas_interaction(self, columns):
        if not self._inter_already_updated:
            self._specific_cols = [
                ("{}:{}".format(column, header), (column, field))
                for column in columns
                for header, (_c, field) in self._specific_cols
            ]
            self._reset_output_header()
            self._inter_already_updated = True
        super()._update_gwas_interaction(columns=columns)

    def handle(self, results):
        if self.subset_info is not None:
            results["MODEL"]["subset_info"] = self.subset_info_str
        super().handle(results=results)


def subscriber_error(message, abort=None):
    """A subscriber asked for a field the test does not produce."""
    logger.critical(
        "A subscriber for this analysis raised an exception. This is be "
        "because an invalid key was accessed from the results of the "
        "statistical test.\nUnknown field: '{}'".format(message))
    if abort:
        abort.set()
    sys.exit(1)
