"""Formula parser for model specifications.

The reference generates its parser with grako from ``modelspec.ebnf``
(``genetest/modelspec/modelspec.ebnf:18-69``, semantics in
``genetest/modelspec/grammar.py:23-70``); grako is not available here, so the
fifteen rules are parsed by a small hand-written recursive-descent parser
with the same PEG behaviour (ordered choice, greedy optional ``as`` names)::

    y ~ x1 + x2                 standard model with covariates
    y ~ x1 + x1 * x2 as z       (named) interaction
    y | male ~ x                stratified analysis
    y | diabetes = 0 ~ x        subgroup analysis
    y ~ g(rs12345) + x          a named variant as covariate
    y ~ SNPs + x                GWAS
    y ~ SNPs + x + SNPs*x       GWAS with a SNP x covariate interaction
    y ~ factor(x) as f          factors
    [tte=t, event=e] ~ x        labelled outcomes
"""

import re
import warnings

from . import core

__all__ = ["parse_formula", "parse_modelspec", "FormulaError"]

_NAME = re.compile(r"[A-Za-z0-9_:]+")
_INT = re.compile(r"[0-9]+")
_STRING = re.compile(r"\".+\"|'.+'")
_WS = re.compile(r"\s*")


class FormulaError(RuntimeError):
    pass


class _Parser(object):
    def __init__(self, text):
        self.text = text
        self.pos = 0

    # -- scanning -----------------------------------------------------------
    def _skip(self):
        self.pos = _WS.match(self.text, self.pos).end()

    def _lit(self, token):
        """Consume a literal token (after whitespace); False if absent."""
        self._skip()
        if self.text.startswith(token, self.pos):
            self.pos += len(token)
            return True
        return False

    def _expect(self, token):
        if not self._lit(token):
            self._fail("expected {!r}".format(token))

    def _regex(self, pattern):
        self._skip()
        m = pattern.match(self.text, self.pos)
        if m is None:
            return None
        self.pos = m.end()
        return m.group(0)

    def _fail(self, what):
        raise FormulaError("{} at position {} of {!r}".format(
            what, self.pos, self.text))

    def _attempt(self, rule):
        """Ordered choice: run ``rule``; rewind and return None if it fails."""
        start = self.pos
        try:
            return rule()
        except FormulaError:
            self.pos = start
            return None

    # -- terminals ------------------------------------------------------------
    def name(self):
        value = self._regex(_NAME)
        if value is None:
            self._fail("expected a name")
        return value

    def optional_as(self):
        start = self.pos
        if self._lit("as"):
            # "as" must be a word of its own ("assay" is a name)
            if self.pos < len(self.text) and \
                    not self.text[self.pos].isspace():
                self.pos = start
                return None
            value = self._regex(_NAME)
            if value is not None:
                return value
        self.pos = start
        return None

    def literal(self):
        value = self._regex(_INT)
        if value is not None:
            return int(value)
        value = self._regex(_STRING)
        if value is not None:
            return value[1:-1]
        self._fail("expected an integer or a quoted string")

    # -- rules ----------------------------------------------------------------
    def phenotype(self):
        return core.phenotypes[self.name()]

    def genotype(self):
        self._expect("g(")
        variant = self.name()
        self._expect(")")
        return core.genotypes[variant]

    def phenotype_or_variant(self):
        found = self._attempt(self.genotype)
        return found if found is not None else self.phenotype()

    def factor(self):
        self._expect("factor(")
        phen = self.phenotype()
        self._expect(")")
        return core.factor(phen, name=self.optional_as())

    def pow(self):
        self._expect("pow(")
        phen = self.phenotype()
        self._expect(",")
        power = self._regex(_INT)
        if power is None:
            self._fail("expected an integer power")
        self._expect(")")
        return core.pow(phen, int(power), name=self.optional_as())

    def ln(self):
        self._expect("ln(")
        phen = self.phenotype()
        self._expect(")")
        self.optional_as()
        return core.ln(phen)

    def log10(self):
        self._expect("log10(")
        phen = self.phenotype()
        self._expect(")")
        self.optional_as()
        return core.log10(phen)

    def snps(self):
        start = self.pos
        if self._lit("SNPs"):
            nxt = self.text[self.pos:self.pos + 1]
            if not _NAME.match(nxt):
                return core.SNPs
        self.pos = start
        self._fail("expected SNPs")

    def _interaction_ahead(self):
        """The rule's lookahead: a first term directly followed by ``*``,
        checked without running any semantic action."""
        start = self.pos
        try:
            self._skip()
            if self.text.startswith("g(", self.pos):
                self._expect("g(")
                self.name()
                self._expect(")")
            elif self.text.startswith("factor(", self.pos):
                self._expect("factor(")
                self.name()
                self._expect(")")
            else:
                self.name()
            return self._lit("*")
        except FormulaError:
            return False
        finally:
            self.pos = start

    def _interaction_term(self):
        for rule in (self.snps, self.factor, self.phenotype_or_variant):
            found = self._attempt(rule)
            if found is not None:
                return found
        self._fail("expected an interaction term")

    def interaction(self):
        if not self._interaction_ahead():
            self._fail("not an interaction")
        terms = [self._interaction_term()]
        while self._lit("*"):
            terms.append(self._interaction_term())
        alias = self.optional_as()
        if core.SNPs in terms:
            others = tuple(t for t in terms if t != core.SNPs)
            return core.gwas_interaction(
                *others, name=alias if alias else "GWAS_INTER")
        return core.interaction(*terms, name=alias)

    def expression(self):
        for rule in (self.interaction, self.snps, self.genotype, self.factor,
                     self.ln, self.log10, self.pow, self.phenotype):
            found = self._attempt(rule)
            if found is not None:
                return found
        self._fail("expected a predictor")

    def predictors(self):
        out = [self.expression()]
        while self._lit("+"):
            out.append(self.expression())
        return out

    def labelled_outcome_group(self):
        self._expect("[")
        tags = {}
        while True:
            key = self.name()
            self._expect("=")
            tags[key] = self.phenotype()
            if not self._lit(","):
                break
        self._expect("]")
        return tags

    def condition(self):
        entity = self.phenotype_or_variant()
        level = None
        if self._lit("="):
            level = self.literal()
        return {"name": entity, "level": level}

    def model(self):
        self._skip()
        if self.text.startswith("[", self.pos):
            outcome = self.labelled_outcome_group()
        else:
            outcome = self.phenotype_or_variant()
        conditions = None
        if self._lit("|"):
            conditions = [self.condition()]
            while self._lit(","):
                conditions.append(self.condition())
        self._expect("~")
        predictors = self.predictors()
        self._skip()
        if self.pos != len(self.text):
            self._fail("unexpected trailing text")
        return {"outcome": outcome, "conditions": conditions,
                "predictors": predictors}


def parse_formula(f):
    """Formula -> ``{"outcome", "conditions", "predictors"}`` ready for
    ``ModelSpec(**model)`` once ``conditions`` is handled (reference
    ``grammar.py:79-107``)."""
    try:
        return _Parser(f).model()
    except FormulaError as e:
        raise RuntimeError("Could not parse the model formula: {}".format(e))


def parse_modelspec(s):
    warnings.warn("use 'parse_formula' instead", DeprecationWarning)
    return parse_formula(s)
