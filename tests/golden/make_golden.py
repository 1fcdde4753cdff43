#!/usr/bin/env python
"""Regenerate the golden fixtures from the reference's own test-suite.

Run in the build container only (it reads ``/root/reference``, which does not
exist on the GPU box)::

    python tests/golden/make_golden.py

Outputs (committed):

* ``tests/golden/data/*.tsv``    -- the reference's known-answer data tables
  (``genetest/tests/data/statistics/*.txt.bz2``), decompressed verbatim.
* ``tests/golden/expected.json`` -- every SAS / R / hand number asserted by
  ``genetest/tests/test_{linear,logistic,coxph,mixedlm}.py``, pulled out of the test
  sources with ``ast`` (the expected-value expressions are evaluated, the
  "actual" side is recorded as (entity, field, transform)).

Nothing here is imported by the product; the tests only read the outputs.
"""

import ast
import bz2
import json
import math
import os
import sys

import numpy as np

REF = "/root/reference/genetest/tests"
HERE = os.path.dirname(os.path.abspath(__file__))

DATA_FILES = ["linear", "linear_missing", "logistic", "logistic_missing",
              "coxph", "factors", "mixedlm"]
TEST_FILES = {"linear": "test_linear.py", "logistic": "test_logistic.py",
              "coxph": "test_coxph.py", "mixedlm": "test_mixedlm.py"}


class _Inter(object):
    """Stand-in for the interaction entity: ``inter.id`` -> ``"INTER"``."""
    id = "INTER"


def _subscript_chain(node):
    """``results[a][b]`` -> ("results", [a_node, b_node]) or None."""
    keys = []
    while isinstance(node, ast.Subscript):
        keys.append(node.slice)
        node = node.value
    if isinstance(node, ast.Name) and keys:
        return node.id, list(reversed(keys))
    return None


def _find_actual(node):
    """Locate the ``results[..][..]`` access inside an assertion argument and
    name the transform wrapped around it."""
    chain = _subscript_chain(node)
    if chain and chain[0] == "results":
        return chain[1], "identity"
    # -np.log10(results[..][..])
    if (isinstance(node, ast.UnaryOp) and isinstance(node.op, ast.USub) and
            isinstance(node.operand, ast.Call) and node.operand.args):
        chain = _subscript_chain(node.operand.args[0])
        if chain and chain[0] == "results":
            return chain[1], "neglog10"
    # results[..][..] ** 2
    if isinstance(node, ast.BinOp) and isinstance(node.op, ast.Pow):
        chain = _subscript_chain(node.left)
        if chain and chain[0] == "results":
            return chain[1], "square"
    return None


def _eval(node, env):
    return eval(compile(ast.Expression(body=node), "<golden>", "eval"),
                {"np": np, "math": math}, env)


def extract(path, nrows):
    tree = ast.parse(open(path).read())
    out = {}
    for cls in [n for n in tree.body if isinstance(n, ast.ClassDef)]:
        for fn in [n for n in cls.body if isinstance(n, ast.FunctionDef)]:
            if not fn.name.startswith("test_"):
                continue
            env = {"inter": _Inter(), "n": nrows}
            rec = {"checks": [], "failed": None, "n_results": None,
                   "expected_failure": any(
                       getattr(d, "attr", None) == "expectedFailure"
                       for d in fn.decorator_list)}
            current = None
            for st in ast.walk(fn):
                pass
            for st in fn.body:
                # with-blocks hold the failed-file assertion
                stmts = [st]
                if isinstance(st, ast.With):
                    stmts = st.body
                for s in stmts:
                    if isinstance(s, ast.Assign) and \
                            isinstance(s.targets[0], ast.Name):
                        name = s.targets[0].id
                        chain = _subscript_chain(s.value)
                        if name == "results" and chain and \
                                chain[0] == "gwas_results":
                            current = _eval(chain[1][0], env)
                        elif name == "results":
                            current = None
                        elif name in ("p", "col", "inter_id"):
                            env[name] = _eval(s.value, env)
                        continue
                    if not (isinstance(s, ast.Expr) and
                            isinstance(s.value, ast.Call) and
                            isinstance(s.value.func, ast.Attribute)):
                        continue
                    call = s.value
                    kind = call.func.attr
                    if kind == "assertRaises":
                        continue
                    if kind not in ("assertEqual", "assertAlmostEqual"):
                        continue
                    places = None
                    for kw in call.keywords:
                        if kw.arg == "places":
                            places = _eval(kw.value, env)
                    a, b = call.args[0], call.args[1]
                    found = None
                    for actual, expected in ((b, a), (a, b)):
                        hit = _find_actual(actual)
                        if hit:
                            found = (hit, expected)
                            break
                    if found is None:
                        # n_results or failed-file content
                        src = ast.unparse(call)
                        if "len(gwas_results" in src:
                            rec["n_results"] = _eval(a, env)
                        elif "line.split" in src:
                            rec["failed"] = _eval(a, env)
                        continue
                    (keys, transform), expected = found
                    keys = [_eval(k, env) for k in keys]
                    rec["checks"].append({
                        "snp": current, "entity": keys[0], "field": keys[1],
                        "transform": transform,
                        "expected": _eval(expected, env),
                        "places": (7 if places is None else places)
                        if kind == "assertAlmostEqual" else None,
                    })
            # assertRaises(StatsError) messages: "str(cm.exception)" asserts
            for node in ast.walk(fn):
                if isinstance(node, ast.Call) and \
                        getattr(node.func, "attr", "") == "assertEqual" and \
                        "cm.exception" in ast.unparse(node):
                    rec["raises"] = _eval(node.args[0], env)
            out[fn.name] = rec
    return out


def main():
    os.makedirs(os.path.join(HERE, "data"), exist_ok=True)
    nrows = {}
    for name in DATA_FILES:
        raw = bz2.open(os.path.join(
            REF, "data", "statistics", name + ".txt.bz2"), "rt").read()
        with open(os.path.join(HERE, "data", name + ".tsv"), "w") as f:
            f.write(raw)
        nrows[name] = len(raw.strip().splitlines()) - 1
    expected = {}
    for model, fname in TEST_FILES.items():
        expected[model] = extract(os.path.join(REF, fname), nrows[model])
    with open(os.path.join(HERE, "expected.json"), "w") as f:
        json.dump(expected, f, indent=1, sort_keys=True)
    for model, tests in expected.items():
        n = sum(len(t["checks"]) for t in tests.values())
        print("{}: {} tests, {} numeric/equality checks".format(
            model, len(tests), n), file=sys.stderr)


if __name__ == "__main__":
    main()
