"""The cases of genetest/tests/test_phenotypes.py replayed on
genetest_b200.phenotypes (the containers ``analysis.execute`` pulls the
outcome, covariates and sample sex from)."""

import numpy as np
import pandas as pd
import pytest

from genetest_b200.phenotypes.core import PhenotypesContainer
from genetest_b200.phenotypes.text import TextPhenotypes

ROWS = [("s1", 1, 1.2, "a", 0), ("s2", 234, 34.3, "sdc", 1),
        ("s3", -999999, np.nan, "csgb", 1), ("s4", 5463, 16.8, "999999", 0),
        ("s5", 2356, 574.8, "jyrhg", 0), ("s6", np.nan, 567.2, "bvnchy", 1),
        ("s7", 57634, 3134.3, np.nan, 0), ("s8", 346, 999999, "dfgfjgsd", np.nan),
        ("s9", 34626, 3421.3, "fhyrj", 0)]
COLS = ["sample", "V1", "V2", "V3", "sex"]


def _write(path, cols, rows):
    with open(path, "w") as f:
        print(*cols, file=f)
        for r in rows:
            print(*["" if (isinstance(v, float) and np.isnan(v)) else v
                    for v in r], file=f)


@pytest.fixture
def pheno(tmp_path):
    path = str(tmp_path / "pheno.txt")
    _write(path, COLS, ROWS)
    expected = pd.DataFrame(ROWS, columns=COLS).set_index("sample")
    params = dict(filename=path, sample_column="sample", field_separator=" ",
                  missing_values=None, repeated_measurements=False,
                  sex_column="sex")
    return params, expected


def test_core_is_abstract():
    with pytest.raises(NotImplementedError):
        PhenotypesContainer().get_phenotypes()
    with pytest.raises(NotImplementedError):
        PhenotypesContainer().close()
    with pytest.raises(ValueError):
        PhenotypesContainer().get_sex()


def test_text_normal_subset_sex_and_repr(pheno):
    params, expected = pheno
    with TextPhenotypes(**params) as c:
        assert expected.equals(c.get_phenotypes())
        assert c.get_nb_samples() == 9 and not c.is_repeated()
        assert expected.loc[:, ["V1", "V3"]].equals(
            c.get_phenotypes(("V1", "V3")))
        with pytest.raises(KeyError):
            c.get_phenotypes(("V1", "V3", "VX"))
        assert expected.loc[:, "sex"].equals(c.get_sex())
        assert str(c) == "TextPhenotypes(9 samples, 4 variables)"
        assert c.get_phenotypes().index.name == "sample_id"


@pytest.mark.parametrize("missing, cells", [
    ("999999", [("s4", "V3"), ("s8", "V2")]),
    (["-999999", "999999"], [("s3", "V1"), ("s4", "V3"), ("s8", "V2")]),
    ({"V3": ["999999"]}, [("s4", "V3")]),
])
def test_text_missing_value_codes(pheno, missing, cells):
    params, expected = pheno
    params["missing_values"] = missing
    for row, col in cells:
        expected.loc[row, col] = np.nan
    assert expected.equals(TextPhenotypes(**params).get_phenotypes())


def test_text_repeated_measurements(pheno):
    params, _ = pheno
    cols = ["sample", "Time", "V1", "V2", "V3", "sex"]
    rows = [("s1", 13, 1, 1.2, "a", 0), ("s1", 125, 1, 1.2, "a", 0),
            ("s1", 356, 1, 1.2, "a", 0), ("s2", 12, 5463, 16.8, "b", 1),
            ("s2", 34, 5463, 16.8, "b", 1), ("s2", 67, 5463, 16.8, "b", 1),
            ("s3", 1, 57634, 3134.3, "c", np.nan),
            ("s3", 5, 57634, 3134.3, "c", np.nan),
            ("s3", 10, 57634, 3134.3, "c", np.nan),
            ("s4", 2, 57635, 3134.6, "b", 1), ("s4", 6, 57635, 3134.6, "b", 1),
            ("s4", 11, 57635, 3134.6, "b", 1)]
    _write(params["filename"], cols, rows)
    with pytest.raises(ValueError):             # duplicated ids need the flag
        TextPhenotypes(**params)
    params["repeated_measurements"] = True
    c = TextPhenotypes(**params)
    assert pd.DataFrame(rows, columns=cols).set_index("sample").equals(
        c.get_phenotypes())
    assert c.get_nb_samples() == 4 and c.is_repeated()
    assert pd.Series([0, 1, np.nan, 1], index=["s1", "s2", "s3", "s4"]) \
        .equals(c.get_sex())


def test_text_sex_errors(pheno):
    params, _ = pheno
    rows = [r if r[0] != "s4" else ("s4", 5463, 16.8, "999999", 2)
            for r in ROWS]
    _write(params["filename"], COLS, rows)
    with pytest.raises(ValueError):
        TextPhenotypes(**params).get_sex()
    no_sex = dict(params)
    del no_sex["sex_column"]
    with pytest.raises(ValueError):
        TextPhenotypes(**no_sex).get_sex()
    cols = ["sample", "Time", "V1", "V2", "V3", "sex"]
    rows = [("s1", 13, 1, 1.2, "a", 0), ("s1", 125, 1, 1.2, "a", 0),
            ("s1", 356, 1, 1.2, "a", 1), ("s2", 12, 5463, 16.8, "b", 1),
            ("s2", 34, 5463, 16.8, "b", 1), ("s2", 67, 5463, 16.8, "b", np.nan),
            ("s3", 1, 57634, 3134.3, "c", np.nan),
            ("s3", 5, 57634, 3134.3, "c", np.nan),
            ("s3", 10, 57634, 3134.3, "c", np.nan)]
    _write(params["filename"], cols, rows)
    params["repeated_measurements"] = True
    with pytest.raises(ValueError) as err:
        TextPhenotypes(**params).get_sex()
    assert str(err.value) == "Samples with duplicated different sex: s1,s2"
