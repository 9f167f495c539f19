"""Freeze the two census rows the household samplers read (sample_households.py:10-33) into a small JSON file.

    python oracle/freeze_census.py      (run where /root/reference exists)

The reference opens `World_Age_2019.csv` / `AgeSpecificFertility.csv` relative to the CWD; on a box without the
reference tree `covid19-demography_b200/population.py` falls back to this frozen copy of the rows it needs.
Data only -- no reference code is copied.
"""
import csv, json, os

REF = os.environ.get("SEIR_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "covid19-demography_b200", "data", "census.json")


def row_of(path, country, count, encoding=None):
    with open(os.path.join(REF, path), encoding=encoding) as f:
        for row in csv.reader(f, delimiter=","):
            if row and row[0] == country:
                return [float(x) for x in row[1:1 + count]]
    raise KeyError(country)


def china_2020_row():
    """The UN 2020 single-year age estimates for China that sample_households_china (:43-45) and p_diabetes_china
    (sample_comorbidities.py:39) carry as an inline list: 101 numbers, read from the source text as data."""
    import re
    text = open(os.path.join(REF, "sample_households.py")).read()
    m = re.search(r"age_distribution = \[([0-9.,\s]+)\]", text)
    vals = [float(x) for x in m.group(1).split(",")]
    assert len(vals) == 101
    return vals


if __name__ == "__main__":
    out = {"china_age_distribution_2020": china_2020_row(),
           "source": "World_Age_2019.csv / AgeSpecificFertility.csv of bwilder0/COVID19-Demography",
           "age_distribution": {c: row_of("World_Age_2019.csv", c, 101) for c in ("Italy", "China")},
           "mother_birth_age_distribution": {c: row_of("AgeSpecificFertility.csv", c, 7, "latin-1") for c in ("Italy", "China")}}
    with open(OUT, "w") as f:
        json.dump(out, f)
    print(OUT, os.path.getsize(OUT), "bytes")
