import csv
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
ENCODING_NAMES = ["cl100k_base", "r50k_base", "p50k_base", "p50k_edit"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def load_golden(name):
    """Rows (input, ids, ids at maxTokens=10) of the reference's <name>_encodings.csv
    (lib/src/test/resources; @CsvFileSource(numLinesToSkip=1) trims the space before a quoted field)."""
    rows = []
    with open(os.path.join(GOLDEN, name + "_encodings.csv"), newline="", encoding="utf-8") as f:
        r = csv.reader(f, skipinitialspace=True)
        next(r)
        for row in r:
            if not row:
                continue
            rows.append((row[0], [int(x) for x in row[1].strip("[]").split(",") if x.strip()],
                         [int(x) for x in row[2].strip("[]").split(",") if x.strip()]))
    return rows


@pytest.fixture(scope="session")
def oracles():
    from oracle import jo
    jo.build()
    return {n: jo.OracleEncoding.builtin(n) for n in ENCODING_NAMES}


@pytest.fixture(scope="session")
def gpu_encodings():
    """One GPU encoding per predefined type, through the registry like the reference's tests."""
    import jtokkit_b200 as jt
    reg = jt.Encodings.new_lazy_encoding_registry()
    return {n: reg.get_encoding(jt.EncodingType.from_name(n)) for n in ENCODING_NAMES}
