"""Parameter-table semantics: the library's C++ reader and the Python reader agree bit for bit with the golden
fixture generated from the reference's shipped run/*.TBL (tests/golden/gen_tables.py)."""
import os

import numpy as np
import pytest

from noahmp_b200 import _capi, tables

REF_RUN = "/root/reference/run"


def _same_bits(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype.kind == "f" or b.dtype.kind == "f":
        return np.array_equal(a.astype(np.float32).view(np.uint32), b.astype(np.float32).view(np.uint32))
    return np.array_equal(a, b)


@pytest.mark.parametrize("dataset", ["USGS", "MODIFIED_IGBP_MODIS_NOAH"])
def test_cpp_reader_roundtrip(built, tmp_path, dataset):
    """golden values -> .TBL files (writer) -> C++ reader == Python reader == golden values."""
    import noahmp_b200
    gold = tables.default_tables(dataset)
    tables.write_tables(str(tmp_path), gold, dataset)
    got = _capi.tables_to_dict(noahmp_b200.read_tables(str(tmp_path), dataset))
    pyd = tables.read_tables(str(tmp_path), dataset)
    for k, v in gold.items():
        assert _same_bits(got[k], v), k
        assert _same_bits(pyd[k], v), k


@pytest.mark.skipif(not os.path.isdir(REF_RUN), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("dataset", ["USGS", "MODIFIED_IGBP_MODIS_NOAH"])
def test_cpp_reader_on_reference_files(built, dataset):
    """The shipped reference files parse to the committed golden fixture."""
    import noahmp_b200
    gold = tables.default_tables(dataset)
    got = _capi.tables_to_dict(noahmp_b200.read_tables(REF_RUN, dataset))
    for k, v in gold.items():
        assert _same_bits(got[k], v), k


def test_known_values(tables_usgs):
    """Spot values straight from run/*.TBL."""
    t = tables_usgs
    assert t["nveg"] == 27 and t["lucats"] == 27 and t["slcats"] == 19
    assert t["isbarren"] == 19 and t["issnow"] == 24 and t["iswater"] == 16
    assert np.float32(t["maxsmc"][0]) == np.float32(0.339)       # sand
    assert np.float32(t["bb"][2]) == np.float32(4.74)            # sandy loam
    assert t["cwpvt"][26] == np.float32(-1.0e36)                 # USGS CWPVT has 24 of 27 values (App. A #18)
    assert t["nrotbl"][0] == 1 and t["nrotbl"][13] == 4


def test_missing_dataset_is_fatal(built, tmp_path):
    import noahmp_b200
    gold = tables.default_tables("USGS")
    tables.write_tables(str(tmp_path), gold, "USGS")
    with pytest.raises(noahmp_b200.NoahmpError):
        noahmp_b200.read_tables(str(tmp_path), "NOT_A_DATASET")
