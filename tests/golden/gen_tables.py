"""Generate tests/golden/tables_{usgs,modis}.json from the reference's shipped run/*.TBL.

Run in the build container (where /root/reference exists):  python tests/golden/gen_tables.py
The JSON files hold the parsed parameter VALUES (fp32, shortest round-trip decimals); they are the
fixture the GPU box uses, where /root/reference does not exist.
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from noahmp_b200 import tables  # noqa: E402

REF = os.environ.get("NOAHMP_REFERENCE_RUN", "/root/reference/run")
here = os.path.dirname(os.path.abspath(__file__))
for dataset, tag in (("USGS", "usgs"), ("MODIFIED_IGBP_MODIS_NOAH", "modis")):
    d = tables.read_tables(REF, dataset)
    tables.tables_to_json(d, os.path.join(here, f"tables_{tag}.json"))
    print(dataset, "nveg", d["nveg"], "lucats", d["lucats"], "slcats", d["slcats"])
