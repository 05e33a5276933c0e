"""The line-audit of the oracle as a test (build container only: it reads /root/reference): for every routine of the
reference's physics the multiset of numeric literals of the Fortran text is compared with the literals of the oracle's
function of the same name (tools/audit_constants.py).  A literal on one side only is a candidate transcription slip;
what remains today is listed here with the reason it is not one — anything new fails the test."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/phys"), reason="reference tree not present")

# routine -> (fortran-only, oracle-only, why this is not a transcription difference)
ALLOW = {
    "ALBEDO": (["100"], [], "statement label of the night-time GOTO 100 (:2356)"),
    "ATM_GLACIER": ([], ["0.1", "0.9"], "the glacier path calls the land ATM; its convective / large-scale split is unused there"),
    "BARE_FLUX": (["50"], ["1e-05"], "TDC statement function (50.) lives in its own oracle function; `EHB2.lt.1.E-5`: the "
                                     "literal touches the relational operator and the Fortran scanner skips it"),
    "CO2FLUX": (["40"], [], "declared but unused local constant"),
    "ERROR": (["256"], [], "CHARACTER(len=256) message"),
    "ERROR_GLACIER": (["256"], ["1000"], "message length; compared with the LAND ERROR (the glacier one is inlined in NOAHMP_GLACIER)"),
    "FRH2O": (["1001", "1002", "80", "920"], [], "statement labels, an unused DICE = 920., line-length comment"),
    "GLACIER_FLUX": (["50"], ["1e-05"], "as BARE_FLUX"),
    "GROUNDWATER": (["8"], [], "REAL(KIND=8) S_NODE -> double"),
    "NOAHMP_GLACIER": (["0.0001", "10", "256"], ["0.01", "0.1", "0.3", "0.378", "0.5", "0.622", "0.7"],
                       "wrf_debug call and message; the oracle inlines ATM_GLACIER / ENERGY_GLACIER into this function"),
    "RADIATION_GLACIER": (["1e-06"], [], "MPE parameter declared, never used"),
    "REDPRM": (["256", "30"], [], "message length; NSLTYPE array bound"),
    "SFCDIF2": (["1e-08"], [], "EPSA parameter declared, never used"),
    "SNOW_AGE": ([], ["800"], "`SNEQV.GT.800.`: literal adjacent to the relational operator, skipped by the Fortran scanner"),
    "SNOW_AGE_GLACIER": ([], ["800"], "as SNOW_AGE"),
    "TSNOSOI": (["0.5", "256"], [], "dead energy-check code after the RETURN (:5797); message length"),
    "VEGE_FLUX": (["50", "80"], [], "TDC statement function; unused loop labels"),
    "WTABLE_MMF_NOAHMP": ([], ["0.45509"], "FANGLE is a PARAMETER of LATERALFLOW, inlined into the oracle's WTABLE"),
    "ZWTEQ": ([], ["0.01"], "`.LE.0.01`-style literal adjacent to a relational operator"),
}
NOT_RESTATED = {"BVOCFLUX": "never called", "LATERALFLOW": "inlined into WTABLE", "NOAHMP_OPTIONS": "options are a struct",
                "NOAHMP_OPTIONS_GLACIER": "options are a struct", "READ_MP_VEG_PARAMETERS": "table reader: product code, tested in test_tables.py",
                "SFCDIF3": "non-functional offline", "SFCDIF4": "non-functional offline"}


def test_numeric_literals_of_every_routine_match_the_reference():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import audit_constants
    diff, missing = audit_constants.audit(ROOT)
    assert sorted(missing) == sorted(NOT_RESTATED), (missing, "a reference routine lost its oracle counterpart")
    unexpected = {}
    for name, (f_only, c_only) in diff.items():
        want = ALLOW.get(name)
        if want is None or (sorted(want[0]), sorted(want[1])) != (f_only, c_only):
            unexpected[name] = (f_only, c_only)
    assert not unexpected, f"numeric literals differ from the reference: {unexpected}"
    stale = [n for n in ALLOW if n not in diff]
    assert not stale, f"allow-list entries no longer needed: {stale}"
