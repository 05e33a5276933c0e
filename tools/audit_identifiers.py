"""Line-audit aid #2: per routine, compare how often each identifier occurs in the executable Fortran statements
and in the oracle's C++ function of the same name.  Prints identifiers whose counts differ (candidates for a wrong
variable in a formula).  Needs /root/reference (build container only)."""
import collections, re, sys
F = ["/root/reference/phys/module_sf_noahmplsm.F90", "/root/reference/phys/module_sf_noahmp_glacier.F90",
     "/root/reference/phys/module_sf_noahmp_groundwater.F90"]
C = ["oracle/nmo_land1.cpp", "oracle/nmo_land2.cpp", "oracle/nmo_land3.cpp", "oracle/nmo_glacier.cpp",
     "oracle/nmo_groundwater.cpp"]
KW = set("IF THEN ELSE ELSEIF ENDIF END DO ENDDO CALL REAL INTEGER INTENT IN OUT INOUT DIMENSION PARAMETER LOGICAL AND OR NOT "
         "MIN MAX AMIN1 AMAX1 ABS EXP LOG ALOG SQRT TANH ATAN COS TAN ACOS MOD SIGN FLOAT NINT INT EXIT CYCLE RETURN GOTO "
         "CONTINUE SUBROUTINE IMPLICIT NONE WRITE PRINT CHARACTER LEN KIND DATA SAVE LT LE GT GE EQ NE TRUE FALSE "
         "FOR INT FLOAT CONST STATIC VOID BOOL AUTO BREAK POW POWI LOG10 ALOG10 WRF_ERROR_FATAL MESSAGE STOP USE ONLY "
         "MIN3 IMIN IMAX DOUBLE DPOW WHILE TRIM".split())
def idents(text, fortran):
    out = collections.Counter()
    for line in text.splitlines():
        line = line.split("!")[0] if fortran else line.split("//")[0]
        if fortran and re.match(r"\s*(REAL|INTEGER|LOGICAL|CHARACTER|USE|IMPLICIT|SUBROUTINE|END SUBROUTINE|DATA)\b", line, re.I):
            continue
        if not fortran and re.match(r"\s*(static\s+)?(void|float|int|const float|const int|bool)\s+\w+\s*[\(;=,]", line) and "(" in line and line.rstrip().endswith(("{", ",")):
            continue
        line = re.sub(r"'[^']*'|\"[^\"]*\"", "", line)
        for m in re.finditer(r"[A-Za-z_]\w*", line):
            t = m.group(0).upper()
            if t not in KW and len(t) > 1 and not re.fullmatch(r"[EDF]\d*", t):
                out[t] += 1
    return out
fr, cr = {}, {}
for f in F:
    txt = open(f, errors="ignore").read()
    for m in re.finditer(r"^\s*SUBROUTINE\s+(\w+).*?^\s*END\s+SUBROUTINE\s+\1", txt, flags=re.S | re.M | re.I):
        fr[m.group(1).upper()] = idents(m.group(0), True)
for f in C:
    txt = open(f).read()
    for m in re.finditer(r"^(?:static\s+)?(?:inline\s+)?(?:void|int|float)\s+(\w+)\s*\([^;{]*\)\s*\{.*?^\}", txt, flags=re.S | re.M):
        cr.setdefault(m.group(1).upper(), collections.Counter()).update(idents(m.group(0), False))
want = sys.argv[1:] or sorted(fr)
for name in want:
    c = cr.get({"WTABLE_MMF_NOAHMP": "WTABLE"}.get(name, name))
    if c is None or name not in fr:
        continue
    a = fr[name]
    diffs = []
    for k in sorted(set(a) | set(c)):
        if a[k] != c[k] and (a[k] > 0 and c[k] > 0) and abs(a[k] - c[k]) >= 1:
            diffs.append(f"{k}:{a[k]}/{c[k]}")
    only_f = [k for k in a if k not in c]
    print(f"{name}: differ {' '.join(diffs)}\n    fortran-only: {' '.join(sorted(only_f))}")
