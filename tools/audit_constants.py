"""Line-audit aid: for every routine of the reference's physics, compare the multiset of numeric literals in the
Fortran text with the literals of the same-named oracle function (C++).  A literal present on one side only is a
candidate transcription error.  Needs /root/reference (build container only)."""
import collections, re, sys
F = ["/root/reference/phys/module_sf_noahmplsm.F90", "/root/reference/phys/module_sf_noahmp_glacier.F90",
     "/root/reference/phys/module_sf_noahmp_groundwater.F90"]
C = ["oracle/nmo_land1.cpp", "oracle/nmo_land2.cpp", "oracle/nmo_land3.cpp", "oracle/nmo_glacier.cpp",
     "oracle/nmo_groundwater.cpp"]
num = re.compile(r"(?<![A-Za-z_\d.])(\d+\.\d*(?:[EeDd][+-]?\d+)?|\.\d+(?:[EeDd][+-]?\d+)?|\d+[EeDd][+-]?\d+|\d+)(?:_\w+)?(?![A-Za-z_\d]|\.\d)")
def norm(t):
    t = t.lower().replace("d", "e").rstrip("f")
    try:
        v = float(t)
    except ValueError:
        return None
    return "%.6g" % v
def lits(text, fortran):
    out = collections.Counter()
    for line in text.splitlines():
        if fortran:
            line = line.split("!")[0]
        else:
            line = re.sub(r"(?<=[\d.])[fF]\b", "", line.split("//")[0])
        for m in num.finditer(line):
            n = norm(m.group(1))
            if n is not None and n not in ("0", "1", "2", "3", "4", "5", "6", "7"):
                out[n] += 1
    return out
def audit(repo="."):
    """{routine: (sorted fortran-only literals, sorted oracle-only literals)} for the routines that differ, and the
    list of reference routines without an oracle function of the same name."""
    import os
    fr = {}
    for f in F:
        txt = open(f, errors="ignore").read()
        for m in re.finditer(r"^\s*SUBROUTINE\s+(\w+).*?^\s*END\s+SUBROUTINE\s+\1", txt, flags=re.S | re.M | re.I):
            fr[m.group(1).upper()] = lits(m.group(0), True)
    cr = {}
    for f in C:
        txt = open(os.path.join(repo, f)).read()
        # crude function splitter: 'name(' at column 0..n followed by body until a line that is just '}'
        for m in re.finditer(r"^(?:static\s+)?(?:inline\s+)?(?:void|int|float)\s+(\w+)\s*\([^;{]*\)\s*\{(?:[^\n]*\}[ \t]*$|.*?^\})", txt, flags=re.S | re.M):
            cr.setdefault(m.group(1).upper(), collections.Counter()).update(lits(m.group(0), False))
    alias = {"WTABLE_MMF_NOAHMP": "WTABLE"}
    diff, missing = {}, []
    for name in sorted(fr):
        c = cr.get(alias.get(name, name))
        if c is None:
            c = cr.get(name.replace("_GLACIER", ""))
            if c is None:
                missing.append(name)
                continue
        a = fr[name]
        only_f = sorted(k for k in a if k not in c)
        only_c = sorted(k for k in c if k not in a)
        if only_f or only_c:
            diff[name] = (only_f, only_c)
    return diff, missing


if __name__ == "__main__":
    diff, missing = audit()
    for name in missing:
        print(f"-- {name}: no oracle function of that name")
    for name, (f_, c_) in diff.items():
        print(f"{name}: fortran-only {f_}  oracle-only {c_}")
