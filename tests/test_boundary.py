"""Pin the drop-in boundary without a Fortran compiler: the ISO_C_BINDING shim
(integration/module_sf_noahmpdrv_b200.F90) is parsed and compared

  * member for member with the C structs of include/noahmp_b200.h (names, order, C types) and with the ctypes mirror
    (noahmp_b200/_capi.py) the tests call through;
  * dummy for dummy (names, order, rank and bounds of every array, INTENT) with the reference's own
    `SUBROUTINE noahmplsm` (phys/module_sf_noahmpdrv.F90:11-44, decls :51-211), `NOAHMP_INIT` (:847-864) and
    `WTABLE_mmf_noahmp` (phys/module_sf_noahmp_groundwater.F90:14-22) when /root/reference is present (it is in the
    build container; the GPU box skips those cases);
  * INTERFACE block by INTERFACE block with the prototypes of the header (symbol exists, argument count, by-value /
    by-reference passing).
"""
import os
import re

import pytest

from noahmp_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "integration", "module_sf_noahmpdrv_b200.F90")
HEADER = os.path.join(ROOT, "include", "noahmp_b200.h")
REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "phys")), reason="reference tree not present")


# ---------------------------------------------------------------------------------------------------------------------
# Fortran (free form) helpers
def fortran_statements(text, defines=()):
    """Logical statements: comments stripped, continuation lines joined, `#ifdef X ... #endif` blocks dropped unless X
    is in `defines`, `;` split."""
    out, cur, skip = [], "", []
    for raw in text.splitlines():
        s = raw.strip()
        if s.startswith("#"):
            if s.startswith("#ifdef"):
                skip.append(s.split()[1] not in defines)
            elif s.startswith("#ifndef"):
                skip.append(s.split()[1] in defines)
            elif s.startswith("#else") and skip:
                skip[-1] = not skip[-1]
            elif s.startswith("#endif") and skip:
                skip.pop()
            continue
        if any(skip):
            continue
        line, q = "", None
        for ch in raw:  # strip the comment, minding quotes
            if q:
                line += ch
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
                line += ch
            elif ch == "!":
                break
            else:
                line += ch
        line = line.strip()
        if not line:
            continue
        if line.startswith("&"):
            line = line[1:].lstrip()
        if line.endswith("&"):
            cur += line[:-1] + " "
            continue
        cur += line
        for st in cur.split(";"):
            if st.strip():
                out.append(st.strip())
        cur = ""
    return out


def find_subroutine(stmts, name):
    """(dummy list, declaration statements) of SUBROUTINE `name`."""
    pat = re.compile(r"^\s*SUBROUTINE\s+" + re.escape(name) + r"\s*\((.*)\)\s*$", re.I)
    for k, s in enumerate(stmts):
        m = pat.match(s)
        if m:
            dummies = [d.strip().lower() for d in m.group(1).split(",") if d.strip()]
            body = []
            for t in stmts[k + 1:]:
                if re.match(r"^\s*END\s+SUBROUTINE", t, re.I):
                    break
                body.append(t)
            return dummies, body
    raise AssertionError(f"SUBROUTINE {name} not found")


def split_top(s):
    """split on commas that are not inside parentheses"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [x.strip() for x in out]


def declarations(body, dummies):
    """name -> dict(type, intent, dims, optional) for every dummy argument declared in `body`."""
    want = set(dummies)
    decl = {}
    for s in body:
        m = re.match(r"^\s*(REAL|INTEGER|LOGICAL|CHARACTER)\b(.*?)::(.*)$", s, re.I)
        if not m:
            continue
        typ, attrs, names = m.group(1).upper(), m.group(2), m.group(3)
        intent = re.search(r"INTENT\s*\(\s*(\w+)\s*\)", attrs, re.I)
        dim = re.search(r"DIMENSION\s*\((.*?)\)\s*(,|$)", attrs.strip() + ",", re.I)
        adim = None
        if dim:
            # re-extract with balanced parentheses
            i = attrs.upper().index("DIMENSION")
            j = attrs.index("(", i)
            depth, k = 0, j
            while True:
                depth += attrs[k] == "("
                depth -= attrs[k] == ")"
                if depth == 0:
                    break
                k += 1
            adim = attrs[j + 1:k]
        for item in split_top(names):
            mm = re.match(r"^(\w+)\s*(\((.*)\))?$", item.strip())
            if not mm:
                continue
            n = mm.group(1).lower()
            if n not in want:
                continue
            dims = mm.group(3) if mm.group(3) is not None else adim
            decl[n] = dict(type=typ, intent=intent.group(1).upper() if intent else None,
                           dims=norm_dims(dims), optional=bool(re.search(r"\bOPTIONAL\b", attrs, re.I)))
    return decl


def norm_dims(d):
    if d is None:
        return None
    parts = []
    for x in split_top(d):
        x = re.sub(r"\s+", "", x).lower()
        if ":" not in x:
            x = "1:" + x
        parts.append(x)
    return tuple(parts)


def bindc_type(stmts, name):
    """[(member, fortran type string)] of TYPE, BIND(C) :: name"""
    out, on = [], False
    for s in stmts:
        if re.match(r"^\s*TYPE\s*,\s*BIND\s*\(\s*C\s*\)\s*::\s*" + name + r"\s*$", s, re.I):
            on = True
            continue
        if on and re.match(r"^\s*END\s+TYPE", s, re.I):
            return out
        if on:
            m = re.match(r"^\s*(.+?)\s*::\s*(.+)$", s)
            for n in split_top(m.group(2)):
                out.append((n.strip().lower(), re.sub(r"\s+", "", m.group(1)).upper()))
    raise AssertionError(f"TYPE {name} not found")


# ---------------------------------------------------------------------------------------------------------------------
# C header helpers
def c_struct(text, name):
    """[(member, kind)] with kind in {'i', 'f', 'pf', 'pi'} of `typedef struct name {...} name;`"""
    m = re.search(r"typedef\s+struct\s+" + name + r"\s*\{(.*?)\}\s*" + name + r"\s*;", text, re.S)
    assert m, name
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    out = []
    for st in body.split(";"):
        st = " ".join(st.split())
        if not st:
            continue
        mm = re.match(r"^(const\s+)?(int32_t|float|int)\s*(\*?)\s*(.*)$", st)
        assert mm, st
        base, first_ptr = mm.group(2), mm.group(3)
        for k, item in enumerate(mm.group(4).split(",")):
            item = item.strip()
            ptr = item.startswith("*") or (k == 0 and first_ptr == "*")
            n = item.lstrip("* ").strip()
            kind = ("p" if ptr else "") + ("f" if base == "float" else "i")
            out.append((n, kind))
    return out


def c_prototypes(text):
    """name -> list of parameter strings, for every noahmp_b200_* function the header declares"""
    body = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(noahmp_b200_\w+)\s*\(([^;{}]*?)\)\s*;", body, re.S):
        params = " ".join(m.group(2).split())
        protos[m.group(1)] = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
    return protos


FKIND = {"i": "INTEGER(C_INT)", "f": "REAL(C_FLOAT)", "pf": "TYPE(C_PTR)", "pi": "TYPE(C_PTR)"}


@pytest.fixture(scope="module")
def shim():
    return fortran_statements(open(SHIM).read())


@pytest.fixture(scope="module")
def header():
    return open(HEADER).read()


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cname,spec", [("noahmp_lsm_args", "ARGS_SPEC"), ("noahmp_wtable_args", "WT_SPEC"),
                                        ("noahmp_init_args", "INIT_SPEC")])
def test_bindc_types_mirror_the_header_structs(shim, header, cname, spec):
    cs = c_struct(header, cname)
    ft = bindc_type(shim, cname)
    assert [n for n, _ in ft] == [n for n, _ in cs], "member names / order differ between shim and header"
    for (n, ftype), (_, kind) in zip(ft, cs):
        assert ftype == FKIND[kind], (cname, n, ftype, kind)
    # and the ctypes mirror the tests call through
    py = list(getattr(_capi, spec))
    assert [(n, k) for n, k in py] == cs, f"_capi.{spec} differs from the header"


def test_small_structs(shim, header):
    assert [n for n, _ in bindc_type(shim, "noahmp_status")] == ["code", "i", "j", "count", "value"]
    assert [n for n, _ in c_struct(header, "noahmp_status")] == ["code", "i", "j", "count", "value"]
    assert [n for n, _ in bindc_type(shim, "noahmp_forcing_fields")] == _capi.FORCING_FIELDS
    assert [n for n, _ in c_struct(header, "noahmp_forcing_fields")] == _capi.FORCING_FIELDS


def test_shim_fills_every_member_from_the_dummy_of_the_same_name(shim):
    """`a%x = X` / `a%x = C_LOC(X)` for every member of the three argument structs."""
    for sub, typ in (("noahmplsm", "noahmp_lsm_args"), ("WTABLE_mmf_noahmp", "noahmp_wtable_args"),
                     ("NOAHMP_INIT", "noahmp_init_args")):
        dummies, body = find_subroutine(shim, sub)
        members = [n for n, _ in bindc_type(shim, typ)]
        assigned = {}
        for s in body:
            m = re.match(r"^(?:IF\s*\(.*?\)\s*)?a%(\w+)\s*=\s*(.+)$", s.strip(), re.I)
            if m:
                assigned.setdefault(m.group(1).lower(), []).append(m.group(2))
        for n in members:
            assert n in assigned, (sub, "member never assigned", n)
            src = " ".join(assigned[n]).lower()
            assert re.search(r"\b" + n + r"\b", src), (sub, n, src)
            assert n in dummies, (sub, n)


def _compare_with_reference(shim, sub, ref_file, widen_out=()):
    ref = fortran_statements(open(os.path.join(REF, ref_file)).read())
    d_ref, b_ref = find_subroutine(ref, sub)
    d_shim, b_shim = find_subroutine(shim, sub)
    assert d_shim == d_ref, f"{sub}: dummy list differs from {ref_file}"
    dr, ds = declarations(b_ref, d_ref), declarations(b_shim, d_shim)
    for n in d_ref:
        assert n in dr, (sub, "reference declaration not parsed", n)
        assert n in ds, (sub, "shim does not declare", n)
        r, s = dr[n], ds[n]
        assert r["type"] == s["type"], (sub, n, r, s)
        assert r["dims"] == s["dims"], (sub, n, r["dims"], s["dims"])
        assert r["optional"] == s["optional"], (sub, n)
        if r["intent"] != s["intent"]:
            # the library reads what the caller holds in OUT arrays at open-water cells (which the reference never
            # writes) to hand it back unchanged: INTENT(OUT) dummies are INTENT(INOUT) in the shim, nothing else differs
            assert (r["intent"], s["intent"]) == ("OUT", "INOUT") or n in widen_out, (sub, n, r["intent"], s["intent"])
    return d_ref, dr


@needs_ref
def test_noahmplsm_dummy_list_equals_the_reference(shim, header):
    d, dr = _compare_with_reference(shim, "noahmplsm", "phys/module_sf_noahmpdrv.F90")
    assert len(d) == 158
    # and the C struct follows the same order
    assert [n for n, _ in c_struct(header, "noahmp_lsm_args")] == d
    # kinds: REAL arrays -> float*, INTEGER arrays -> int32_t*, scalars by value
    for (n, kind) in c_struct(header, "noahmp_lsm_args"):
        r = dr[n]
        assert kind == ("p" if r["dims"] else "") + ("f" if r["type"] == "REAL" else "i"), (n, kind, r)
    # INTENT groups of the header comment: IN 41, INOUT 55, OUT 44, bounds 18
    intents = [dr[n]["intent"] for n in d]
    assert intents.count("OUT") == 44 and intents.count("INOUT") == 55


@needs_ref
def test_noahmp_init_dummy_list_equals_the_reference(shim, header):
    # MMINLU is not in the struct (the tables were read by noahmp_b200_start); LOGICALs travel as int32
    d, dr = _compare_with_reference(shim, "NOAHMP_INIT", "phys/module_sf_noahmpdrv.F90", widen_out=("tmn",))
    cs = c_struct(header, "noahmp_init_args")
    assert [n for n, _ in cs] == [n for n in d if n != "mminlu"]
    for n, kind in cs:
        r = dr[n]
        base = "f" if r["type"] == "REAL" else "i"  # LOGICAL, INTEGER -> int32
        assert kind == ("p" if (r["dims"] or n == "stepwtd") else "") + base, (n, kind, r)


@needs_ref
def test_wtable_dummy_list_equals_the_reference(shim, header):
    d, dr = _compare_with_reference(shim, "WTABLE_mmf_noahmp", "phys/module_sf_noahmp_groundwater.F90")
    cs = c_struct(header, "noahmp_wtable_args")
    assert [n for n, _ in cs] == d
    for n, kind in cs:
        r = dr[n]
        assert kind == ("p" if r["dims"] else "") + ("f" if r["type"] == "REAL" else "i"), (n, kind, r)
    # the call site hands the arrays over in this order (driver/module_hrldas_noahmp_driver.F90:424-436)
    drv = fortran_statements(open(os.path.join(REF, "driver/module_hrldas_noahmp_driver.F90")).read())
    call = [s for s in drv if re.match(r"^call\s+WTABLE_MMF_NOAHMP", s, re.I)]
    assert len(call) == 1 and len(split_top(call[0][call[0].index("(") + 1:call[0].rindex(")")])) == len(d)


def test_interface_blocks_match_the_header_prototypes(shim, header):
    protos = c_prototypes(header)
    seen = set()
    k = 0
    while k < len(shim):
        m = re.match(r"^\s*(FUNCTION|SUBROUTINE)\s+(\w+)\s*\((.*?)\)\s*BIND\s*\(\s*C\s*,\s*NAME\s*=\s*\"(\w+)\"\s*\)", shim[k], re.I)
        if not m:
            k += 1
            continue
        fargs = [a.strip().lower() for a in m.group(3).split(",") if a.strip()]
        cname = m.group(4)
        assert cname in protos, f"shim binds {cname}, which include/noahmp_b200.h does not declare"
        seen.add(cname)
        cparams = protos[cname]
        assert len(fargs) == len(cparams), (cname, fargs, cparams)
        body = []
        k += 1
        while not re.match(r"^\s*END\s+(FUNCTION|SUBROUTINE)", shim[k], re.I):
            body.append(shim[k])
            k += 1
        decl = {}
        for s in body:
            mm = re.match(r"^\s*(.+?)\s*::\s*(.+)$", s)
            if mm and not s.upper().startswith("IMPORT"):
                for n in split_top(mm.group(2)):
                    decl[re.sub(r"\(.*\)", "", n).strip().lower()] = re.sub(r"\s+", "", mm.group(1)).upper()
        for fa, cp in zip(fargs, cparams):
            assert fa in decl, (cname, "argument without declaration", fa)
            t = decl[fa]
            by_value = ",VALUE" in t
            if "*" in cp:
                # pointer parameter: a C_PTR by value, or anything else by reference
                assert (t.startswith("TYPE(C_PTR)") and by_value) or not by_value, (cname, fa, t, cp)
            else:
                assert by_value, (cname, fa, "scalar C parameter must be passed BY VALUE", t, cp)
                want = "REAL(C_FLOAT)" if cp.split()[0] == "float" else "INTEGER(C_INT)"
                assert t.startswith(want), (cname, fa, t, cp)
    # the entry points INTEGRATION.md tells a maintainer to call from Fortran are all bound
    need = {"noahmp_b200_read_tables", "noahmp_b200_create", "noahmp_b200_destroy", "noahmp_b200_set_mode",
            "noahmp_b200_set_fetch", "noahmp_b200_set_push", "noahmp_b200_set_forcing_hints", "noahmp_b200_noahmplsm",
            "noahmp_b200_sync_host", "noahmp_b200_init", "noahmp_b200_output_begin", "noahmp_b200_output_wait",
            "noahmp_b200_wtable", "noahmp_b200_wtable_begin", "noahmp_b200_wtable_exchange", "noahmp_b200_wtable_end",
            "noahmp_b200_wtable_sync_host", "noahmp_b200_comm_unique_id", "noahmp_b200_comm_init",
            "noahmp_b200_budget_enable", "noahmp_b200_budget_read", "noahmp_b200_forcing_static",
            "noahmp_b200_forcing_upload", "noahmp_b200_forcing_swap", "noahmp_b200_forcing_apply",
            "noahmp_b200_noahmplsm_device_forcing", "noahmp_b200_get_status"}
    assert need <= seen, sorted(need - seen)


def test_tables_size_constant_of_the_shim(shim):
    import ctypes as C
    m = [s for s in shim if "NOAHMP_TABLES_BYTES" in s and "PARAMETER" in s.upper()]
    assert m and int(re.search(r"=\s*(\d+)", m[0]).group(1)) == C.sizeof(_capi.NoahmpTables)
