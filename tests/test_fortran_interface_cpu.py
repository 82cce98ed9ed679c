"""No Fortran compiler exists in this image, so fortran/fedem_b200_mod.f90 is checked the only way left: both the C header
and the interface blocks are parsed and compared entry point by entry point -- number of arguments, pass-by-value vs
pass-by-reference, C type vs ISO_C_BINDING kind, function result.  A mismatch here would be a silent stack / register
corruption the day the module meets gfortran."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _c_prototypes():
    txt = open(os.path.join(ROOT, "include", "fedem_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", " ", txt, flags=re.S)
    txt = re.sub(r"#.*", " ", txt)
    txt = re.sub(r"typedef\s+struct\s+\w+\s*\{.*?\}\s*\w+\s*;", " ", txt, flags=re.S)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(\w+)\s*\(([^()]*)\)\s*;", txt):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith("typedef") or not ret:
            continue
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                a = re.sub(r"\bconst\b", "", a).strip()
                is_ptr = "*" in a or "[" in a
                base = re.sub(r"[\*\[\]\d]", " ", a).split()
                # drop the parameter name (last token) unless the declaration is type-only
                ctype = " ".join(base[:-1]) if len(base) > 1 else base[0]
                nstar = a.count("*")
                params.append((ctype, is_ptr, nstar))
        protos[name] = (" ".join(re.sub(r"\bconst\b", "", ret).split()), params)
    return protos


def _f_interfaces():
    src = open(os.path.join(ROOT, "fortran", "fedem_b200_mod.f90")).read()
    src = re.sub(r"&\s*\n\s*&?", " ", src)            # continuation lines
    out = {}
    pat = re.compile(r"^\s*(function|subroutine)\s+(\w+)\s*\(([^)]*)\)\s*bind\s*\(\s*C\s*,\s*name\s*=\s*[\"'](\w+)[\"']\s*\)(?:\s*result\s*\((\w+)\))?(.*?)^\s*end\s+\1",
                     re.S | re.M | re.I)
    for m in pat.finditer(src):
        kind, fname, args, cname, res, body = m.groups()
        names = [a.strip().lower() for a in args.split(",") if a.strip()]
        decl = {}
        for line in body.splitlines():
            line = line.split("!")[0].strip()
            if "::" not in line or line.lower().startswith("import"):
                continue
            left, right = line.split("::", 1)
            left = left.lower()
            by_value = bool(re.search(r"\bvalue\b", left))
            typ = left.split(",")[0].strip()
            for v in re.split(r",(?![^()]*\))", right):
                v = v.strip().lower()
                vn = re.match(r"\w+", v).group(0)
                decl[vn] = (typ.replace(" ", ""), by_value, "(" in v)
        out[cname] = (kind.lower(), names, decl, res.lower() if res else None)
    return out


SCALAR = {"int": "integer(c_int)", "double": "real(c_double)", "long long": "integer(c_long_long)", "size_t": "integer(c_size_t)",
          "unsigned": "integer(c_int)", "bool": "logical(c_bool)"}


def test_every_interface_block_matches_its_c_prototype():
    protos, ifs = _c_prototypes(), _f_interfaces()
    assert len(protos) > 100 and len(ifs) > 100
    problems = []
    for name, (ret, params) in sorted(protos.items()):
        if name not in ifs:
            problems.append(f"{name}: no interface block")
            continue
        kind, names, decl, res = ifs[name]
        if len(names) != len(params):
            problems.append(f"{name}: {len(params)} C parameters, {len(names)} Fortran dummy arguments")
            continue
        # result
        if ret == "void":
            if kind != "subroutine":
                problems.append(f"{name}: void in C but a Fortran function")
        else:
            if kind != "function":
                problems.append(f"{name}: returns {ret} in C but is a Fortran subroutine")
            else:
                rt = decl.get(res or name.lower(), ("?", False, False))[0]
                want = "type(c_ptr)" if "*" in ret else SCALAR.get(ret.replace("*", "").strip(), "?")
                if rt != want:
                    problems.append(f"{name}: result {ret} vs {rt}")
        for (ctype, is_ptr, nstar), an in zip(params, names):
            if an not in decl:
                problems.append(f"{name}: dummy argument {an} has no declaration")
                continue
            typ, by_value, is_array = decl[an]
            if not is_ptr:                       # C scalar by value
                want = SCALAR.get(ctype)
                if want is None:
                    problems.append(f"{name}: unexpected C scalar type '{ctype}' for {an}")
                elif typ != want or not by_value:
                    problems.append(f"{name}: {an} is '{ctype}' by value in C, Fortran has {typ}{', value' if by_value else ' by reference'}")
            else:                                # C pointer: Fortran by reference, or type(c_ptr) by value
                if typ == "type(c_ptr)":
                    if by_value and nstar == 2 and ctype not in ("char",):
                        problems.append(f"{name}: {an} is '{ctype}**' in C but type(c_ptr), value in Fortran (needs by reference)")
                    if not by_value and nstar == 1:
                        problems.append(f"{name}: {an} is '{ctype}*' in C but type(c_ptr) by reference (= {ctype}**) in Fortran")
                    continue
                if by_value:
                    problems.append(f"{name}: {an} is a pointer in C but passed by value as {typ}")
                    continue
                want = {"int": "integer(c_int)", "double": "real(c_double)", "char": "character(kind=c_char)", "long long": "integer(c_long_long)",
                        "void": None, "float": "real(c_float)"}.get(ctype, f"type({ctype})")
                if want is not None and typ != want:
                    problems.append(f"{name}: {an} is '{ctype}*' in C, {typ} in Fortran")
    assert not problems, "\n".join(problems)


def test_derived_types_mirror_the_c_structs():
    hdr = open(os.path.join(ROOT, "include", "fedem_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    f90 = open(os.path.join(ROOT, "fortran", "fedem_b200_mod.f90")).read()
    for sname in ("fsr_sam", "fsr_elmdata", "fsr_options", "fsr_rosette", "fsr_rdb_options"):
        body = re.search(r"typedef\s+struct\s+%s\s*\{(.*?)\}\s*%s\s*;" % (sname, sname), hdr, re.S).group(1)
        cmembers = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            decl = re.sub(r"\bconst\b", "", decl)
            is_ptr = "*" in decl
            base = decl.replace("*", " ").split()
            ctype = base[0] if base[0] != "unsigned" else "unsigned"
            for v in " ".join(base[1:]).split(","):
                v = v.strip()
                dim = re.search(r"\[(\d+)\]", v)
                cmembers.append((re.sub(r"\[.*", "", v).strip().lower(), "ptr" if is_ptr else ctype, int(dim.group(1)) if dim else 0))
        fb = re.search(r"type\s*,\s*bind\(C\)\s*::\s*%s(.*?)end\s+type" % sname, f90, re.S | re.I).group(1)
        fmembers = []
        for line in fb.splitlines():
            line = line.split("!")[0]
            if "::" not in line:
                continue
            typ, names = line.split("::")
            typ = typ.strip().lower().replace(" ", "")
            for v in re.split(r",(?![^()]*\))", names):
                v = v.strip().lower()
                dim = re.search(r"\((\d+)\)", v)
                kind = {"integer(c_int)": "int", "real(c_double)": "double", "type(c_ptr)": "ptr"}[typ]
                fmembers.append((re.match(r"\w+", v).group(0), kind, int(dim.group(1)) if dim else 0))
        cm = [(n, "int" if t == "unsigned" else t, d) for n, t, d in cmembers]
        assert cm == fmembers, (sname, cm, fmembers)
