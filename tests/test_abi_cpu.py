"""CPU (-m "not gpu"): the ctypes binding (clipdlm/_lib.py) and include/clipdlm.h must describe the SAME ABI.

Two independent checks, neither launches a kernel:
  * struct layouts: a C program compiled by gcc against the header prints sizeof / offsetof / field size of every struct member;
    the ctypes Structures must agree byte for byte (a drifted field silently shifts every pointer behind it);
  * prototypes: every `clipdlm_*` declaration of the header is parsed and its return / argument classes (pointer, i32, i64, u32, u64,
    f32, f64, size_t) are compared with the argtypes / restype the binding installs.
"""
import ctypes as C
import os
import re
import subprocess

import pytest

from _util import ROOT

HEADER = os.path.join(ROOT, "include", "clipdlm.h")


def _structs():
    from clipdlm import _lib as L
    return {"clipdlm_bf_t": L.Bf, "clipdlm_gemm_t": L.Gemm, "clipdlm_embed_t": L.Embed, "clipdlm_config_t": L.Config,
            "clipdlm_buffers_t": L.Buffers, "clipdlm_pass_t": L.Pass, "clipdlm_loss_cfg_t": L.LossCfg, "clipdlm_dp_buffers_t": L.DpBuffers,
            "clipdlm_prof_t": L.Prof}


def test_every_header_struct_has_a_binding():
    header = open(HEADER).read()
    declared = set(re.findall(r"\}\s*(clipdlm_[a-z0-9_]+_t)\s*;", header))
    assert declared == set(_structs()), declared ^ set(_structs())


def test_struct_layouts_match_the_header(tmp_path):
    structs = _structs()
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "clipdlm.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} . %zu 0\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu %zu\\n", offsetof({cname}, {fname}), sizeof((({cname}*)0)->{fname}));')
    lines += ["  return 0;", "}"]
    src, exe = tmp_path / "layout.c", tmp_path / "layout"
    src.write_text("\n".join(lines))
    r = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr   # a field the binding names but the header lacks fails HERE
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    seen = 0
    for row in out.splitlines():
        cname, fname, a, b = row.split()
        cls = structs[cname]
        if fname == ".":
            assert C.sizeof(cls) == int(a), f"sizeof({cname}): header {a}, ctypes {C.sizeof(cls)}"
        else:
            f = getattr(cls, fname)
            assert (f.offset, f.size) == (int(a), int(b)), f"{cname}.{fname}: header (offset {a}, size {b}), ctypes ({f.offset}, {f.size})"
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in structs.values())


def test_header_struct_field_counts_match():
    """The layout test walks the BINDING's field list; a field added at the END of a header struct would escape it when the struct's
    size did not change through padding. Count the declarators of every header struct as well."""
    header = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for cname, cls in _structs().items():
        m = re.search(r"typedef\s+struct\s+\w+\s*\{([^{}]*)\}\s*" + cname + r"\s*;", header, flags=re.S)
        assert m, cname
        n = 0
        for decl in m.group(1).split(";"):
            decl = decl.strip()
            if decl:
                n += decl.count(",") + 1
        assert n == len(cls._fields_), f"{cname}: header declares {n} members, the binding {len(cls._fields_)}"


_SCALARS = {"int": "i32", "int32_t": "i32", "uint32_t": "u32", "int64_t": "i64", "uint64_t": "u64", "float": "f32", "double": "f64",
            "size_t": "size", "clipdlm_stream": "ptr", "void": "void"}


def _c_class(decl: str) -> str:
    decl = re.sub(r"\b(const|struct)\b", " ", decl).strip()
    if "*" in decl or "[" in decl:
        return "ptr"
    ty = decl.split()[0]
    assert ty in _SCALARS, f"unknown C type in header prototype: {decl!r}"
    return _SCALARS[ty]


def _ctypes_class(t) -> str:
    if t is None:
        return "void"
    if t in (C.c_void_p, C.c_char_p) or isinstance(t, type) and issubclass(t, C._Pointer):
        return "ptr"
    table = {C.c_int32: "i32", C.c_uint32: "u32", C.c_int64: "i64", C.c_uint64: "u64", C.c_float: "f32", C.c_double: "f64", C.c_size_t: "size"}
    # c_int is c_int32 and c_size_t is c_uint64 on this ABI: ctypes aliases them, so classify by width where the alias collapsed
    if t in table:
        return table[t]
    raise AssertionError(f"unclassified ctypes type {t}")


def test_prototypes_match_the_binding():
    from clipdlm import _lib as L
    header = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    protos = re.findall(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(clipdlm_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", header)
    found = {}
    for ret, name, args in protos:
        args = args.strip()
        arg_classes = [] if args in ("", "void") else [_c_class(a) for a in args.split(",")]
        found[name] = (_c_class(ret + " x") if ret.strip() else "void", arg_classes)
    assert set(found) == set(L._SIGS), set(found) ^ set(L._SIGS)
    same = {"size": "u64"}  # ctypes collapses c_size_t into c_uint64 on LP64
    for name, (res, argtypes) in L._SIGS.items():
        c_ret, c_args = found[name]
        got_ret, got_args = _ctypes_class(res), [_ctypes_class(a) for a in argtypes]
        norm = lambda xs: [same.get(x, x) for x in xs]
        assert norm([c_ret]) == norm([got_ret]), f"{name}: returns {c_ret} in the header, {got_ret} in the binding"
        assert norm(c_args) == norm(got_args), f"{name}: header {c_args}\n binding {got_args}"


def test_enums_match_the_binding(tmp_path):
    from clipdlm import _lib as L
    names = {"CLIPDLM_EPI_STORE": L.EPI_STORE, "CLIPDLM_EPI_WGRAD": L.EPI_WGRAD, "CLIPDLM_EPI_LSE": L.EPI_LSE, "CLIPDLM_EPI_SMGRAD": L.EPI_SMGRAD,
             "CLIPDLM_EPI_LSE_EXP": L.EPI_LSE_EXP, "CLIPDLM_EPI_STORE_ROWSCALE": L.EPI_STORE_ROWSCALE,
             "CLIPDLM_OPT_FUSED_SOFTMAX_GRAD": L.OPT_FUSED_SOFTMAX_GRAD, "CLIPDLM_OPT_EXP_SHIFT_PTR": L.OPT_EXP_SHIFT_PTR,
             "CLIPDLM_EPI_STORE_GELU_DERIV": L.EPI_STORE_GELU_DERIV, "CLIPDLM_EPI_STORE_MULAUX": L.EPI_STORE_MULAUX,
             "CLIPDLM_OPT_GELU_DERIV_STORE": L.OPT_GELU_DERIV_STORE,
             "CLIPDLM_P_POS": L.P_POS, "CLIPDLM_P_SEG": L.P_SEG, "CLIPDLM_P_LAYER0": L.P_LAYER0, "CLIPDLM_PL_QKV_W": L.PL_QKV_W,
             "CLIPDLM_PL_LN2_B": L.PL_LN2_B, "CLIPDLM_P_PER_LAYER": L.P_PER_LAYER, "CLIPDLM_PROF_NCAT": len(L.PROF_CATEGORIES),
             "CLIPDLM_MAX_PEERS": L.MAX_PEERS}
    for i, cat in enumerate(L.PROF_CATEGORIES):
        names["CLIPDLM_PROF_" + cat.upper()] = i
    body = "\n".join(f'  printf("{n} %d\\n", (int){n});' for n in names)
    src, exe = tmp_path / "enums.c", tmp_path / "enums"
    src.write_text('#include <stdio.h>\n#include "clipdlm.h"\nint main(void) {\n' + body + "\n  return 0;\n}\n")
    r = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for row in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines():
        n, v = row.split()
        assert names[n] == int(v), f"{n}: header {v}, binding {names[n]}"
