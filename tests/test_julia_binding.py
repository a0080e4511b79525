"""The Julia wrapper cannot be executed here (no Julia in the image); what can be checked without it is that every
`ccall` it makes binds a symbol declared in include/tfhe_b200.h with the same arity and the same integer widths, that
the exported API surface covers the reference's (src/TFHE.jl:24-61), and that the reference's broadcasting idioms
(docs/src/manual.md:28-35, lwe.jl:32, api.jl:103,130) have methods."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JULIA_DIR = os.path.join(ROOT, "tfhe.jl_b200", "julia")

C_TYPES = {
    "int": "i32", "void": "void", "size_t": "usize", "uint32_t": "u32", "int32_t": "i32", "uint64_t": "u64",
    "const char*": "cstr", "double*": "ptr_f64", "const int32_t*": "ptr_i32", "int32_t*": "ptr_i32", "const int*": "ptr_i32",
    "tfhe_b200_ctx*": "ptr_void", "const tfhe_b200_ctx*": "ptr_void", "tfhe_b200_multi*": "ptr_void",
    "const tfhe_b200_multi*": "ptr_void", "void*": "ptr_void", "tfhe_b200_ctx**": "ptr_ptr", "tfhe_b200_multi**": "ptr_ptr",
    "const tfhe_b200_params*": "ptr_params", "const uint8_t*": "ptr_u8", "uint8_t*": "ptr_u8", "double": "f64",
}
JL_TYPES = {
    "Cint": "i32", "Cvoid": "void", "Csize_t": "usize", "UInt32": "u32", "Int32": "i32", "UInt64": "u64", "Cstring": "cstr",
    "Ptr{Int32}": "ptr_i32", "Ptr{Cint}": "ptr_i32", "Ptr{Cvoid}": "ptr_void", "Ref{Ptr{Cvoid}}": "ptr_ptr",
    "Ref{CParams}": "ptr_params", "Ptr{Float64}": "ptr_f64", "Ref{Float64}": "ptr_f64", "Ref{Cdouble}": "ptr_f64",
    "Ptr{UInt8}": "ptr_u8", "Cdouble": "f64", "Float64": "f64",
}


def header_signatures():
    text = open(os.path.join(ROOT, "include", "tfhe_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    sigs = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ ]*?[\s\*]+)(tfhe_b200_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        def canon(t):
            t = re.sub(r"\s+", " ", t.strip())
            t = re.sub(r"\s*\*", "*", t)
            return C_TYPES[t]
        arg_types = []
        for a in args.split(","):
            a = a.strip()
            if a in ("void", ""):
                continue
            a = re.sub(r"\b[a-z_][a-z0-9_]*$", "", a).strip()          # drop the parameter name
            arg_types.append(canon(a))
        sigs[name] = (canon(ret), arg_types)
    return sigs


def split_top_level(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "{(":
            depth += 1
        elif ch in "})":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def julia_ccalls(path):
    text = open(path).read()
    text = re.sub(r"#[^\n]*", "", text)
    calls = []
    for m in re.finditer(r"ccall\(\(\s*:([a-z0-9_]+)\s*,\s*LIB\s*\)\s*,\s*([A-Za-z0-9{}]+)\s*,\s*\(", text):
        i = m.end(); depth = 1; j = i
        while depth:
            depth += {"(": 1, ")": -1}.get(text[j], 0); j += 1
        types = split_top_level(text[i:j - 1])
        # the call arguments follow, up to the closing parenthesis of ccall
        k = j; depth = 1
        while depth:
            depth += {"(": 1, ")": -1, "[": 1, "]": -1}.get(text[k], 0); k += 1
        nargs = len(split_top_level(text[j:k - 1].lstrip(", \n")))
        calls.append((m.group(1), m.group(2), types, nargs))
    return calls


def test_every_julia_ccall_matches_the_header():
    sigs = header_signatures()
    assert len(sigs) >= 40
    seen = set()
    for fn in sorted(os.listdir(JULIA_DIR)):
        if not fn.endswith(".jl"):
            continue
        for name, ret, types, nargs in julia_ccalls(os.path.join(JULIA_DIR, fn)):
            assert name in sigs, f"{fn}: ccall of {name}, which include/tfhe_b200.h does not declare"
            cret, cargs = sigs[name]
            assert JL_TYPES[ret] == cret, f"{fn}: {name} returns {cret} in C, {ret} in Julia"
            assert len(types) == len(cargs) == nargs, f"{fn}: {name} takes {len(cargs)} arguments, ccall lists {len(types)} types and passes {nargs}"
            for i, (jt, ct) in enumerate(zip(types, cargs)):
                assert JL_TYPES[jt] == ct, f"{fn}: {name} argument {i}: C {ct}, Julia {jt}"
            seen.add(name)
    # the wrapper reaches the whole gate path, single- and multi-device, single-key and MK
    for must in ("tfhe_b200_create", "tfhe_b200_multi_create", "tfhe_b200_gate_batch", "tfhe_b200_multi_gate_batch",
                 "tfhe_b200_mk_nand_batch", "tfhe_b200_multi_mk_nand_batch", "tfhe_b200_mk_bootstrap_batch",
                 "tfhe_b200_bootstrap_batch", "tfhe_b200_polymul_batch", "tfhe_b200_load_bk", "tfhe_b200_mk_load_ksk"):
        assert must in seen, must


def test_julia_wrapper_keeps_the_reference_api_surface():
    src = open(os.path.join(JULIA_DIR, "TFHEB200.jl")).read()
    exported = set(re.findall(r"[A-Za-z_0-9]+", " ".join(re.findall(r"^export (.*)$", src, flags=re.M))))
    # src/TFHE.jl:24-61 of the reference
    reference = {"make_key_pair", "LweSample", "SecretKey", "CloudKey", "encrypt", "decrypt", "tfhe_parameters_80",
                 "tfhe_parameters_128", "SchemeParameters", "gate_nand", "gate_or", "gate_and", "gate_xor", "gate_xnor",
                 "gate_not", "gate_constant", "gate_nor", "gate_andny", "gate_andyn", "gate_orny", "gate_oryn", "gate_mux",
                 "SharedKey", "CloudKeyPart", "MKCloudKey", "mk_encrypt", "mk_decrypt", "mk_gate_nand",
                 "mktfhe_parameters_2party", "mktfhe_parameters_4party", "mktfhe_parameters_8party"}
    assert reference <= exported, reference - exported
    # lwe.jl:21-29: the sample carries params, a, b, current_variance; lwe.jl:32 / api.jl:103,130: keys and samples broadcast as scalars
    body = re.search(r"mutable struct LweSample\n(.*?)\nend", src, flags=re.S).group(1)
    for field in ("params::LweParams", "a::Array{Torus32, 1}", "b::Torus32", "current_variance::Float64"):
        assert field in body, field
    for t in ("lwe::LweSample", "sk::SecretKey", "ck::CloudKey"):
        assert f"Base.Broadcast.broadcastable({t}) = (" in src, t
    # one library call per broadcast over vectors of samples (docs/src/manual.md:28-35)
    assert "Base.Broadcast.broadcasted(::typeof($f), ck::CloudKey, x::SampleVec, y::SampleVec)" in src
    assert "Base.Broadcast.broadcasted(::typeof(encrypt), rng::Base.RefValue{<:AbstractRNG}, key::SecretKey" in src
    assert "Base.Broadcast.broadcasted(::typeof(decrypt), key::SecretKey, xs::SampleVec)" in src
    # all arithmetic on torus words stays in Int32 (the round-1 advisor finding: `sum` over Int32 widens to Int64)
    assert not re.search(r"\bsum\(prod", src)
