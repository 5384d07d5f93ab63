"""CPU-only, mechanical check of the Julia host layer (julia/B200Backend/src/B200BackendModule.jl) against the C ABI and
against the reference's backend interface.  There is no Julia toolchain here, so nothing executes the .jl file; this
test READS it:

  1. every `ccall` / `@b200` names a symbol declared in include/b200ka.h, with the same arity, a compatible return type and
     compatible argument types (pointer <-> Ptr/Ref/Cstring, int32_t <-> Int32, int64_t <-> Int64, size_t <-> Csize_t, ...);
  2. the by-value records (LaunchDesc, CArg) and the status constants match the header field for field;
  3. every method of SURVEY.md Appendix C (the `src/pocl/backend.jl` template) has a definition with a matching signature;
  4. every host-side name the reference's testsuite files and examples call on a backend (tests/golden/api_calls.json, frozen
     by tests/golden/make_api_calls.py) is either defined here or listed as inherited from KA with the reason;
  5. the expansion of `KI.@kernel backend ... f(args...)` (reference src/intrinsics.jl:352-360) type-checks against the
     methods defined: argconvert(backend, f::Function) must hand the function back (round 1 shipped `error(...)` there, which
     killed examples/histogram.jl and examples/performant_matmul.jl on the real host layer), kernel_function must accept it and
     derive the name with nameof, and the launch must accept the ORIGINAL arguments;
  6. runtime state is task local (reference src/pocl/pocl.jl:12-41): every checked call binds device + stream first.
"""
import json
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200ka.h")
JL = os.path.join(ROOT, "kernelabstractions.jl_b200", "julia", "B200Backend", "src", "B200BackendModule.jl")


# ------------------------------------------------------------------------------------------------------------
# tiny parsers
# ------------------------------------------------------------------------------------------------------------
def strip_c_comments(src):
    return re.sub(r"/\*.*?\*/", "", src, flags=re.S)


def strip_jl_comments(src):
    out = []
    for line in src.splitlines():
        # a '#' outside a string literal starts a comment (the file has no '#' inside strings except "$..." interpolation-free ones)
        in_str, i, cut = False, 0, len(line)
        while i < len(line):
            c = line[i]
            if c == '"' and (i == 0 or line[i - 1] != "\\"):
                in_str = not in_str
            elif c == "#" and not in_str:
                cut = i
                break
            i += 1
        out.append(line[:cut])
    return "\n".join(out)


def split_top(s, sep=","):
    """split at top-level separators (outside (), [], {} and string literals)"""
    parts, depth, cur, in_str = [], 0, [], False
    for i, c in enumerate(s):
        if c == '"' and (i == 0 or s[i - 1] != "\\"):
            in_str = not in_str
        if not in_str:
            if c in "([{":
                depth += 1
            elif c in ")]}":
                depth -= 1
            elif depth == 0 and (c == sep or (sep == " " and c.isspace())):
                if "".join(cur).strip():
                    parts.append("".join(cur).strip())
                cur = []
                continue
        cur.append(c)
    if "".join(cur).strip():
        parts.append("".join(cur).strip())
    return parts


def balanced(s, start):
    """s[start] == '(' -> index one past the matching ')'"""
    assert s[start] == "("
    depth, in_str = 0, False
    for i in range(start, len(s)):
        c = s[i]
        if c == '"' and s[i - 1] != "\\":
            in_str = not in_str
        if in_str:
            continue
        if c == "(":
            depth += 1
        elif c == ")":
            depth -= 1
            if depth == 0:
                return i + 1
    raise AssertionError("unbalanced parenthesis")


def header_prototypes():
    src = strip_c_comments(open(HEADER).read())
    protos = {}
    for m in re.finditer(r"(?:^|\n)\s*((?:const\s+)?[A-Za-z_0-9]+\s*\*?)\s*(b200_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if name in ("b200_kernel_handler",):
            continue
        argl = [] if args in ("void", "") else [a.strip() for a in split_top(args)]
        protos[name] = (ret.replace(" ", ""), argl)
    return protos


def c_class(ctype):
    """coarse class of a C parameter declaration"""
    t = re.sub(r"/\*.*", "", ctype).strip()
    if "*" in t or "b200_kernel_handler" in t:
        return "ptr"
    t = re.sub(r"\bconst\b", "", t).strip()
    base = t.split()[0]
    return {"int32_t": "i32", "b200_status": "i32", "int64_t": "i64", "size_t": "usize", "uint64_t": "u64", "double": "f64",
            "float": "f32", "int": "i32"}[base]


def jl_class(jtype):
    t = jtype.strip()
    if t.startswith(("Ptr{", "Ref{")) or t in ("Cstring", "Ptr"):
        return "ptr"
    return {"Int32": "i32", "Cint": "i32", "Int64": "i64", "Csize_t": "usize", "UInt64": "u64", "Float64": "f64", "Cdouble": "f64",
            "Float32": "f32"}[t]


def julia_calls():
    """[(symbol, ret, [argtypes], nargs, line)] for every ccall and @b200 in the Julia module"""
    src = strip_jl_comments(open(JL).read())
    calls = []
    for m in re.finditer(r"ccall\(", src):
        if "macro b200" in src[max(0, m.start() - 200):m.start()] and "QuoteNode" in src[m.start():m.start() + 80]:
            continue   # the macro body itself
        end = balanced(src, m.end() - 1)
        parts = split_top(src[m.end():end - 1])
        sym = re.match(r"\(\s*:(b200_[a-z0-9_]+)\s*,\s*libb200ka\s*\)", parts[0])
        assert sym, f"ccall target not (:b200_x, libb200ka): {parts[0]}"
        types = split_top(parts[2].strip()[1:-1])
        calls.append((sym.group(1), parts[1].strip(), types, len(parts) - 3, src.count("\n", 0, m.start()) + 1))
    for m in re.finditer(r"@b200\s+(b200_[a-z0-9_]+)\s+\(", src):
        tstart = m.end() - 1
        tend = balanced(src, tstart)
        types = split_top(src[tstart + 1:tend - 1])
        rest = src[tend:].split("\n", 1)[0]
        # the macro call ends at the line end, or at a closing `;` / `)` of an enclosing one-liner
        toks = []
        for tok in split_top(rest, sep=" "):
            if tok.startswith((";", ")")) or tok in ("end",):
                break
            toks.append(tok.rstrip(";"))
            if tok.endswith(";"):
                break
        calls.append((m.group(1), "Int32", types, len(toks), src.count("\n", 0, m.start()) + 1))
    return calls


# ------------------------------------------------------------------------------------------------------------
# 1. ccall symbols / arity / types
# ------------------------------------------------------------------------------------------------------------
def test_every_ccall_matches_the_header():
    protos = header_prototypes()
    assert len(protos) >= 80, sorted(protos)
    calls = julia_calls()
    assert len(calls) >= 50
    for sym, ret, types, nargs, line in calls:
        assert sym in protos, f"B200BackendModule.jl:{line}: {sym} is not declared in include/b200ka.h"
        cret, cargs = protos[sym]
        assert len(types) == len(cargs), f"B200BackendModule.jl:{line}: {sym} takes {len(cargs)} arguments, ccall declares {len(types)}"
        assert nargs == len(types), f"B200BackendModule.jl:{line}: {sym}: {len(types)} argument types but {nargs} arguments passed"
        want_ret = "ptr" if "*" in cret else c_class(cret)
        got_ret = "ptr" if ret in ("Cstring",) or ret.startswith("Ptr") else jl_class(ret)
        assert want_ret == got_ret, f"B200BackendModule.jl:{line}: {sym} returns {cret}, ccall says {ret}"
        for k, (jt, ct) in enumerate(zip(types, cargs)):
            assert jl_class(jt) == c_class(ct), f"B200BackendModule.jl:{line}: {sym} argument {k + 1}: Julia {jt} vs C `{ct}`"


def test_binding_covers_the_reference_facing_entry_points():
    used = {c[0] for c in julia_calls()}
    for sym in ["b200_malloc", "b200_free", "b200_memcpy_h2d_async", "b200_memcpy_d2h_async", "b200_copy", "b200_fill", "b200_launch",
                "b200_stream_query", "b200_bind", "b200_stream_create", "b200_device_count", "b200_host_register",
                "b200_kernel_max_work_group_size", "b200_max_threads_per_block", "b200_sm_count", "b200_last_error",
                "b200_reduce", "b200_count_mismatch", "b200_norms", "b200_matmul_f32"]:
        assert sym in used, f"{sym} is never called from the Julia layer"


# ------------------------------------------------------------------------------------------------------------
# 2. records and constants
# ------------------------------------------------------------------------------------------------------------
def c_struct_fields(name):
    src = strip_c_comments(open(HEADER).read())
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), src, flags=re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        m = re.match(r"(.+?)\s*(\*?)\s*([a-z_]+)(\[[A-Z0-9_]+\])?$", decl)
        ctype, star, fname, arr = m.group(1).strip(), m.group(2), m.group(3), m.group(4)
        cls = "ptr" if star else c_class(ctype)
        fields.append((fname, cls + ("x4" if arr else "")))
    return fields


def jl_struct_fields(name):
    src = strip_jl_comments(open(JL).read())
    body = re.search(r"\nstruct %s\b(.*?)\nend" % name, src, flags=re.S).group(1)
    fields = []
    for line in body.splitlines():
        line = line.strip()
        if "::" not in line:
            continue
        fname, jt = [x.strip() for x in line.split("::")]
        if jt.startswith("NTuple{4"):
            cls = jl_class(jt[len("NTuple{4,"):-1].strip()) + "x4"
        else:
            cls = jl_class(jt)
        fields.append((fname, cls))
    return fields


def test_records_match_the_header_field_for_field():
    assert jl_struct_fields("LaunchDesc") == c_struct_fields("b200_launch_desc")
    assert jl_struct_fields("CArg") == c_struct_fields("b200_arg")


def test_status_and_enum_constants_match():
    hdr = strip_c_comments(open(HEADER).read())
    jl = strip_jl_comments(open(JL).read())
    consts = dict(re.findall(r"const (B200_[A-Z_]+) = Int32\((-?\d+)\)", jl))
    assert len(consts) >= 6
    for k, v in consts.items():
        m = re.search(r"\b%s = (-?\d+)" % k, hdr)
        assert m and int(m.group(1)) == int(v), f"{k}: Julia {v} vs header {m and m.group(1)}"
    assert re.search(r"const B200_MAX_NDIM = 4", jl) and re.search(r"#define B200_MAX_NDIM 4", hdr)
    # eltype codes
    codes = dict(re.findall(r"B200_T_([A-Z0-9]+) = (\d+)", hdr))
    jmap = {"Int8": "I8", "UInt8": "U8", "Int16": "I16", "UInt16": "U16", "Int32": "I32", "UInt32": "U32", "Int64": "I64",
            "UInt64": "U64", "Float16": "F16", "Float32": "F32", "Float64": "F64", "Bool": "BOOL"}
    for jt, v in re.findall(r"eltype_code\(::Type\{(\w+)\}\) = Int32\((\d+)\)", jl):
        assert codes[jmap[jt]] == v, f"eltype_code({jt})"
    # launch kinds: KA launch = 0, KI launch = 1
    assert re.search(r"desc = LaunchDesc\(0, N,", jl) and re.search(r"desc = LaunchDesc\(1, nd,", jl)
    assert re.search(r"B200_LAUNCH_KA = 0, B200_LAUNCH_KI = 1", hdr)


# ------------------------------------------------------------------------------------------------------------
# 3. Appendix C method checklist (template: reference src/pocl/backend.jl)
# ------------------------------------------------------------------------------------------------------------
CHECKLIST = [
    # (what, regex over the comment-stripped module, reference line)
    ("struct B200Backend <: KA.GPU", r"struct B200Backend <: KA\.GPU end", "src/pocl/backend.jl:19-20"),
    ("B200Array <: AbstractArray", r"mutable struct B200Array\{T, N\} <: AbstractArray\{T, N\}", "Appendix C"),
    ("finalizer on allocation", r"finalizer\(KA\.unsafe_free!, A\)", "nanoOpenCL.jl:962-966"),
    ("KA.allocate", r"KA\.allocate\(::B200Backend, ::Type\{T\}, dims::Tuple; unified::Bool = false\) where \{T\}", ":25"),
    ("KA.zeros", r"function KA\.zeros\(backend::B200Backend, ::Type\{T\}, dims::Tuple; kwargs\.\.\.\) where \{T\}", ":27-32"),
    ("KA.ones", r"function KA\.ones\(backend::B200Backend, ::Type\{T\}, dims::Tuple; kwargs\.\.\.\) where \{T\}", ":33-38"),
    ("KA.copyto! d<-d", r"function KA\.copyto!\(::B200Backend, A::B200Array\{T\}, B::B200Array\{T\}\) where \{T\}", ":40-54"),
    ("KA.copyto! d<-h", r"function KA\.copyto!\(::B200Backend, A::B200Array\{T\}, B::Array\{T\}\) where \{T\}", ":52"),
    ("KA.copyto! h<-d", r"function KA\.copyto!\(::B200Backend, A::Array\{T\}, B::B200Array\{T\}\) where \{T\}", ":52"),
    ("length check text", r'error\("Arrays must match in length"\)', ":42-44"),
    ("KA.functional", r"KA\.functional\(::B200Backend\)", ":56"),
    ("KA.pagelock!", r"KA\.pagelock!\(::B200Backend, x::Array\)", ":57"),
    ("KA.get_backend", r"KA\.get_backend\(::B200Array\) = B200Backend\(\)", ":59"),
    ("KA.synchronize", r"function KA\.synchronize\(::B200Backend\)", ":67"),
    ("KA.supports_float64", r"KA\.supports_float64\(::B200Backend\) = true", ":68"),
    ("KA.supports_atomics", r"KA\.supports_atomics\(::B200Backend\) = true", ":69"),
    ("KA.supports_unified", r"KA\.supports_unified\(::B200Backend\) = true", "KernelAbstractions.jl:623-640"),
    ("KA.ndevices", r"function KA\.ndevices\(::B200Backend\)", "KernelAbstractions.jl:670-691"),
    ("KA.device", r"KA\.device\(::B200Backend\)", "KernelAbstractions.jl:670-691"),
    ("KA.device!", r"function KA\.device!\(backend::B200Backend, id::Int\)", "KernelAbstractions.jl:686-691"),
    ("device! ArgumentError", r'throw\(ArgumentError\("Device id \$id out of bounds\."\)\)', "test/devices.jl:10-11"),
    ("KA.priority!", r"function KA\.priority!\(::B200Backend, prio::Symbol\)", "KernelAbstractions.jl:658-663"),
    ("KA.unsafe_free!", r"function KA\.unsafe_free!\(A::B200Array\)", "KernelAbstractions.jl:193-195"),
    ("KA.mkcontext (2 forms)", r"function KA\.mkcontext\(kernel::KA\.Kernel\{B200Backend\}, _ndrange, iterspace\)", ":74-76"),
    ("KA.mkcontext dynamic", r"function KA\.mkcontext\(kernel::KA\.Kernel\{B200Backend\}, I, _ndrange, iterspace, ::Dynamic\) where \{Dynamic\}", ":77-82"),
    ("KA.launch_config", r"function KA\.launch_config\(kernel::KA\.Kernel\{B200Backend\}, ndrange, workgroupsize\)", ":84-106"),
    ("KA.Kernel call", r"function \(obj::KA\.Kernel\{B200Backend\}\)\(args\.\.\.; ndrange = nothing, workgroupsize = nothing\)", ":117-147"),
    ("groups == 0 returns", r"prod\(blocks\) == 0 && return nothing", ":136-138"),
    ("KA.argconvert", r"KA\.argconvert\(::KA\.Kernel\{B200Backend\}, arg\) = carg\(arg\)", ":242"),
    ("KI.argconvert(f::Function) = f", r"KI\.argconvert\(::B200Backend, f::F\) where \{F <: Function\} = f\b", ":149"),
    ("KI.argconvert generic", r"KI\.argconvert\(::B200Backend, arg\) = carg\(arg\)", ":149"),
    ("KI.kernel_function", r"function KI\.kernel_function\(::B200Backend, f::F, tt::TT = Tuple\{\}; name = nothing, kwargs\.\.\.\) where \{F, TT\}", ":151-154"),
    ("KI.Kernel call", r"function \(obj::KI\.Kernel\{B200Backend\}\)\(args\.\.\.; numworkgroups = 1, workgroupsize = 1\)", ":156-168"),
    ("KI.check_launch_args first", r"KI\.check_launch_args\(numworkgroups, workgroupsize\)", "src/intrinsics.jl:182-188"),
    ("KI.kernel_max_work_group_size", r"function KI\.kernel_max_work_group_size\(kernel::KI\.Kernel\{<:B200Backend\}; max_work_items::Int = typemax\(Int\)\)::Int", ":170-173"),
    ("KI.max_work_group_size", r"function KI\.max_work_group_size\(::B200Backend\)::Int", ":174-176"),
    ("KI.multiprocessor_count", r"function KI\.multiprocessor_count\(::B200Backend\)::Int", ":177-179"),
    ("Adapt host->device", r"Adapt\.adapt_storage\(::B200Backend, a::Array\) = B200Array\(a\)", "KernelAbstractions.jl:559-561"),
    ("Adapt identity", r"Adapt\.adapt_storage\(::B200Backend, a::B200Array\) = a", "test/test.jl:107-113"),
    ("Adapt device->host", r"Adapt\.adapt_storage\(::KA\.CPU, a::B200Array\) = Array\(a\)", "test/test.jl:107-113"),
    ("Base.size", r"Base\.size\(A::B200Array\)", "Appendix C"),
    ("Base.similar", r"Base\.similar\(A::B200Array\{T\}", "Appendix C"),
    ("Base.Array", r"function Base\.Array\(A::B200Array\{T, N\}\) where \{T, N\}", "Appendix C"),
    ("Base.copyto! h->d", r"function Base\.copyto!\(dst::B200Array\{T\}, src::Array\{T\}\) where \{T\}", "Appendix C"),
    ("Base.copyto! d->h", r"function Base\.copyto!\(dst::Array\{T\}, src::B200Array\{T\}\) where \{T\}", "Appendix C"),
    ("Base.fill!", r"function Base\.fill!\(A::B200Array\{T\}, x\) where \{T\}", "Appendix C"),
    ("scalar getindex guarded", r"A\.unified \|\| error\(", "test/test.jl:82-87"),
    ("B200Array{T}(undef, dims...)", r"B200Array\{T\}\(::UndefInitializer, dims::Integer\.\.\.\) where \{T\}", "Appendix C"),
    ("mul!", r"function LinearAlgebra\.mul!\(C::B200Array\{Float32, 2\}", "examples/performant_matmul.jl:92"),
    ("isapprox", r"function Base\.isapprox\(A::B200Array\{T\}, B::B200Array\{T\};", "examples/matmul.jl:36"),
    ("==", r"function Base\.:\(==\)\(A::B200Array\{T\}, B::B200Array\{T\}\) where \{T\}", "examples/memcopy.jl:23"),
    ("maximum", r"function Base\.maximum\(A::B200Array\{T\}\) where \{T <: Reducible\}", "examples/histogram.jl:10"),
]


@pytest.mark.parametrize("what,pattern,where", CHECKLIST, ids=[c[0] for c in CHECKLIST])
def test_appendix_c_method_exists(what, pattern, where):
    src = strip_jl_comments(open(JL).read())
    assert re.search(pattern, src), f"{what}: no definition matching /{pattern}/ (reference {where})"


# ------------------------------------------------------------------------------------------------------------
# 4. every name the reference's callers use
# ------------------------------------------------------------------------------------------------------------
# names that need no backend method: KA defines them generically on top of the methods above, or they are device-side
INHERITED = {
    "KA.@context": "macro, expands to the hidden __ctx__ argument (src/KernelAbstractions.jl:85-96); device side lives in ka_device.cuh",
    "KA.Extras": "submodule with @unroll (src/extras/loopinfo.jl) -- loop metadata only, reused unchanged",
    "KA.Kernel": "the struct itself (src/KernelAbstractions.jl:735-746); the backend defines its call method",
    "KA.KernelIntrinsics": "submodule; the backend implements its host methods, device ids live in the .so",
    "KA.NDIteration": "submodule reused unchanged (src/nditeration.jl)",
    "KA.adapt": "Adapt.adapt -> Adapt.adapt_storage methods defined here",
    "KA.backend": "accessor of KA.Kernel (src/KernelAbstractions.jl:749)",
    "KA.partition": "reused unchanged (src/KernelAbstractions.jl:751-806); the launch calls it",
    "KI.@kernel": "macro (src/intrinsics.jl:300-373): expands to argconvert / kernel_function / the KI.Kernel call defined here",
    "KI.barrier": "device side: __syncthreads() in the precompiled kernels (csrc/ka_device.cuh)",
    "KI.get_global_id": "device side (csrc/ka_device.cuh)", "KI.get_global_size": "device side (csrc/ka_device.cuh)",
    "KI.get_group_id": "device side (csrc/ka_device.cuh)", "KI.get_local_id": "device side (csrc/ka_device.cuh)",
    "KI.get_local_size": "device side (csrc/ka_device.cuh)", "KI.get_num_groups": "device side (csrc/ka_device.cuh)",
    "KI.localmemory": "device side: static __shared__ arrays in the precompiled kernels",
}


def test_every_reference_call_has_a_method():
    calls = json.load(open(os.path.join(ROOT, "tests", "golden", "api_calls.json")))["calls"]
    assert len(calls) >= 30
    src = strip_jl_comments(open(JL).read())
    for name, where in calls.items():
        if name in INHERITED:
            continue
        mod, fn = name.split(".", 1)
        pat = r"(?:function\s+)?%s\.%s\(" % (mod, re.escape(fn))
        assert re.search(pat, src), f"{name} (first used at reference {where}) has no method in B200BackendModule.jl"
    for name in INHERITED:
        assert name in calls, f"stale INHERITED entry {name}"


# ------------------------------------------------------------------------------------------------------------
# 5. the KI.@kernel expansion, statement by statement (reference src/intrinsics.jl:352-360)
# ------------------------------------------------------------------------------------------------------------
def method_bodies(src, head_regex):
    """source text of every method whose head matches (one-liners: to the line end; function blocks: to the matching end)"""
    out = []
    for m in re.finditer(head_regex, src):
        line_start = src.rfind("\n", 0, m.start()) + 1
        if src[line_start:m.start()].strip().startswith("function") or src[m.start():m.start() + 8] == "function":
            depth, i = 0, line_start
            for ln in src[line_start:].splitlines(keepends=True):
                st = ln.strip()
                if re.match(r"(function|if|for|while|let|begin|try|do|quote|struct|mutable struct|macro)\b", st) or re.search(r"\bdo\b[^\n]*$", st) or \
                        re.search(r"=\s*(if|begin|let|try)\b", st) or re.search(r"\bbegin$", st):
                    depth += 1
                if st == "end" or st.startswith("end ") or st.startswith("end)") or st.endswith(" end"):
                    depth -= 1
                i += len(ln)
                if depth <= 0:
                    break
            out.append(src[line_start:i])
        else:
            out.append(src[line_start:src.find("\n", m.end())])
    return out


def test_ki_kernel_macro_expansion_resolves():
    src = strip_jl_comments(open(JL).read())
    # $kernel_f = argconvert(backend, f_var): the most specific method for a Function must return it unchanged
    fn_methods = method_bodies(src, r"KI\.argconvert\(::B200Backend, \w+::\w+\) where \{\w+ <: Function\}")
    assert len(fn_methods) == 1, "KI.argconvert needs exactly one method for kernel functions"
    assert re.search(r"=\s*f\s*$", fn_methods[0].strip()), fn_methods[0]
    assert "error(" not in fn_methods[0]
    # $kernel_args = map(x -> argconvert(backend, x), args): arrays, integers, floats, Val and types all have a carg method
    for sig in [r"carg\(A::B200Array\{T, N\}\)", r"carg\(x::Integer\)", r"carg\(x::AbstractFloat\)", r"carg\(::Val\{x\}\)", r"carg\(::Type\{T\}\)"]:
        assert re.search(sig, src), sig
    # carg(::Type) must not take sizeof of an abstract type
    ty = [b for b in method_bodies(src, r"carg\(::Type\{T\}\) where \{T\}")][0]
    assert "isconcretetype" in ty or "isbitstype" in ty, "carg(::Type{T}) calls sizeof(T) unguarded"
    # $kernel = kernel_function(backend, kernel_f, kernel_tt; name...): receives the FUNCTION, names it with nameof
    kf = method_bodies(src, r"function KI\.kernel_function\(::B200Backend")[0]
    assert "kernel_name(f)" in kf and re.search(r"kernel_name\(f::Function\) = String\(nameof\(f\)\)", src)
    assert "KI.Kernel{B200Backend, B200Function}(B200Backend(), B200Function(kname))" in kf
    # $kernel(args...; numworkgroups, workgroupsize): converts the ORIGINAL args itself
    call = method_bodies(src, r"function \(obj::KI\.Kernel\{B200Backend\}\)\(args\.\.\.")[0]
    assert "launch(obj.kern.name, desc, args)" in call
    launch = method_bodies(src, r"function launch\(name::String, desc::LaunchDesc, args\)")[0]
    assert "CArg[carg(a) for a in args]" in launch and "GC.@preserve args cargs" in launch
    # names the two KI examples launch exist in the catalogue of the library under exactly nameof(f)
    import b200ka as KA
    names = set(KA.registered_kernels())
    for n in ["histogram_kernel!", "coalesced_matmul_kernel!", "test_intrinsics_kernel"]:
        assert n in names


# ------------------------------------------------------------------------------------------------------------
# 6. task-local runtime state
# ------------------------------------------------------------------------------------------------------------
def test_runtime_state_is_task_local():
    src = strip_jl_comments(open(JL).read())
    assert "task_local_storage()" in src and ":B200State" in src
    macro = re.search(r"macro b200\(name, argtypes, args\.\.\.\)(.*?)\nend", src, flags=re.S).group(1)
    assert macro.index("bind!()") < macro.index("ccall("), "@b200 must bind the task's device and stream before the call"
    sync = method_bodies(src, r"function KA\.synchronize\(::B200Backend\)")[0]
    assert sync.index("bind!()") < sync.index("b200_stream_query") and "yield()" in sync
    # device! / priority! / device touch only the task's state -- the OS thread's is set by bind!
    for head in [r"function KA\.device!\(backend::B200Backend, id::Int\)", r"function KA\.priority!\(::B200Backend, prio::Symbol\)"]:
        body = method_bodies(src, head)[0]
        assert "task_state()" in body and "b200_set_device" not in body and "b200_set_priority" not in body
    # async host copies keep both arrays alive for the enqueue
    for head in [r"function KA\.copyto!\(::B200Backend, A::B200Array\{T\}, B::Array\{T\}\)", r"function KA\.copyto!\(::B200Backend, A::Array\{T\}, B::B200Array\{T\}\)"]:
        assert "GC.@preserve A B" in method_bodies(src, head)[0]


# ------------------------------------------------------------------------------------------------------------
# 7. shallow syntax check: block openers and `end`s balance, brackets balance (nothing here can parse Julia properly, but an
#    unbalanced edit is the most likely slip in a file that is never executed)
# ------------------------------------------------------------------------------------------------------------
def test_julia_blocks_and_brackets_balance():
    src = strip_jl_comments(open(JL).read())
    src = re.sub(r'"(?:\\.|[^"\\])*"', '""', src)                      # string literals out of the way
    depth_paren = depth_brack = depth_brace = 0
    for ch in src:
        depth_paren += ch == "("
        depth_paren -= ch == ")"
        depth_brack += ch == "["
        depth_brack -= ch == "]"
        depth_brace += ch == "{"
        depth_brace -= ch == "}"
        assert depth_paren >= 0 and depth_brack >= 0 and depth_brace >= 0
    assert (depth_paren, depth_brack, depth_brace) == (0, 0, 0)
    no_index = src
    while True:                                                          # `a[end]` is not a block end, a comprehension's `for` no opener
        nxt = re.sub(r"\[[^\[\]]*\]", "<>", no_index)
        if nxt == no_index:
            break
        no_index = nxt
    openers = len(re.findall(r"(?<![\w.:@])(?:function|if|for|while|let|begin|try|quote|struct|macro|module|do)\b(?!\s*=)", no_index))
    openers -= len(re.findall(r"\bmutable\s+struct\b", no_index)) * 0  # `mutable struct` counts once (matched at `struct`)
    ends = len(re.findall(r"(?<![\w.:])end\b", no_index))
    assert openers == ends, f"{openers} block openers vs {ends} `end`s"
