# B200Backend.jl -- the Julia host layer: `struct B200Backend <: KA.GPU` over libb200ka.so (`ccall`).
#
# This file is the twin of the in-tree backend glue, reference src/pocl/backend.jl (all 244 lines), with the
# OpenCL runtime (src/pocl/nanoOpenCL.jl) replaced by the C ABI declared in include/b200ka.h.  KA's front end
# (`@kernel`, `@index`, `@localmem`, `@private`, `@synchronize`, `@Const`), `KA.partition` and NDIteration are
# reused UNCHANGED; only the methods a backend must provide are defined here.
#
# NOTE: there is no Julia toolchain in the build environment of this repository, so this file is not executed there.
# It is checked mechanically instead (tests/test_julia_binding.py): every ccall against include/b200ka.h (symbol, arity,
# argument and return types, record layouts, constants), every method of the backend interface against the template
# src/pocl/backend.jl, every host-side name the reference's tests and examples call, and the KI.@kernel expansion
# (src/intrinsics.jl:352-360) statement by statement.  Behaviour is pinned through the Python mirror
# (kernelabstractions.jl_b200/b200ka), which implements the same methods over the same C ABI and runs the ported
# testsuite on a B200.
#
# There is no JIT.  A `@kernel` launched on B200Backend is dispatched BY NAME (`nameof(obj.f)`, i.e.
# `:gpu_<kernel name>`, reference src/macros.jl:32) to a precompiled sm_100a kernel inside the library; an
# unknown name raises an ErrorException.  The catalogue is `B200Backend.registered_kernels()`.
module B200BackendModule

import KernelAbstractions as KA
import KernelAbstractions.KernelIntrinsics as KI
import Adapt
import LinearAlgebra

export B200Backend, B200Array

# ---------------------------------------------------------------------------------------------------------
# library handle + checked calls (twin of `@checked`, reference src/pocl/nanoOpenCL.jl:22-52,392-408)
# ---------------------------------------------------------------------------------------------------------
const libb200ka = get(ENV, "B200KA_LIB", joinpath(@__DIR__, "..", "..", "..", "libb200ka.so"))

const B200_OK = Int32(0)
const B200_NOT_READY = Int32(1)
const B200_ERR_UNKNOWN_KERNEL = Int32(-3)
const B200_ERR_LAUNCH_DIMS = Int32(-5)
const B200_ERR_LENGTH_MISMATCH = Int32(-6)
const B200_ERR_ALIAS = Int32(-7)
const B200_ERR_BAD_DEVICE = Int32(-12)

struct B200Error <: Exception
    code::Int32
    msg::String
end
Base.showerror(io::IO, e::B200Error) = print(io, "B200Error($(e.code)): ", e.msg)

last_error() = unsafe_string(ccall((:b200_last_error, libb200ka), Cstring, ()))

function check(status::Int32)
    status == B200_OK && return nothing
    msg = last_error()
    # same exception *types* as the reference (tests assert them: test/intrinsics.jl:79-80, test/devices.jl:10-11)
    (status == B200_ERR_LAUNCH_DIMS || status == B200_ERR_BAD_DEVICE) && throw(ArgumentError(msg))
    (status == B200_ERR_UNKNOWN_KERNEL || status == B200_ERR_LENGTH_MISMATCH || status == B200_ERR_ALIAS) && error(msg)
    throw(B200Error(status, msg))
end

# ---------------------------------------------------------------------------------------------------------
# Task-local runtime state (twin of the task-local platform/device/context/queue, reference src/pocl/pocl.jl:12-41).
# The library keeps "current device" and "current stream" per OS thread, but a Julia task migrates between OS threads at
# every yield point -- and KA.synchronize itself yields.  So device, priority and the stream handles live in
# `task_local_storage()`, and every checked call first binds them to whatever OS thread the task is running on right now
# (`b200_bind`: cudaSetDevice + adopt the stream).  There is no yield point between the bind and the ccall that follows it.
# ---------------------------------------------------------------------------------------------------------
mutable struct TaskState
    device::Int32                                        # 0-based
    priority::Int32                                      # -1 :low, 0 :normal, 1 :high
    streams::Dict{Tuple{Int32, Int32}, Ptr{Cvoid}}       # (device, priority) -> cudaStream_t owned by this task
end

function destroy_streams(st::TaskState)
    for s in values(st.streams)
        ccall((:b200_stream_destroy, libb200ka), Int32, (Ptr{Cvoid},), s)     # waits for nothing: cudaStreamDestroy defers
    end
    empty!(st.streams)
    return nothing
end

function task_state()
    return get!(task_local_storage(), :B200State) do
        finalizer(destroy_streams, TaskState(Int32(0), Int32(0), Dict{Tuple{Int32, Int32}, Ptr{Cvoid}}()))
    end::TaskState
end

function bind!()
    st = task_state()
    s = get(st.streams, (st.device, st.priority), C_NULL)
    if s == C_NULL
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:b200_bind, libb200ka), Int32, (Int32, Ptr{Cvoid}), st.device, C_NULL))
        check(ccall((:b200_stream_create, libb200ka), Int32, (Ptr{Ptr{Cvoid}}, Int32), r, st.priority))
        s = st.streams[(st.device, st.priority)] = r[]
    end
    check(ccall((:b200_bind, libb200ka), Int32, (Int32, Ptr{Cvoid}), st.device, s))
    return nothing
end

macro b200(name, argtypes, args...)
    return quote
        bind!()
        check(ccall(($(QuoteNode(name)), libb200ka), Int32, $(esc(argtypes)), $(map(esc, args)...)))
    end
end

# ---------------------------------------------------------------------------------------------------------
# Back-end definition (reference src/pocl/backend.jl:19-20)
# ---------------------------------------------------------------------------------------------------------
struct B200Backend <: KA.GPU end

# ---------------------------------------------------------------------------------------------------------
# Device array: pointer + dims, column-major, owner of its allocation (SURVEY.md Appendix C)
# ---------------------------------------------------------------------------------------------------------
mutable struct B200Array{T, N} <: AbstractArray{T, N}
    ptr::Ptr{T}
    dims::Dims{N}
    unified::Bool
    parent::Any           # keeps the owner alive for views
    function B200Array{T, N}(::UndefInitializer, dims::Dims{N}; unified::Bool = false) where {T, N}
        p = Ref{Ptr{Cvoid}}(C_NULL)
        @b200 b200_malloc (Ptr{Ptr{Cvoid}}, Csize_t, Int32) p (prod(dims) * sizeof(T)) Int32(unified)
        A = new{T, N}(convert(Ptr{T}, p[]), dims, unified, nothing)
        finalizer(KA.unsafe_free!, A)
        return A
    end
    B200Array{T, N}(ptr::Ptr{T}, dims::Dims{N}, unified::Bool, parent) where {T, N} = new{T, N}(ptr, dims, unified, parent)
end
B200Array{T}(::UndefInitializer, dims::Integer...) where {T} = B200Array{T, length(dims)}(undef, Dims(dims))
B200Array{T}(::UndefInitializer, dims::Dims{N}) where {T, N} = B200Array{T, N}(undef, dims)
function B200Array(h::Array{T, N}) where {T, N}
    A = B200Array{T, N}(undef, size(h))
    copyto!(A, h)
    return A
end

Base.size(A::B200Array) = A.dims
Base.length(A::B200Array) = prod(A.dims)
Base.eltype(::B200Array{T}) where {T} = T
Base.elsize(::Type{<:B200Array{T}}) where {T} = sizeof(T)
Base.pointer(A::B200Array) = A.ptr
Base.similar(A::B200Array{T}, ::Type{S} = T, dims::Dims = size(A)) where {T, S} = B200Array{S}(undef, dims)
Base.unsafe_convert(::Type{Ptr{T}}, A::B200Array{T}) where {T} = A.ptr
Base.dataids(A::B200Array) = (UInt(A.ptr),)

# scalar getindex throws unless the memory is unified (reference test/test.jl:82-87)
function Base.getindex(A::B200Array{T}, i::Int) where {T}
    A.unified || error("Scalar indexing of a B200Array is disallowed; copy it to the host with Array(A)")
    KA.synchronize(B200Backend())
    return unsafe_load(A.ptr, i)
end
Base.getindex(A::B200Array, I::Int...) = A[LinearIndices(A.dims)[I...]]

function Base.Array(A::B200Array{T, N}) where {T, N}
    h = Array{T, N}(undef, A.dims)
    copyto!(h, A)
    return h
end
function Base.copyto!(dst::B200Array{T}, src::Array{T}) where {T}   # synchronous host -> device
    length(dst) == length(src) || error("Arrays must match in length")
    GC.@preserve src begin
        @b200 b200_memcpy_h2d_async (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t) dst.ptr pointer(src) sizeof(src)
        @b200 b200_synchronize ()
    end
    return dst
end
function Base.copyto!(dst::Array{T}, src::B200Array{T}) where {T}   # synchronous device -> host
    length(dst) == length(src) || error("Arrays must match in length")
    GC.@preserve dst begin
        @b200 b200_memcpy_d2h_async (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t) pointer(dst) src.ptr sizeof(dst)
        @b200 b200_synchronize ()
    end
    return dst
end
function Base.fill!(A::B200Array{T}, x) where {T}                 # also `A .= x` via the broadcast fallback below
    v = Ref(convert(T, x))
    GC.@preserve v @b200 b200_fill (Ptr{Cvoid}, Csize_t, Int32, Ptr{Cvoid}) A.ptr length(A) Int32(sizeof(T)) Base.unsafe_convert(Ptr{Cvoid}, v)
    return A
end
Base.Broadcast.materialize!(A::B200Array, bc::Base.Broadcast.Broadcasted{<:Any, <:Any, typeof(identity), <:Tuple{Number}}) =
    fill!(A, bc.args[1])
# ---------------------------------------------------------------------------------------------------------
# Device-array conveniences (SURVEY.md 8(f)-4): what Base / LinearAlgebra give a CPU() user on plain Arrays, computed on the
# device (csrc/kernels/arrayops.cu) so the examples run without host round trips.  Mixed host/device operands fall back
# to a host copy.
# ---------------------------------------------------------------------------------------------------------
# (eltype_code(T): the B200_T_* code of an element type, defined with the argument marshalling below)
const Reducible = Union{Int32, UInt32, Int64, UInt64, Float32, Float64}
sumtype(::Type{T}) where {T <: Signed} = Int64
sumtype(::Type{T}) where {T <: Unsigned} = UInt64
sumtype(::Type{T}) where {T <: AbstractFloat} = Float64

function reduce_device(op::Integer, A::B200Array{T}, ::Type{R}) where {T <: Reducible, R}
    out = Ref{R}(zero(R))
    GC.@preserve out @b200 b200_reduce (Int32, Int32, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}) Int32(op) eltype_code(T) A.ptr length(A) Base.unsafe_convert(Ptr{Cvoid}, out)
    return out[]
end
Base.sum(A::B200Array{T}) where {T <: Reducible} = reduce_device(0, A, sumtype(T))
function Base.maximum(A::B200Array{T}) where {T <: Reducible}       # reference examples/histogram.jl:10,88-90
    isempty(A) && throw(ArgumentError("reducing over an empty collection is not allowed"))
    return reduce_device(1, A, T)
end
function Base.minimum(A::B200Array{T}) where {T <: Reducible}
    isempty(A) && throw(ArgumentError("reducing over an empty collection is not allowed"))
    return reduce_device(2, A, T)
end

function Base.:(==)(A::B200Array{T}, B::B200Array{T}) where {T}     # reference examples/memcopy.jl:23, test/copyto.jl:17-19
    size(A) == size(B) || return false
    isbitstype(T) || return Array(A) == Array(B)
    bad = Ref{Int64}(0)
    @b200 b200_count_mismatch (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Int32, Int32, Ptr{Int64}) A.ptr B.ptr length(A) eltype_code(T) Int32(sizeof(T)) bad
    return bad[] == 0
end
Base.:(==)(A::B200Array, B::AbstractArray) = Array(A) == (B isa B200Array ? Array(B) : collect(B))
Base.:(==)(A::AbstractArray, B::B200Array) = B == A

function Base.isapprox(A::B200Array{T}, B::B200Array{T}; atol::Real = 0, rtol::Real = atol > 0 ? 0 : sqrt(eps(T))) where {T <: Union{Float32, Float64}}
    size(A) == size(B) || throw(DimensionMismatch("isapprox"))
    out = zeros(Float64, 3)                                         # sum (x-y)^2, sum x^2, sum y^2   (examples/matmul.jl:36)
    GC.@preserve out @b200 b200_norms (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Int32, Ptr{Float64}) A.ptr B.ptr length(A) Int32(T === Float64) pointer(out)
    return sqrt(out[1]) <= max(atol, rtol * max(sqrt(out[2]), sqrt(out[3])))
end
Base.isapprox(A::B200Array, B::AbstractArray; kw...) = isapprox(Array(A), B isa B200Array ? Array(B) : collect(B); kw...)
Base.isapprox(A::AbstractArray, B::B200Array; kw...) = isapprox(B, A; kw...)

# LinearAlgebra.mul!(C, A, B) / A * B -> the FP32 SIMT GEMM in the reference kernel's summation order (mode 0)
function LinearAlgebra.mul!(C::B200Array{Float32, 2}, A::B200Array{Float32, 2}, B::B200Array{Float32, 2})
    (size(A, 2) == size(B, 1) && size(C) == (size(A, 1), size(B, 2))) || throw(DimensionMismatch("mul!"))
    M, K, N = size(A, 1), size(A, 2), size(B, 2)
    (M == 0 || N == 0) && return C
    K == 0 && return fill!(C, 0.0f0)
    @b200 b200_matmul_f32 (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int64, Int64, Int32) C.ptr A.ptr B.ptr M K N M K M Int32(0)
    return C
end
Base.:*(A::B200Array{Float32, 2}, B::B200Array{Float32, 2}) = LinearAlgebra.mul!(B200Array{Float32, 2}(undef, (size(A, 1), size(B, 2))), A, B)

# host-array forms (csrc/hostpipe.cu): plain Arrays in and out, slabs streamed through the GPU with h2d | kernel | d2h overlapped;
# stream ordered -- KA.synchronize(B200Backend()) completes them; page-lock the arrays (KA.pagelock!) for full PCIe speed
function transpose_host!(a::Matrix{T}, b::Matrix{T}) where {T}
    size(a) == reverse(size(b)) || (println("Matrix size mismatch!"); return nothing)      # examples/naive_transpose.jl:12-15
    GC.@preserve a b @b200 b200_transpose_host (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int32) pointer(a) pointer(b) size(a, 1) size(a, 2) size(a, 1) size(b, 1) Int32(sizeof(T))
    return nothing
end
function histogram_host!(bins::Vector{Int32}, input::Vector{Int32})
    GC.@preserve bins input @b200 b200_histogram_i32_host (Ptr{Int32}, Int64, Ptr{Int32}, Int64) pointer(bins) length(bins) pointer(input) length(input)
    return nothing
end
function matmul_host!(C::Matrix{Float32}, A::Matrix{Float32}, B::Matrix{Float32}; mode::Integer = 0)
    (size(A, 2) == size(B, 1) && size(C) == (size(A, 1), size(B, 2))) || throw(DimensionMismatch("matmul_host!"))
    GC.@preserve C A B @b200 b200_matmul_f32_host (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int64, Int64, Int32) pointer(C) pointer(A) pointer(B) size(A, 1) size(A, 2) size(B, 2) size(A, 1) size(B, 1) size(C, 1) Int32(mode)
    return nothing
end

# opt-in tensor-core products (no reference twin): C = A * B with TF32 / BF16 tcgen05 arithmetic, FP32 storage
function matmul_tf32!(C::B200Array{Float32, 2}, A::B200Array{Float32, 2}, B::B200Array{Float32, 2})
    M, K, N = size(A, 1), size(A, 2), size(B, 2)
    @b200 b200_matmul_tf32 (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int64, Int64) C.ptr A.ptr B.ptr M K N M K M
    return C
end
function matmul_tf32x3!(C::B200Array{Float32, 2}, A::B200Array{Float32, 2}, B::B200Array{Float32, 2};      # FP32-grade, rtol 1e-5
                        workspace::B200Array{Float32, 1} = B200Array{Float32, 1}(undef, (length(A) + length(B),)))
    M, K, N = size(A, 1), size(A, 2), size(B, 2)
    @b200 b200_matmul_f32_tf32x3 (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int64, Int64, Ptr{Cvoid}) C.ptr A.ptr B.ptr M K N M K M workspace.ptr
    return C
end
function matmul_bf16!(C::B200Array{Float32, 2}, A::B200Array{Float32, 2}, B::B200Array{Float32, 2};
                      workspace::B200Array{UInt16, 1} = B200Array{UInt16, 1}(undef, (length(A) + length(B),)))
    M, K, N = size(A, 1), size(A, 2), size(B, 2)
    @b200 b200_matmul_f32_via_bf16 (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int64, Int64, Ptr{Cvoid}) C.ptr A.ptr B.ptr M K N M K M workspace.ptr
    return C
end
# B converted once and kept resident (the sharded form: B is replicated at set-up, only the A row panel changes per call)
function prepare_bf16_b(B::B200Array{Float32, 2})
    out = B200Array{UInt16, 1}(undef, (length(B),))
    @b200 b200_convert_f32_bf16 (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int32) out.ptr B.ptr size(B, 1) size(B, 2) size(B, 1) Int32(0)
    return out
end
function matmul_bf16_resident_b!(C::B200Array{Float32, 2}, A::B200Array{Float32, 2}, Bbf::B200Array{UInt16, 1};
                                 workspace::B200Array{UInt16, 1} = B200Array{UInt16, 1}(undef, (length(A),)))
    M, K, N = size(A, 1), size(A, 2), size(C, 2)
    length(Bbf) == K * N || throw(DimensionMismatch("matmul_bf16_resident_b!"))
    @b200 b200_matmul_f32_bf16_resident_b (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int64, Ptr{Cvoid}) C.ptr A.ptr Bbf.ptr M K N M M workspace.ptr
    return C
end
Base.view(A::B200Array{T}, r::UnitRange{Int}) where {T} =
    B200Array{T, 1}(A.ptr + (first(r) - 1) * sizeof(T), (length(r),), A.unified, A)

# ---------------------------------------------------------------------------------------------------------
# Memory operations (reference src/pocl/backend.jl:25-57)
# ---------------------------------------------------------------------------------------------------------
KA.allocate(::B200Backend, ::Type{T}, dims::Tuple; unified::Bool = false) where {T} =
    B200Array{T, length(dims)}(undef, Dims(dims); unified)

function KA.zeros(backend::B200Backend, ::Type{T}, dims::Tuple; kwargs...) where {T}
    return fill!(KA.allocate(backend, T, dims; kwargs...), zero(T))      # init_kernel(arr, zero, T) -> b200_fill
end
function KA.ones(backend::B200Backend, ::Type{T}, dims::Tuple; kwargs...) where {T}
    return fill!(KA.allocate(backend, T, dims; kwargs...), one(T))
end

function KA.unsafe_free!(A::B200Array)
    if A.parent === nothing && A.ptr != C_NULL
        ccall((:b200_free, libb200ka), Int32, (Ptr{Cvoid},), A.ptr)
        A.ptr = C_NULL
    end
    return nothing
end

# KA.copyto! is asynchronous and stream ordered; the caller keeps host buffers alive until synchronize
# (contract: reference src/KernelAbstractions.jl:118-151).
function KA.copyto!(::B200Backend, A::B200Array{T}, B::B200Array{T}) where {T}
    length(A) == length(B) || error("Arrays must match in length")              # :42-44
    # the alias check (Base.mightalias, :45-47) is done in the library on the byte ranges -> B200_ERR_ALIAS
    @b200 b200_copy (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t) A.ptr B.ptr (length(A) * sizeof(T))
    return A
end
# host legs: the arrays are preserved for the duration of the ENQUEUE (reference src/intrinsics.jl:353,
# src/pocl/compiler/execution.jl:63); the caller keeps them alive until synchronize (src/KernelAbstractions.jl:127-142)
function KA.copyto!(::B200Backend, A::B200Array{T}, B::Array{T}) where {T}
    length(A) == length(B) || error("Arrays must match in length")
    GC.@preserve A B @b200 b200_memcpy_h2d_async (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t) A.ptr pointer(B) sizeof(B)
    return A
end
function KA.copyto!(::B200Backend, A::Array{T}, B::B200Array{T}) where {T}
    length(A) == length(B) || error("Arrays must match in length")
    GC.@preserve A B @b200 b200_memcpy_d2h_async (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t) pointer(A) B.ptr sizeof(A)
    return A
end

KA.functional(::B200Backend) = begin
    n = Ref{Int32}(0)
    ccall((:b200_device_count, libb200ka), Int32, (Ptr{Int32},), n) == B200_OK && n[] > 0
end
# The registration is dropped together with the array: a stale one would poison the address range for whatever the allocator
# puts there next (a later copy from a buffer lying PARTLY inside it fails with "invalid argument").
function KA.pagelock!(::B200Backend, x::Array)
    GC.@preserve x @b200 b200_host_register (Ptr{Cvoid}, Csize_t) pointer(x) sizeof(x)
    finalizer(a -> ccall((:b200_host_unregister, libb200ka), Int32, (Ptr{Cvoid},), pointer(a)), x)
    return nothing
end
KA.get_backend(::B200Array) = B200Backend()
KA.supports_float64(::B200Backend) = true
KA.supports_atomics(::B200Backend) = true
KA.supports_unified(::B200Backend) = true

# Cooperative synchronize: never block inside the driver, yield to the Julia scheduler
# (reference docs/src/implementations.md:3-11; examples/mpi.jl:30 relies on it).
function KA.synchronize(::B200Backend)
    while true
        bind!()                                   # the task may have moved to another OS thread during the last yield
        st = ccall((:b200_stream_query, libb200ka), Int32, ())
        st == B200_OK && break
        st == B200_NOT_READY || check(st)
        yield()
    end
    return nothing
end

function KA.priority!(::B200Backend, prio::Symbol)
    prio in (:high, :normal, :low) || error("priority must be one of :high, :normal, :low")   # KernelAbstractions.jl:658-663
    # task local like the reference's queue: later launches of THIS task go to its stream of that priority
    task_state().priority = Int32(prio === :high ? 1 : (prio === :low ? -1 : 0))
    return nothing
end

function KA.ndevices(::B200Backend)
    n = Ref{Int32}(0)
    check(ccall((:b200_device_count, libb200ka), Int32, (Ptr{Int32},), n))
    return Int(n[])
end
KA.device(::B200Backend) = Int(task_state().device) + 1
function KA.device!(backend::B200Backend, id::Int)
    (1 <= id <= KA.ndevices(backend)) || throw(ArgumentError("Device id $id out of bounds."))   # KernelAbstractions.jl:686-691
    task_state().device = Int32(id - 1)           # bound to the OS thread by the next call (bind!)
    return nothing
end

Adapt.adapt_storage(::B200Backend, a::Array) = B200Array(a)
Adapt.adapt_storage(::B200Backend, a::B200Array) = a
Adapt.adapt_storage(::KA.CPU, a::B200Array) = Array(a)

# ---------------------------------------------------------------------------------------------------------
# Argument marshalling: the C records of include/b200ka.h
# ---------------------------------------------------------------------------------------------------------
const B200_MAX_NDIM = 4

struct LaunchDesc            # b200_launch_desc
    kind::Int32
    ndim::Int32
    ndrange::NTuple{4, Int64}
    wgs::NTuple{4, Int64}
    blocks::NTuple{4, Int64}
    dynamic_check::Int32
    flags::Int32
end

struct CArg                  # b200_arg
    kind::Int32
    eltype::Int32
    elsize::Int32
    ndim::Int32
    dims::NTuple{4, Int64}
    ptr::Ptr{Cvoid}
    scalar::Int64
    fscalar::Float64
end

eltype_code(::Type{Int8}) = Int32(1);  eltype_code(::Type{UInt8}) = Int32(2)
eltype_code(::Type{Int16}) = Int32(3); eltype_code(::Type{UInt16}) = Int32(4)
eltype_code(::Type{Int32}) = Int32(5); eltype_code(::Type{UInt32}) = Int32(6)
eltype_code(::Type{Int64}) = Int32(7); eltype_code(::Type{UInt64}) = Int32(8)
eltype_code(::Type{Float16}) = Int32(9); eltype_code(::Type{Float32}) = Int32(11)
eltype_code(::Type{Float64}) = Int32(12); eltype_code(::Type{Bool}) = Int32(13)
eltype_code(::Type) = Int32(0)       # opaque isbits element (CartesianIndex{N}, KernelData, ...)

pad4(t::Tuple) = ntuple(i -> i <= length(t) ? Int64(t[i]) : Int64(1), 4)

# KA.argconvert / KI.argconvert (reference src/pocl/backend.jl:149,242)
carg(A::B200Array{T, N}) where {T, N} = CArg(0, eltype_code(T), sizeof(T), N, pad4(size(A)), convert(Ptr{Cvoid}, A.ptr), 0, 0.0)
carg(x::Integer) = CArg(1, eltype_code(Int64), 8, 0, pad4(()), C_NULL, Int64(x), Float64(x))
carg(x::AbstractFloat) = CArg(1, eltype_code(typeof(x)), sizeof(x), 0, pad4(()), C_NULL, isinteger(x) ? Int64(x) : 0, Float64(x))
carg(::Val{x}) where {x} = CArg(2, eltype_code(Int64), 8, 0, pad4(()), C_NULL, Int64(x), Float64(x))
# type arguments (convert_kernel!(A, B, ::Type{T})): only the eltype code travels; abstract / union types have no size
carg(::Type{T}) where {T} = CArg(2, eltype_code(T), (isconcretetype(T) && isbitstype(T)) ? sizeof(T) : 0, 0, pad4(()), C_NULL, Int64(eltype_code(T)), 0.0)
# function-valued ARGUMENTS of a kernel (init_kernel(arr, f, T), src/KernelAbstractions.jl:869-872): only zero / one have a twin
carg(f::Function) = f === zero ? carg(0) : f === one ? carg(1) : error("Don't know how to convert function $f for Kernel{B200Backend}")
carg(x) = error("Don't know how to convert arguments of type $(typeof(x)) for Kernel{B200Backend}")
KA.argconvert(::KA.Kernel{B200Backend}, arg) = carg(arg)
# KI.@kernel calls argconvert on the kernel FUNCTION first (reference src/intrinsics.jl:354; src/pocl/backend.jl:149 passes it
# through clconvert unchanged) and hands the result to kernel_function: the function must come back as itself -- its NAME
# selects the precompiled kernel.  Every other value becomes its C record (only the types are used by the macro, :355-356;
# the launch receives the ORIGINAL arguments, :358).
KI.argconvert(::B200Backend, f::F) where {F <: Function} = f
KI.argconvert(::B200Backend, arg) = carg(arg)

function launch(name::String, desc::LaunchDesc, args)
    cargs = CArg[carg(a) for a in args]
    GC.@preserve args cargs begin
        st = ccall((:b200_launch, libb200ka), Int32, (Cstring, Ref{LaunchDesc}, Ptr{CArg}, Int32),
                   name, Ref(desc), cargs, Int32(length(cargs)))
        check(st)
    end
    return nothing
end

# Symmetric device buffers + distributed transpose (b200_sym_*, b200_transpose_alltoall): one process per GPU (MPI.jl), the
# 64-byte IPC handles are exchanged with the launcher's transport, e.g. `all = MPI.Allgather(export_handle(s), comm)`.
mutable struct SymBuffer
    handle::Ptr{Cvoid}
    nbytes::Int
end
function SymBuffer(nbytes::Integer)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    @b200 b200_sym_alloc (Ptr{Ptr{Cvoid}}, Csize_t) h nbytes
    return SymBuffer(h[], nbytes)
end
function export_handle(s::SymBuffer)
    blob = Vector{UInt8}(undef, 64)
    GC.@preserve blob @b200 b200_sym_export (Ptr{Cvoid}, Ptr{Cvoid}) s.handle pointer(blob)
    return blob
end
connect!(s::SymBuffer, all_handles::Vector{UInt8}, world::Integer, rank::Integer) =
    (GC.@preserve all_handles @b200 b200_sym_connect (Ptr{Cvoid}, Int32, Int32, Ptr{Cvoid}) s.handle Int32(world) Int32(rank) pointer(all_handles); s)
function Base.unsafe_wrap(::Type{B200Array{T, N}}, s::SymBuffer, dims::Dims{N}) where {T, N}      # the own slab as a non-owning array
    p = Ref{Ptr{Cvoid}}(C_NULL)
    @b200 b200_sym_ptr (Ptr{Cvoid}, Int32, Ptr{Ptr{Cvoid}}) s.handle Int32(-1) p
    return B200Array{T, N}(Ptr{T}(p[]), dims, false, s)
end
barrier(s::SymBuffer) = (@b200 b200_sym_barrier (Ptr{Cvoid},) s.handle; nothing)
Base.close(s::SymBuffer) = (@b200 b200_sym_free (Ptr{Cvoid},) s.handle; s.handle = C_NULL; nothing)
# a = transpose(b) for a global N0 x N1 matrix sharded by columns: b_local is N0 x N1/world, the slab of `a_sym` N1 x N0/world
transpose_alltoall!(a_sym::SymBuffer, b_local::B200Array{T, 2}, N0::Integer, N1::Integer) where {T} =
    (@b200 b200_transpose_alltoall (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int32) a_sym.handle b_local.ptr N0 N1 Int32(sizeof(T)); nothing)

# CUDA graphs for launch-bound sequences: `g = record(backend) do ... launches ... end; replay(g)` (b200_graph_*).
mutable struct Graph
    handle::Ptr{Cvoid}
end
function record(f::Function, ::B200Backend)
    @b200 b200_graph_begin ()
    h = Ref{Ptr{Cvoid}}(C_NULL)
    try
        f()
    catch
        ccall((:b200_graph_end, libb200ka), Int32, (Ptr{Ptr{Cvoid}},), h)      # leave capture mode, drop the partial recording
        h[] != C_NULL && ccall((:b200_graph_destroy, libb200ka), Int32, (Ptr{Cvoid},), h[])
        rethrow()
    end
    @b200 b200_graph_end (Ptr{Ptr{Cvoid}},) h
    g = Graph(h[])
    finalizer(x -> ccall((:b200_graph_destroy, libb200ka), Int32, (Ptr{Cvoid},), x.handle), g)
    return g
end
replay(g::Graph) = (@b200 b200_graph_launch (Ptr{Cvoid},) g.handle; nothing)

# Run-time registration of a precompiled kernel from the user's own shared library (b200_register_kernel): `handler` is the
# address of a native `b200_kernel_handler`, e.g. `Libdl.dlsym(Libdl.dlopen("libmykernels.so"), :my_handler)`.  `name` is
# `"gpu_" * string(kernel function name)` for `@kernel`s (reference src/macros.jl:32), the function name for `KI.@kernel`.
function register_kernel(name::AbstractString, handler::Ptr{Cvoid}; ki::Bool = false, user::Ptr{Cvoid} = C_NULL)
    @b200 b200_register_kernel (Cstring, Int32, Ptr{Cvoid}, Ptr{Cvoid}) String(name) Int32(ki ? 1 : 0) handler user
    return nothing
end
unregister_kernel(name::AbstractString) = (@b200 b200_unregister_kernel (Cstring,) String(name); nothing)

function registered_kernels()
    n = Ref{Int32}(0)
    @b200 b200_kernel_count (Ptr{Int32},) n
    return [unsafe_string(ccall((:b200_kernel_name, libb200ka), Cstring, (Int32,), Int32(i))) for i in 0:(n[] - 1)]
end

# ---------------------------------------------------------------------------------------------------------
# Kernel launch (reference src/pocl/backend.jl:74-147)
# ---------------------------------------------------------------------------------------------------------
function KA.mkcontext(kernel::KA.Kernel{B200Backend}, _ndrange, iterspace)
    return KA.CompilerMetadata{KA.ndrange(kernel), KA.DynamicCheck}(_ndrange, iterspace)
end
function KA.mkcontext(kernel::KA.Kernel{B200Backend}, I, _ndrange, iterspace, ::Dynamic) where {Dynamic}
    return KA.CompilerMetadata{KA.ndrange(kernel), Dynamic}(I, _ndrange, iterspace)
end

function KA.launch_config(kernel::KA.Kernel{B200Backend}, ndrange, workgroupsize)
    if ndrange isa Integer
        ndrange = (ndrange,)
    end
    if workgroupsize isa Integer
        workgroupsize = (workgroupsize,)
    end
    # partition checked that the ndrange's agreed
    if KA.ndrange(kernel) <: KA.StaticSize
        ndrange = nothing
    end
    iterspace, dynamic = if KA.workgroupsize(kernel) <: KA.DynamicSize && workgroupsize === nothing
        KA.partition(kernel, ndrange, ndrange)          # use ndrange as preliminary workgroupsize
    else
        KA.partition(kernel, ndrange, workgroupsize)
    end
    return ndrange, workgroupsize, iterspace, dynamic
end

function threads_to_workgroupsize(threads, ndrange)       # reference src/pocl/backend.jl:108-115
    total = 1
    return map(ndrange) do n
        x = min(div(threads, total), n)
        total *= x
        return x
    end
end

const MAX_WORK_GROUP_SIZE = 1024

function (obj::KA.Kernel{B200Backend})(args...; ndrange = nothing, workgroupsize = nothing)
    ndrange, workgroupsize, iterspace, dynamic = KA.launch_config(obj, ndrange, workgroupsize)
    ctx = KA.mkcontext(obj, ndrange, iterspace)

    # figure out the workgroupsize automatically: greedy fill of the device maximum (:126-131)
    if KA.workgroupsize(obj) <: KA.DynamicSize && workgroupsize === nothing
        wg_size_nd = threads_to_workgroupsize(MAX_WORK_GROUP_SIZE, ndrange)
        iterspace, dynamic = KA.partition(obj, ndrange, wg_size_nd)
        ctx = KA.mkcontext(obj, ndrange, iterspace)
    end

    blocks = size(KA.blocks(iterspace))
    items = size(KA.workitems(iterspace))
    prod(blocks) == 0 && return nothing                    # groups == 0 (:136-138)
    N = length(blocks)
    N <= B200_MAX_NDIM || throw(ArgumentError("ndranges of more than $B200_MAX_NDIM dimensions are not supported"))
    prod(items) <= MAX_WORK_GROUP_SIZE || throw(ArgumentError("workgroupsize $items exceeds $MAX_WORK_GROUP_SIZE work-items"))

    full_ndrange = size(KA.__ndrange(ctx))
    flags = Int32((KA.ndrange(obj) <: KA.StaticSize ? 1 : 0) | (KA.workgroupsize(obj) <: KA.StaticSize ? 2 : 0))
    # the flattened CompilerMetadata; DynamicCheck always, like the in-tree backend (:74-76)
    desc = LaunchDesc(0, N, pad4(full_ndrange), pad4(items), pad4(blocks), 1, flags)
    launch(String(nameof(obj.f)), desc, args)               # "gpu_<kernel name>" (src/macros.jl:32)
    return nothing
end

# ---------------------------------------------------------------------------------------------------------
# KernelIntrinsics (reference src/pocl/backend.jl:149-179)
# ---------------------------------------------------------------------------------------------------------
struct B200Function
    name::String
end

kernel_name(f::Function) = String(nameof(f))
kernel_name(f::B200Function) = f.name
kernel_name(f::Union{AbstractString, Symbol}) = String(f)

# `f` is the kernel function itself (see KI.argconvert above); `name` is accepted for interface parity but cannot rename a
# precompiled kernel, so the catalogue lookup always uses the function's own name.
function KI.kernel_function(::B200Backend, f::F, tt::TT = Tuple{}; name = nothing, kwargs...) where {F, TT}
    kname = kernel_name(f)
    out = Ref{Int64}(0)
    @b200 b200_kernel_max_work_group_size (Cstring, Int64, Ptr{Int64}) kname Int64(MAX_WORK_GROUP_SIZE) out   # existence check
    return KI.Kernel{B200Backend, B200Function}(B200Backend(), B200Function(kname))
end

function (obj::KI.Kernel{B200Backend})(args...; numworkgroups = 1, workgroupsize = 1)
    KI.check_launch_args(numworkgroups, workgroupsize)       # ArgumentError for > 3 dims (src/intrinsics.jl:182-188)
    ng = Tuple(numworkgroups)
    ws = Tuple(workgroupsize)
    nd = max(length(ng), length(ws))
    desc = LaunchDesc(1, nd, pad4(ntuple(i -> (i <= length(ng) ? ng[i] : 1) * (i <= length(ws) ? ws[i] : 1), nd)),
                      pad4(ws), pad4(ng), 0, 0)
    launch(obj.kern.name, desc, args)
    return nothing
end

function KI.kernel_max_work_group_size(kernel::KI.Kernel{<:B200Backend}; max_work_items::Int = typemax(Int))::Int
    out = Ref{Int64}(0)
    @b200 b200_kernel_max_work_group_size (Cstring, Int64, Ptr{Int64}) kernel.kern.name Int64(max_work_items) out
    return Int(out[])
end
function KI.max_work_group_size(::B200Backend)::Int
    v = Ref{Int32}(0)
    @b200 b200_max_threads_per_block (Ptr{Int32},) v
    return Int(v[])
end
function KI.multiprocessor_count(::B200Backend)::Int
    v = Ref{Int32}(0)
    @b200 b200_sm_count (Ptr{Int32},) v
    return Int(v[])
end

# Device-side overrides (KI.get_*_id, KI.localmemory, KI.barrier, KA.__validindex, KA.Scratchpad; reference
# src/pocl/backend.jl:183-237) have no Julia twin: they live in csrc/ka_device.cuh inside the library.

end # module
