"""The finite kernel catalogue: one constructor per `@kernel` the reference defines on the hot path
(SURVEY.md 2.2).  `copy_kernel_` is Julia's `copy_kernel!` etc.  KI kernels are launched by name through
`ki_kernel(backend, "histogram_kernel!", ...)`."""
from .kernel import KernelFunction

# KA internals (reference src/KernelAbstractions.jl:869-877)
init_kernel = KernelFunction("init_kernel")
copy_kernel = KernelFunction("copy_kernel")
# examples/
copy_kernel_ = KernelFunction("copy_kernel!")                    # examples/memcopy.jl:4-7, memcopy_static.jl:4-7
naive_transpose_kernel_ = KernelFunction("naive_transpose_kernel!")  # examples/naive_transpose.jl:4-7
matmul_kernel_ = KernelFunction("matmul_kernel!")                # examples/matmul.jl:5-15
saxpy_kernel_ = KernelFunction("saxpy_kernel!")                  # benchmark/benchmarks.jl:29-32
saxpy_kernel = KernelFunction("saxpy_kernel")                    # examples/numa_aware.jl:10-13 (a, X, Y) in place
# examples/performance.jl:17-138
simple_copy_kernel_ = KernelFunction("simple_copy_kernel!")
simple_transpose_kernel_ = KernelFunction("simple_transpose_kernel!")
lmem_copy_kernel_ = KernelFunction("lmem_copy_kernel!")
lmem_transpose_kernel_ = KernelFunction("lmem_transpose_kernel!")
coalesced_copy_kernel_ = KernelFunction("coalesced_copy_kernel!")
coalesced_transpose_kernel_ = KernelFunction("coalesced_transpose_kernel!")
# test/localmem.jl
localmem = KernelFunction("localmem")
localmem2 = KernelFunction("localmem2")
localmem_unsafe_indices = KernelFunction("localmem_unsafe_indices")
many_localmem = KernelFunction("many_localmem")
# test/private.jl
stmt_form = KernelFunction("stmt_form")
typetest = KernelFunction("typetest")
private = KernelFunction("private")
forloop = KernelFunction("forloop")
reduce_private = KernelFunction("reduce_private")
# test/unroll.jl
kernel_unroll_ = KernelFunction("kernel_unroll!")
kernel_unroll2_ = KernelFunction("kernel_unroll2!")
# test/test.jl
index_linear_global = KernelFunction("index_linear_global")
index_linear_local = KernelFunction("index_linear_local")
index_linear_group = KernelFunction("index_linear_group")
index_cartesian_global = KernelFunction("index_cartesian_global")
index_cartesian_local = KernelFunction("index_cartesian_local")
index_cartesian_group = KernelFunction("index_cartesian_group")
constarg = KernelFunction("constarg")
kernel_val_ = KernelFunction("kernel_val!")
kernel_empty = KernelFunction("kernel_empty")
context_kernel = KernelFunction("context_kernel")                        # test/test.jl:275-282
gpu_return_kernel_ = KernelFunction("gpu_return_kernel!")                # test/test.jl:307-313
unaliased_accumulate_ = KernelFunction("unaliased_accumulate!")          # test/test.jl:339-345
unaliased_accumulate_local_ = KernelFunction("unaliased_accumulate_local!")  # test/test.jl:347-356
convert_kernel_ = KernelFunction("convert_kernel!")                      # test/convert.jl:5-44
gamma_knl = KernelFunction("gamma_knl")                                  # test/specialfunctions.jl:5-18
erf_knl = KernelFunction("erf_knl")
erfc_knl = KernelFunction("erfc_knl")
