# Runs the reference's backend-parametric testsuite against B200Backend (reference test/testsuite.jl:43-97).
# NOTE: not executable in this repository's build environment (no Julia toolchain); the Python mirror
# (tests/test_gpu_testsuite.py) replays the same testsets over the same C ABI.
using Test
using KernelAbstractions
using B200BackendModule

include(joinpath(pkgdir(KernelAbstractions), "test", "testsuite.jl"))

@testset "KernelAbstractions on B200Backend" begin
    if KernelAbstractions.functional(B200Backend())
        Testsuite.testsuite(
            B200Backend, "B200", B200BackendModule, B200Array, B200Array;
            skip_tests = Set([
                "Const", "Reflection", "Printing", "sparse",       # need Julia IR / device printf / SparseArrays adaptors
                "fallback test: callable types",                     # a kernel defined on a callable type has no name to dispatch on
            ]),
        )
    else
        @test_skip "no B200 available"
    end
end
