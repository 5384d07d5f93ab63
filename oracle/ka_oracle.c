/*
 * ka_oracle.c -- CPU restatement of the KernelAbstractions.jl hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load it.  It is never on the product path (the product path is
 * libb200ka.so and fails loudly without a GPU).
 *
 * Parity status: PINNED by the known-answer values the reference's own tests hold for this path
 * (tests/test_oracle_golden.py replays them).  The reference itself (Julia + PoCL)
 * cannot be executed in this environment, so there are no reference-generated fixtures.
 *
 * Everything is 1-based and column-major exactly like the Julia source; `Int` is int64_t.
 * Work-groups run in parallel over OpenMP threads; inside a group the work-items of one
 * @synchronize-delimited region all run before the next region starts, which is the structure
 * `split`/`emit` give a kernel body (reference src/macros.jl:113-196, docs/src/index.md:96-113).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t Int;
#define KA_MAX_NDIM 4

/* Flattened CompilerMetadata + NDRange (reference src/compiler.jl:1-33, src/nditeration.jl:49-60). */
typedef struct {
    int32_t ndim;
    Int ndrange[KA_MAX_NDIM];
    Int wgs[KA_MAX_NDIM];
    Int blocks[KA_MAX_NDIM];
    int32_t dynamic_check;
} ka_ctx;

int ka_oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* launchers such as torchrun export OMP_NUM_THREADS=1; the timed CPU baseline sets its thread count explicitly */
void ka_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------
 * NDIteration.partition  (reference src/nditeration.jl:143-164)
 *   workgroupsize padded with ones up to length(ndrange) (a 0 entry also becomes 1);
 *   blocks[d] = fld1(ndrange[d], wgs[d]);  dynamic iff any mod(ndrange[d], wgs[d]) != 0.
 * Returns -1 when length(wgs) > length(ndrange) (the @assert at :144).
 */
static Int fld1(Int x, Int y) { /* Julia fld1 for y > 0: fld(x - 1, y) + 1 with floor division */
    Int a = x - 1;
    Int q = a / y;
    if ((a % y != 0) && ((a < 0) != (y < 0))) q -= 1;
    return q + 1;
}

int ka_partition(int32_t n, const Int *ndrange, int32_t nw, const Int *wgs_in, Int *blocks, Int *wgs,
                 int32_t *dynamic) {
    if (nw > n) return -1;
    int dyn = 0;
    for (int d = 0; d < n; ++d) {
        Int w = (d >= nw || wgs_in[d] == 0) ? 1 : wgs_in[d];
        wgs[d] = w;
        Int r = ndrange[d] % w;
        if (r < 0) r += w;
        dyn |= (r != 0);
        blocks[d] = fld1(ndrange[d], w);
    }
    *dynamic = dyn;
    return 0;
}

/* threads_to_workgroupsize (reference src/pocl/backend.jl:108-115): greedy fill of dim 1, then 2, ... */
void ka_threads_to_workgroupsize(Int threads, int32_t n, const Int *ndrange, Int *out) {
    Int total = 1;
    for (int d = 0; d < n; ++d) {
        Int x = threads / total;
        if (ndrange[d] < x) x = ndrange[d];
        total *= x;
        out[d] = x;
    }
}

/* Column-major unravel of a 1-based linear index (CartesianIndices(dims)[lin]). */
static inline void unravel1(Int lin, int n, const Int *dims, Int *out) {
    Int r = lin - 1;
    for (int d = 0; d < n; ++d) {
        Int s = dims[d];
        out[d] = (s > 0 ? r % s : 0) + 1;
        r = s > 0 ? r / s : 0;
    }
}

/* expand(ndrange, groupidx::Integer, idx::Integer)  (reference src/nditeration.jl:74-82,115-126):
 * I_d = (g_d - 1) * wgs_d + l_d  with g = blocks[groupidx], l = workitems[idx]. */
void ka_expand(const ka_ctx *c, Int g, Int l, Int *I) {
    Int gI[KA_MAX_NDIM], lI[KA_MAX_NDIM];
    unravel1(g, c->ndim, c->blocks, gI);
    unravel1(l, c->ndim, c->wgs, lI);
    for (int d = 0; d < c->ndim; ++d) I[d] = (gI[d] - 1) * c->wgs[d] + lI[d];
}
void ka_group_cartesian(const ka_ctx *c, Int g, Int *I) { unravel1(g, c->ndim, c->blocks, I); }
void ka_local_cartesian(const ka_ctx *c, Int l, Int *I) { unravel1(l, c->ndim, c->wgs, I); }

/* __validindex (reference src/pocl/backend.jl:207-214): `expand(...) in ndrange` under DynamicCheck. */
int ka_validindex(const ka_ctx *c, Int g, Int l) {
    if (!c->dynamic_check) return 1;
    Int I[KA_MAX_NDIM];
    ka_expand(c, g, l, I);
    for (int d = 0; d < c->ndim; ++d)
        if (I[d] < 1 || I[d] > c->ndrange[d]) return 0;
    return 1;
}

static inline Int prodn(const Int *v, int n) {
    Int p = 1;
    for (int d = 0; d < n; ++d) p *= v[d];
    return p;
}
Int ka_num_groups(const ka_ctx *c) { return prodn(c->blocks, c->ndim); }
Int ka_group_items(const ka_ctx *c) { return prodn(c->wgs, c->ndim); }

/* @index(Global, Linear) = KI.get_global_id().x of the 1-D launch of groups*items work-items
 * (reference src/KernelAbstractions.jl:499-501, src/pocl/backend.jl:133-143). */
static inline Int global_linear(const ka_ctx *c, Int g, Int l) { return (g - 1) * ka_group_items(c) + l; }

/* Fill `out` (length groups*items*ndim, ndim fastest) with expand(g, l) for every linear g, l:
 * the `linear_iteration` helper of the reference test (test/nditeration.jl:20-30). */
void ka_linear_iteration(const ka_ctx *c, Int *out) {
    Int G = ka_num_groups(c), L = ka_group_items(c);
    for (Int g = 1; g <= G; ++g)
        for (Int l = 1; l <= L; ++l) ka_expand(c, g, l, out + ((g - 1) * L + (l - 1)) * c->ndim);
}

/* ==========================================================================================
 * KA @kernel twins.  Each is the body produced by transform_gpu! (reference src/macros.jl:55-88):
 *   __active_lane__ = __validindex(ctx); regions guarded by it; barriers between regions.
 * ========================================================================================== */

/* copy_kernel / copy_kernel!: A[I] = B[I], I = @index(Global)
 * (reference src/KernelAbstractions.jl:874-877, examples/memcopy.jl:4-7, memcopy_static.jl:4-7). */
void orc_copy_kernel(const ka_ctx *c, void *A, const void *B, Int elsize) {
    Int G = ka_num_groups(c), L = ka_group_items(c);
#pragma omp parallel for schedule(static)
    for (Int g = 1; g <= G; ++g)
        for (Int l = 1; l <= L; ++l) {
            if (!ka_validindex(c, g, l)) continue;
            Int I = global_linear(c, g, l);
            memcpy((char *)A + (I - 1) * elsize, (const char *)B + (I - 1) * elsize, (size_t)elsize);
        }
}

/* init_kernel: arr[I] = f(T)  (reference src/KernelAbstractions.jl:869-872); value = f(T) bits. */
void orc_init_kernel(const ka_ctx *c, void *arr, const void *value, Int elsize) {
    Int G = ka_num_groups(c), L = ka_group_items(c);
#pragma omp parallel for schedule(static)
    for (Int g = 1; g <= G; ++g)
        for (Int l = 1; l <= L; ++l) {
            if (!ka_validindex(c, g, l)) continue;
            Int I = global_linear(c, g, l);
            memcpy((char *)arr + (I - 1) * elsize, value, (size_t)elsize);
        }
}

/* naive_transpose_kernel!: (i, j) = @index(Global, NTuple); a[i, j] = b[j, i]
 * (reference examples/naive_transpose.jl:4-7).  rows_a = size(a,1), rows_b = size(b,1). */
void orc_naive_transpose(const ka_ctx *c, void *a, const void *b, Int rows_a, Int rows_b, Int elsize) {
    Int G = ka_num_groups(c), L = ka_group_items(c);
#pragma omp parallel for schedule(static)
    for (Int g = 1; g <= G; ++g)
        for (Int l = 1; l <= L; ++l) {
            if (!ka_validindex(c, g, l)) continue;
            Int I[KA_MAX_NDIM];
            ka_expand(c, g, l, I);
            Int i = I[0], j = I[1];
            memcpy((char *)a + ((i - 1) + (j - 1) * rows_a) * elsize,
                   (const char *)b + ((j - 1) + (i - 1) * rows_b) * elsize, (size_t)elsize);
        }
}

/* Same kernel, as a work-group-vectorising CPU compiler (PoCL's work-item loops) would run it: the group's
 * Cartesian index is computed once per group, the work-items of the group become the inner loops.  Identical
 * results to orc_naive_transpose (tests/test_oracle_golden.py checks it); used as the timed CPU baseline so the
 * baseline is not slowed down by this file's generic per-item div/mod. 4-byte elements, 2-d ndrange. */
void orc_naive_transpose_wg_f32(const ka_ctx *c, float *a, const float *b, Int rows_a, Int rows_b) {
    Int G = ka_num_groups(c);
    const Int W0 = c->wgs[0], W1 = c->wgs[1], N0 = c->ndrange[0], N1 = c->ndrange[1];
#pragma omp parallel for schedule(static)
    for (Int g = 1; g <= G; ++g) {
        Int gI[KA_MAX_NDIM];
        ka_group_cartesian(c, g, gI);
        const Int i0 = (gI[0] - 1) * W0, j0 = (gI[1] - 1) * W1;
        for (Int lj = 1; lj <= W1; ++lj) {
            const Int j = j0 + lj;
            if (j > N1) continue;
            const Int imax = (i0 + W0 <= N0) ? W0 : N0 - i0;
            float *arow = a + (j - 1) * rows_a + i0;
            const float *bcol = b + (j - 1) + i0 * rows_b;
            for (Int li = 0; li < imax; ++li) arow[li] = bcol[li * rows_b];
        }
    }
}

/* Work-group-vectorised 1-d copy and saxpy, same idea: one bounds computation per group, the work-items of the group are the
 * inner (vectorisable) loop -- how a CPU OpenCL runtime executes `A[I] = B[I]` (examples/memcopy.jl:4-7) and
 * `Z[I] = a*X[I] + Y[I]` (benchmark/benchmarks.jl:29-32).  Identical results to orc_copy / orc_saxpy_f32 (checked in
 * tests/test_oracle_golden.py); used only as timed CPU baselines. */
void orc_copy_wg_f32(const ka_ctx *c, float *A, const float *B) {
    const Int G = ka_num_groups(c), W = c->wgs[0], N = c->ndrange[0];
#pragma omp parallel for schedule(static)
    for (Int g = 0; g < G; ++g) {
        const Int i0 = g * W, cnt = (i0 + W <= N) ? W : N - i0;
        for (Int l = 0; l < cnt; ++l) A[i0 + l] = B[i0 + l];
    }
}
void orc_saxpy_wg_f32(const ka_ctx *c, float *Z, float a, const float *X, const float *Y) {
    const Int G = ka_num_groups(c), W = c->wgs[0], N = c->ndrange[0];
#pragma omp parallel for schedule(static)
    for (Int g = 0; g < G; ++g) {
        const Int i0 = g * W, cnt = (i0 + W <= N) ? W : N - i0;
        for (Int l = 0; l < cnt; ++l) {
            const float p = a * X[i0 + l];
            Z[i0 + l] = p + Y[i0 + l];
        }
    }
}

/* matmul_kernel!: tmp = zero(T); for k: tmp += a[i,k]*b[k,j]; output[i,j] = tmp
 * (reference examples/matmul.jl:5-15).  No @simd / @fastmath: separate multiply and add roundings,
 * so this file must be compiled with -ffp-contract=off. */
void orc_matmul_naive_f32(const ka_ctx *c, float *out, const float *a, const float *b, Int M, Int K) {
    Int G = ka_num_groups(c), L = ka_group_items(c);
#pragma omp parallel for schedule(static)
    for (Int g = 1; g <= G; ++g)
        for (Int l = 1; l <= L; ++l) {
            if (!ka_validindex(c, g, l)) continue;
            Int I[KA_MAX_NDIM];
            ka_expand(c, g, l, I);
            Int i = I[0], j = I[1];
            float tmp = 0.0f;
            for (Int k = 1; k <= K; ++k) {
                float p = a[(i - 1) + (k - 1) * M] * b[(k - 1) + (j - 1) * K];
                tmp = tmp + p;
            }
            out[(i - 1) + (j - 1) * M] = tmp;
        }
}
void orc_matmul_naive_f64(const ka_ctx *c, double *out, const double *a, const double *b, Int M, Int K) {
    Int G = ka_num_groups(c), L = ka_group_items(c);
#pragma omp parallel for schedule(static)
    for (Int g = 1; g <= G; ++g)
        for (Int l = 1; l <= L; ++l) {
            if (!ka_validindex(c, g, l)) continue;
            Int I[KA_MAX_NDIM];
            ka_expand(c, g, l, I);
            Int i = I[0], j = I[1];
            double tmp = 0.0;
            for (Int k = 1; k <= K; ++k) {
                double p = a[(i - 1) + (k - 1) * M] * b[(k - 1) + (j - 1) * K];
                tmp = tmp + p;
            }
            out[(i - 1) + (j - 1) * M] = tmp;
        }
}

/* ---- examples/performance.jl:17-138: six copy/transpose variants ----------------------------------------
 * which: 0 simple_copy (:17-20)        output[I,J] = input[I,J]               (bounds-checked lanes)
 *        1 simple_transpose (:22-25)   output[J,I] = input[I,J]
 *        2 lmem_copy (:29-46)          unsafe_indices: tile[i,j] = input[I,J]; @synchronize; output[I,J] = tile[i,j]
 *        3 lmem_transpose (:48-76)     tile[i,j] = input[I,J]; @synchronize; pivot group; output[I',J'] = tile[j,i]
 *        4 coalesced_copy (:80-106)    groupsize (TILE_DIM, BLOCK_ROWS), TILE_DIM/BLOCK_ROWS elements per lane
 *        5 coalesced_transpose (:108-138)
 * tile is (N + BANK, M) resp. (TILE_DIM + BANK, TILE_DIM), column-major.  4-byte elements (const T = Float32, :9).
 * rows_out = size(output,1), rows_in = size(input,1).  The unsafe_indices kernels have no lane guard (macros.jl:72-76):
 * the caller guarantees the launch stays inside the arrays, as the script does (N = 2048 divisible by the tiles). */
void orc_performance_kernel(const ka_ctx *c, int which, uint32_t *output, const uint32_t *input, Int rows_out,
                            Int rows_in, Int bank) {
    Int G = ka_num_groups(c), L = ka_group_items(c);
    const Int N = c->wgs[0], M = c->wgs[1];
    if (which <= 1) {
#pragma omp parallel for schedule(static)
        for (Int g = 1; g <= G; ++g)
            for (Int l = 1; l <= L; ++l) {
                if (!ka_validindex(c, g, l)) continue;
                Int I[KA_MAX_NDIM];
                ka_expand(c, g, l, I);
                if (which == 0)
                    output[(I[0] - 1) + (I[1] - 1) * rows_out] = input[(I[0] - 1) + (I[1] - 1) * rows_in];
                else
                    output[(I[1] - 1) + (I[0] - 1) * rows_out] = input[(I[0] - 1) + (I[1] - 1) * rows_in];
            }
        return;
    }
    const Int trows = N + bank;                          /* tile leading dimension */
    const Int tcols = (which >= 4) ? N : M;              /* coalesced: (TILE_DIM + BANK, TILE_DIM) */
#pragma omp parallel for schedule(static)
    for (Int g = 1; g <= G; ++g) {
        uint32_t *tile = (uint32_t *)calloc((size_t)(trows * tcols), sizeof(uint32_t));
        Int gI[KA_MAX_NDIM];
        ka_group_cartesian(c, g, gI);
        const Int gi = gI[0], gj = gI[1];
        if (which == 2 || which == 3) {
            for (Int l = 1; l <= L; ++l) {
                Int lI[KA_MAX_NDIM];
                ka_local_cartesian(c, l, lI);
                const Int i = lI[0], j = lI[1];
                const Int I = (gi - 1) * N + i, J = (gj - 1) * M + j;
                tile[(i - 1) + (j - 1) * trows] = input[(I - 1) + (J - 1) * rows_in];
            }
            /* @synchronize */
            for (Int l = 1; l <= L; ++l) {
                Int lI[KA_MAX_NDIM];
                ka_local_cartesian(c, l, lI);
                const Int i = lI[0], j = lI[1];
                if (which == 2) {
                    const Int I = (gi - 1) * N + i, J = (gj - 1) * M + j;
                    output[(I - 1) + (J - 1) * rows_out] = tile[(i - 1) + (j - 1) * trows];
                } else {
                    const Int I = (gj - 1) * M + i, J = (gi - 1) * N + j;   /* pivoted group index */
                    output[(I - 1) + (J - 1) * rows_out] = tile[(j - 1) + (i - 1) * trows];
                }
            }
        } else {
            const Int TILE = N, ROWS = M;
            for (Int l = 1; l <= L; ++l) {
                Int lI[KA_MAX_NDIM];
                ka_local_cartesian(c, l, lI);
                const Int i = lI[0], j = lI[1];
                const Int I = (gi - 1) * TILE + i, J = (gj - 1) * TILE + j;
                for (Int k = 0; k <= TILE - 1; k += ROWS)
                    tile[(i - 1) + (j + k - 1) * trows] = input[(I - 1) + (J + k - 1) * rows_in];
            }
            /* @synchronize */
            for (Int l = 1; l <= L; ++l) {
                Int lI[KA_MAX_NDIM];
                ka_local_cartesian(c, l, lI);
                const Int i = lI[0], j = lI[1];
                if (which == 4) {
                    const Int I = (gi - 1) * TILE + i, J = (gj - 1) * TILE + j;
                    for (Int k = 0; k <= TILE - 1; k += ROWS)
                        output[(I - 1) + (J + k - 1) * rows_out] = tile[(i - 1) + (j + k - 1) * trows];
                } else {
                    const Int I = (gj - 1) * TILE + i, J = (gi - 1) * TILE + j;   /* transposed block offsets */
                    for (Int k = 0; k <= TILE - 1; k += ROWS)
                        output[(I - 1) + (J + k - 1) * rows_out] = tile[(j + k - 1) + (i - 1) * trows];
                }
            }
        }
        free(tile);
    }
}

/* ---- test/localmem.jl:4-65 (lowering: SURVEY.md Appendix A) -------------------------------- */
/* variant 0 localmem, 1 localmem2, 2 localmem_unsafe_indices, 3 many_localmem; A::Int64[lenA] */
void orc_localmem(const ka_ctx *c, int variant, Int *A, Int lenA) {
    Int G = ka_num_groups(c), N = ka_group_items(c);
#pragma omp parallel for schedule(static)
    for (Int g = 1; g <= G; ++g) {
        Int *lmem = (Int *)malloc(sizeof(Int) * (size_t)N * 2);
        Int *lmem2 = lmem + N;
        if (variant == 2) { /* unsafe_indices: no guard, own bounds check (test/localmem.jl:37-48) */
            for (Int i = 1; i <= N; ++i) lmem[i - 1] = i;
            /* @synchronize */
            for (Int i = 1; i <= N; ++i) {
                Int I = (g - 1) * N + i;
                if (I <= lenA) A[I - 1] = lmem[N - i + 1 - 1];
            }
        } else if (variant == 0) {
            for (Int i = 1; i <= N; ++i)
                if (ka_validindex(c, g, i)) lmem[i - 1] = i;
            for (Int i = 1; i <= N; ++i)
                if (ka_validindex(c, g, i)) A[global_linear(c, g, i) - 1] = lmem[N - i + 1 - 1];
        } else if (variant == 1) {
            for (Int i = 1; i <= N; ++i)
                if (ka_validindex(c, g, i)) lmem[i - 1] = i + 3;
            for (Int j = 1; j <= 2; ++j) { /* loop header on all lanes, body guarded, barrier inside */
                for (Int i = 1; i <= N; ++i)
                    if (ka_validindex(c, g, i)) lmem[i - 1] -= j;
            }
            for (Int i = 1; i <= N; ++i)
                if (ka_validindex(c, g, i)) A[global_linear(c, g, i) - 1] = lmem[N - i + 1 - 1];
        } else {
            for (Int i = 1; i <= N; ++i)
                if (ka_validindex(c, g, i)) {
                    lmem[i - 1] = i - 1;
                    lmem2[i - 1] = 1;
                }
            for (Int i = 1; i <= N; ++i)
                if (ka_validindex(c, g, i))
                    A[global_linear(c, g, i) - 1] = lmem[N - i + 1 - 1] + lmem2[N - i + 1 - 1];
        }
        free(lmem);
    }
}

/* ---- test/private.jl:4-59 ------------------------------------------------------------------- */
/* private(A): priv[1] = N - i + 1; @synchronize; A[I] = priv[1]   (test/private.jl:18-28) */
void orc_private(const ka_ctx *c, Int *A) {
    Int G = ka_num_groups(c), N = ka_group_items(c);
#pragma omp parallel for schedule(static)
    for (Int g = 1; g <= G; ++g) {
        Int *priv = (Int *)malloc(sizeof(Int) * (size_t)N);
        for (Int i = 1; i <= N; ++i)
            if (ka_validindex(c, g, i)) priv[i - 1] = N - i + 1;
        for (Int i = 1; i <= N; ++i)
            if (ka_validindex(c, g, i)) A[global_linear(c, g, i) - 1] = priv[i - 1];
        free(priv);
    }
}

/* typetest(A, B): B[I] = eltype(priv) === eltype(A)  -> true  (test/private.jl:10-16) */
void orc_typetest(const ka_ctx *c, uint8_t *B) {
    Int G = ka_num_groups(c), N = ka_group_items(c);
#pragma omp parallel for schedule(static)
    for (Int g = 1; g <= G; ++g)
        for (Int i = 1; i <= N; ++i)
            if (ka_validindex(c, g, i)) B[global_linear(c, g, i) - 1] = 1;
}

/* forloop(A, Val(N)) (test/private.jl:31-45): A is rowsA x N Int64; one work-item per row. */
void orc_forloop(const ka_ctx *c, Int *A, Int rowsA, Int N) {
    Int G = ka_num_groups(c), W = ka_group_items(c);
    for (Int g = 1; g <= G; ++g) { /* groups touch the same A[k,1] cells: keep serial like one queue */
        Int *priv = (Int *)malloc(sizeof(Int) * (size_t)(W * N));
        for (Int i = 1; i <= W; ++i) {
            if (!ka_validindex(c, g, i)) continue;
            Int I = global_linear(c, g, i);
            for (Int j = 1; j <= N; ++j) priv[(i - 1) * N + (j - 1)] = A[(I - 1) + (j - 1) * rowsA];
            A[(I - 1)] = 0;
        }
        for (Int j = 1; j <= N; ++j) {
            for (Int i = 1; i <= W; ++i) {
                if (!ka_validindex(c, g, i)) continue;
                Int k = ((j + i - 1 - 1) % N + N) % N + 1; /* mod1(j + i - 1, N) */
                A[(k - 1)] += priv[(i - 1) * N + (j - 1)];
            }
        }
        free(priv);
    }
}

/* reduce_private(out, A): priv[1] = sum_k A[I..., k]; out[I...] = priv[1]  (test/private.jl:47-59)
 * A is (n x K) Float32 with a 1-D ndrange (n,). */
void orc_reduce_private_f32(const ka_ctx *c, float *out, const float *A, Int n, Int K) {
    Int G = ka_num_groups(c), W = ka_group_items(c);
#pragma omp parallel for schedule(static)
    for (Int g = 1; g <= G; ++g)
        for (Int l = 1; l <= W; ++l) {
            if (!ka_validindex(c, g, l)) continue;
            Int I[KA_MAX_NDIM];
            ka_expand(c, g, l, I);
            float p = 0.0f;
            for (Int k = 1; k <= K; ++k) p = p + A[(I[0] - 1) + (k - 1) * n];
            out[I[0] - 1] = p;
        }
}

/* ---- test/unroll.jl:5-37 -------------------------------------------------------------------- */
/* variant 0: a[i] = i, i in 1:5; variant 1: a[i-5] = i for i in 6:N+5; variant 2: kernel_unroll2! */
void orc_unroll_f32(const ka_ctx *c, int variant, float *a, Int N) {
    Int G = ka_num_groups(c), W = ka_group_items(c);
    for (Int g = 1; g <= G; ++g)
        for (Int l = 1; l <= W; ++l) {
            if (!ka_validindex(c, g, l)) continue;
            if (variant == 0) {
                for (Int i = 1; i <= 5; ++i) a[i - 1] = (float)i;
            } else if (variant == 1) {
                for (Int i = 6; i <= N + 5; ++i) a[i - 5 - 1] = (float)i;
            } else {
                float va[3] = {1, 2, 3}, vb[3] = {3, 2, 1}, vc[9];
                Int I = global_linear(c, g, l);
                for (Int m = 1; m <= 3; ++m) {
                    for (int j = 1; j <= 3; ++j)
                        for (int i = 1; i <= 3; ++i) vc[0 + (j - 1) * 3] = (float)m * va[0] * vb[j - 1];
                    a[I - 1] = vc[0];
                }
            }
        }
}

/* ---- test/test.jl:46-73,116-160 index kernels ------------------------------------------------
 * which: 0 A[I]=I (Global Linear), 1 A[I]=local linear, 2 A[I]=group linear   (I = Global Linear)   */
void orc_index_linear(const ka_ctx *c, int which, Int *A) {
    Int G = ka_num_groups(c), W = ka_group_items(c);
#pragma omp parallel for schedule(static)
    for (Int g = 1; g <= G; ++g)
        for (Int l = 1; l <= W; ++l) {
            if (!ka_validindex(c, g, l)) continue;
            Int I = global_linear(c, g, l);
            A[I - 1] = which == 0 ? I : (which == 1 ? l : g);
        }
}
/* which: 0 A[I]=I, 1 A[I]=local cartesian, 2 A[I]=group cartesian, I = @index(Global, Cartesian);
 * A holds CartesianIndex{ndim} = ndim consecutive Int64; dimsA = size(A) (ndimA dims). */
void orc_index_cartesian(const ka_ctx *c, int which, Int *A, int32_t ndimA, const Int *dimsA) {
    Int G = ka_num_groups(c), W = ka_group_items(c);
    int n = c->ndim;
#pragma omp parallel for schedule(static)
    for (Int g = 1; g <= G; ++g)
        for (Int l = 1; l <= W; ++l) {
            if (!ka_validindex(c, g, l)) continue;
            Int I[KA_MAX_NDIM], v[KA_MAX_NDIM];
            ka_expand(c, g, l, I);
            if (which == 0) memcpy(v, I, sizeof(Int) * n);
            else if (which == 1) ka_local_cartesian(c, l, v);
            else ka_group_cartesian(c, g, v);
            /* A[I]: a CartesianIndex{n} into an ndimA-dim array; n == ndimA, or n == 1 (linear) */
            Int lin;
            if (n == ndimA) {
                lin = 0;
                Int stride = 1;
                for (int d = 0; d < n; ++d) {
                    lin += (I[d] - 1) * stride;
                    stride *= dimsA[d];
                }
            } else {
                lin = I[0] - 1;
            }
            memcpy(A + lin * n, v, sizeof(Int) * n);
        }
}

/* kernel_val!(a, Val(m)): a[I] = m  (test/test.jl:216-224) */
void orc_kernel_val(const ka_ctx *c, Int *a, Int m) {
    Int G = ka_num_groups(c), W = ka_group_items(c);
#pragma omp parallel for schedule(static)
    for (Int g = 1; g <= G; ++g)
        for (Int l = 1; l <= W; ++l)
            if (ka_validindex(c, g, l)) a[global_linear(c, g, l) - 1] = m;
}

/* context_kernel (test/test.jl:275-282): a[I] = 1;  gpu_return_kernel! (:307-313): if i <= length(x) / 2: x[i] = 1 */
void orc_misc_i64(const ka_ctx *c, int which, Int *a, Int len) {
    Int G = ka_num_groups(c), W = ka_group_items(c);
#pragma omp parallel for schedule(static)
    for (Int g = 1; g <= G; ++g)
        for (Int l = 1; l <= W; ++l) {
            if (!ka_validindex(c, g, l)) continue;
            Int I = global_linear(c, g, l);
            if (which == 0 || I <= len / 2) a[I - 1] = 1;
        }
}

/* unaliased_accumulate! / unaliased_accumulate_local! (test/test.jl:339-356): plain adds, k ascending from j to n.
 * Both forms produce the same bits when output starts at zero(T) (0 + x == x exactly). */
void orc_unaliased_accumulate_f32(const ka_ctx *c, int local, float *output, const float *input, Int n, Int rows_out,
                                  Int rows_in) {
    Int G = ka_num_groups(c), W = ka_group_items(c);
#pragma omp parallel for schedule(static)
    for (Int g = 1; g <= G; ++g)
        for (Int l = 1; l <= W; ++l) {
            if (!ka_validindex(c, g, l)) continue;
            Int I[KA_MAX_NDIM];
            ka_expand(c, g, l, I);
            const Int i = I[0], j = I[1];
            float acc = local ? 0.0f : output[(i - 1) + (j - 1) * rows_out];
            for (Int k = j; k <= n; ++k) acc = acc + input[(i - 1) + (k - 1) * rows_in];
            output[(i - 1) + (j - 1) * rows_out] = acc;
        }
}

/* ==========================================================================================
 * KI.@kernel twins: raw <=3-D launches, ids 1-based, no implicit bounds check
 * (reference src/intrinsics.jl:19-103,300-373; src/pocl/backend.jl:156-205).
 * ========================================================================================== */

/* create_histogram: the serial baseline of the reference (examples/histogram.jl:9-15). */
void orc_create_histogram_i32(int32_t *out, const int32_t *input, Int n) {
    for (Int t = 0; t < n; ++t) out[input[t] - 1] += 1;
}

/* histogram_kernel! (reference examples/histogram.jl:18-58), launched with workgroupsize = gs and
 * numworkgroups = cld(length(input), gs) (:63).  `out` has N bins and is accumulated into.
 * Atomics are integer adds: any interleaving gives the same counts, so the group loop is
 * parallelised with per-thread partial bins merged at the end. */
void orc_histogram_i32(int32_t *out, Int N, const int32_t *input, Int n, Int gs) {
    Int groups = (n + gs - 1) / gs;
#pragma omp parallel
    {
        int32_t *part = (int32_t *)calloc((size_t)(N > 0 ? N : 1), sizeof(int32_t));
        int32_t *shared = (int32_t *)malloc(sizeof(int32_t) * (size_t)gs);
#pragma omp for schedule(static)
        for (Int gid = 1; gid <= groups; ++gid) {
            for (Int min_element = 1; min_element <= N; min_element += gs) {
                for (Int lid = 1; lid <= gs; ++lid) shared[lid - 1] = 0;
                /* KI.barrier() */
                Int max_element = min_element + gs;
                if (max_element > N) max_element = N + 1;
                for (Int lid = 1; lid <= gs; ++lid) {
                    Int tid = (gid - 1) * gs + lid;
                    Int bin = tid <= n ? (Int)input[tid - 1] : 0;
                    if (bin >= min_element && bin < max_element) {
                        bin -= min_element - 1;
                        shared[bin - 1] += 1; /* @atomic */
                    }
                }
                /* KI.barrier() */
                for (Int lid = 1; lid <= gs; ++lid)
                    if (lid + min_element - 1 <= N)
                        part[lid + min_element - 1 - 1] += shared[lid - 1]; /* @atomic on global */
            }
        }
#pragma omp critical
        for (Int b = 0; b < N; ++b) out[b] += part[b];
        free(part);
        free(shared);
    }
}

/* coalesced_matmul_kernel! (reference examples/performant_matmul.jl:15-77): output is N x M,
 * input1 N x R, input2 R x M; TDIM x TDIM groups over (cld(N,TDIM), cld(M,TDIM)).
 * outval starts at -0.0 (:27); per tile `out = 0; for k in 1:TDIM out += t1[i,k]*t2[k,j]` (:59-62) then
 * `outval += out` (:63).  The @simd at :60 licenses reassociation/contraction, so the reference's own
 * rounding is unspecified; this oracle fixes k ascending with a fused multiply-add, which is what the
 * CUDA path computes bit for bit. */
void orc_coalesced_matmul_f32(float *output, const float *in1, const float *in2, Int N, Int R, Int M,
                              Int TDIM) {
    Int ngi = (N + TDIM - 1) / TDIM, ngj = (M + TDIM - 1) / TDIM;
    Int NUM_TILES = (R + TDIM - 1) / TDIM;
#pragma omp parallel
    {
        float *tile1 = (float *)malloc(sizeof(float) * (size_t)((TDIM + 1) * TDIM) * 2);
        float *tile2 = tile1 + (TDIM + 1) * TDIM;
        float *outval = (float *)malloc(sizeof(float) * (size_t)(TDIM * TDIM));
#pragma omp for collapse(2) schedule(static)
        for (Int gj = 1; gj <= ngj; ++gj)
            for (Int gi = 1; gi <= ngi; ++gi) {
                for (Int x = 0; x < TDIM * TDIM; ++x) outval[x] = -0.0f;
                for (Int t = 0; t < NUM_TILES; ++t) {
                    for (Int j = 1; j <= TDIM; ++j)
                        for (Int i = 1; i <= TDIM; ++i) {
                            Int I = (gi - 1) * TDIM + i, J = (gj - 1) * TDIM + j;
                            tile1[(i - 1) + (j - 1) * (TDIM + 1)] =
                                (I <= N && t * TDIM + j <= R) ? in1[(I - 1) + (t * TDIM + j - 1) * N] : 0.0f;
                            tile2[(i - 1) + (j - 1) * (TDIM + 1)] =
                                (t * TDIM + i <= R && J <= M) ? in2[(t * TDIM + i - 1) + (J - 1) * R] : 0.0f;
                        }
                    /* KI.barrier() */
                    for (Int j = 1; j <= TDIM; ++j)
                        for (Int i = 1; i <= TDIM; ++i) {
                            float out = 0.0f;
                            for (Int k = 1; k <= TDIM; ++k)
                                out = fmaf(tile1[(i - 1) + (k - 1) * (TDIM + 1)],
                                           tile2[(k - 1) + (j - 1) * (TDIM + 1)], out);
                            outval[(i - 1) + (j - 1) * TDIM] += out;
                        }
                    /* KI.barrier() */
                }
                for (Int j = 1; j <= TDIM; ++j)
                    for (Int i = 1; i <= TDIM; ++i) {
                        Int I = (gi - 1) * TDIM + i, J = (gj - 1) * TDIM + j;
                        if (I <= N && J <= M) output[(I - 1) + (J - 1) * N] = outval[(i - 1) + (j - 1) * TDIM];
                    }
            }
        free(tile1);
        free(outval);
    }
}

/* Same summation structure on operands rounded to BF16 / TF32 is produced by the Python side
 * (oracle/ka_oracle.py) rounding the inputs and calling orc_coalesced_matmul_f32. */

/* Running-accumulator FP32 GEMM (k ascending, fused multiply-add, acc starts at +0): the order of
 * b200_matmul_f32 mode B200_MATMUL_RUNNING.  Not a reference kernel; kept to pin that mode. */
void orc_matmul_running_f32(float *C, const float *A, const float *B, Int M, Int K, Int N) {
#pragma omp parallel for schedule(static)
    for (Int j = 0; j < N; ++j)
        for (Int i = 0; i < M; ++i) {
            float acc = 0.0f;
            for (Int k = 0; k < K; ++k) acc = fmaf(A[i + k * M], B[k + j * K], acc);
            C[i + j * M] = acc;
        }
}

/* FP64 ground truth for reporting the FP32 kernels' distance from exact arithmetic. */
void orc_matmul_f64_truth(double *C, const float *A, const float *B, Int M, Int K, Int N) {
#pragma omp parallel for schedule(static)
    for (Int j = 0; j < N; ++j)
        for (Int i = 0; i < M; ++i) {
            double acc = 0.0;
            for (Int k = 0; k < K; ++k) acc += (double)A[i + k * M] * (double)B[k + j * K];
            C[i + j * M] = acc;
        }
}

/* test_intrinsics_kernel (reference test/intrinsics.jl:11-25): six Int64 per element
 * {global_size, global_id, local_size, local_id, num_groups, group_id} (.x components). */
void orc_test_intrinsics(Int *results, Int len, Int numworkgroups, Int workgroupsize) {
    for (Int g = 1; g <= numworkgroups; ++g)
        for (Int l = 1; l <= workgroupsize; ++l) {
            Int i = (g - 1) * workgroupsize + l;
            if (i <= len) {
                Int *r = results + (i - 1) * 6;
                r[0] = numworkgroups * workgroupsize;
                r[1] = i;
                r[2] = workgroupsize;
                r[3] = l;
                r[4] = numworkgroups;
                r[5] = g;
            }
        }
}

/* launch_kernel1d/2d/3d (reference test/intrinsics.jl:31-72):
 * arr[(gi-1)*ngi + i, (gj-1)*ngj + j, (gk-1)*ngk + k] = 1f0 with ng* = KI.get_num_groups(). */
void orc_launch_kernel_nd(float *arr, const Int *dims, const Int *numgroups, const Int *wgs) {
    for (Int gk = 1; gk <= numgroups[2]; ++gk)
        for (Int gj = 1; gj <= numgroups[1]; ++gj)
            for (Int gi = 1; gi <= numgroups[0]; ++gi)
                for (Int k = 1; k <= wgs[2]; ++k)
                    for (Int j = 1; j <= wgs[1]; ++j)
                        for (Int i = 1; i <= wgs[0]; ++i) {
                            Int x = (gi - 1) * numgroups[0] + i, y = (gj - 1) * numgroups[1] + j,
                                z = (gk - 1) * numgroups[2] + k;
                            arr[(x - 1) + (y - 1) * dims[0] + (z - 1) * dims[0] * dims[1]] = 1.0f;
                        }
}

/* ---- examples/performance.jl:17-138 and benchmark/benchmarks.jl:29-32 ("next" rows) ----------- */
/* saxpy_kernel!: Z[I] = a*X[I] + Y[I]  (reference benchmark/benchmarks.jl:29-32; muladd allowed by
 * `@inbounds Z[I] = a * X[I] + Y[I]`? No: plain * and + in Julia are not contracted) */
void orc_saxpy_f32(const ka_ctx *c, float *Z, float a, const float *X, const float *Y) {
    Int G = ka_num_groups(c), W = ka_group_items(c);
#pragma omp parallel for schedule(static)
    for (Int g = 1; g <= G; ++g)
        for (Int l = 1; l <= W; ++l) {
            if (!ka_validindex(c, g, l)) continue;
            Int I = global_linear(c, g, l);
            float p = a * X[I - 1];
            Z[I - 1] = p + Y[I - 1];
        }
}

void orc_saxpy_f64(const ka_ctx *c, double *Z, double a, const double *X, const double *Y) {
    Int G = ka_num_groups(c), W = ka_group_items(c);
#pragma omp parallel for schedule(static)
    for (Int g = 1; g <= G; ++g)
        for (Int l = 1; l <= W; ++l) {
            if (!ka_validindex(c, g, l)) continue;
            Int I = global_linear(c, g, l);
            double p = a * X[I - 1];
            Z[I - 1] = p + Y[I - 1];
        }
}

/* Deterministic input generators shared by the tests and bench.py: a counter hash so
 * the same stream can be produced on the GPU side by the test harness with numpy. */
static inline uint32_t mix32(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x = x ^ (x >> 31);
    return (uint32_t)(x >> 32);
}
void orc_gen_uniform_f32(float *out, Int n, uint64_t seed) {
#pragma omp parallel for schedule(static)
    for (Int i = 0; i < n; ++i) out[i] = (float)(mix32(seed * 0x100000001B3ull + (uint64_t)i) >> 8) * (1.0f / 16777216.0f);
}
void orc_gen_bins_i32(int32_t *out, Int n, uint64_t seed, int32_t nbins) {
#pragma omp parallel for schedule(static)
    for (Int i = 0; i < n; ++i) out[i] = (int32_t)(mix32(seed * 0x100000001B3ull + (uint64_t)i) % (uint32_t)nbins) + 1;
}
