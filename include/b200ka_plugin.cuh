/* b200ka_plugin.cuh -- for authors of kernels registered with b200_register_kernel (include/b200ka.h).
 *
 * A KA `@kernel` launch is one-dimensional on the device: prod(blocks) CTAs of prod(workgroupsize) threads; the N-d indices
 * are recovered per work-item (reference src/pocl/backend.jl:133-143, src/nditeration.jl:74-82,115-126).  This header gives a
 * plug-in the same building blocks the built-in kernel twins use:
 *   device:  b200::KaCtx, ka_global_linear(c) / ka_local_linear(c) / ka_group_linear(c), ka_expand(c, I), ka_valid(c)
 *            (__validindex), ka_group_cartesian / ka_local_cartesian -- strength-reduced: no integer division on the device
 *   host:    b200_plugin_ctx(desc, &ctx, &groups, &items)  -- flatten a b200_launch_desc of kind B200_LAUNCH_KA, choose the
 *            lowering (native <=3-d grid, or 1-d launch with multiply-high unravelling) and precompute its constants;
 *            b200::ka_launch_dims(ctx, &grid, &block) -- the launch geometry that goes with it (ALWAYS launch with these)
 * A handler looks like
 *   extern "C" int32_t my_handler(const b200_launch_desc *d, const b200_arg *a, int32_t n, void *stream, void *user) {
 *       b200::KaCtx c; long long groups, items; dim3 grid, block;
 *       if (b200_plugin_ctx(d, &c, &groups, &items) != B200_OK || n != 2 || a[0].eltype != B200_T_F32) return B200_ERR_BAD_ARGS;
 *       b200::ka_launch_dims(c, &grid, &block);
 *       if (groups) my_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(c, (float *)a[0].ptr, (float)a[1].fscalar);
 *       return B200_OK;
 *   }
 */
#pragma once
#include "b200ka.h"
#include "../kernelabstractions.jl_b200/csrc/ka_device.cuh"

static inline int32_t b200_plugin_ctx(const b200_launch_desc *d, b200::KaCtx *c, long long *groups, long long *items) {
    if (!d || d->kind != B200_LAUNCH_KA || d->ndim < 1 || d->ndim > b200::KA_MAX_NDIM) return B200_ERR_BAD_ARGS;
    c->ndim = d->ndim;
    c->dynamic_check = d->dynamic_check;
    *groups = 1;
    *items = 1;
    for (int k = 0; k < b200::KA_MAX_NDIM; ++k) {
        const bool in = k < d->ndim;
        c->ndrange[k] = in ? d->ndrange[k] : 1;
        c->wgs[k] = in ? d->wgs[k] : 1;
        c->blocks[k] = in ? d->blocks[k] : 1;
        *groups *= c->blocks[k];
        *items *= c->wgs[k];
    }
    if (*items < 1 || *items > 1024 || *groups > 2147483647LL) return B200_ERR_LAUNCH_DIMS;
    b200::ka_ctx_finish(c);
    return B200_OK;
}
