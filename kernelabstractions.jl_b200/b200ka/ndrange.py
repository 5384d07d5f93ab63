"""Iteration-space host logic: the Python mirror of KernelAbstractions.NDIteration and KA.partition.

In the Julia host layer these functions are NOT re-implemented -- B200Backend.jl calls the unchanged
`KA.partition` / `NDIteration.partition` (reference src/KernelAbstractions.jl:751-806,
src/nditeration.jl:143-164).  This module exists so the same logic can be exercised (and parity-tested against
the oracle) from Python, where the testsuite mirror lives.  Indices are 1-based like the reference.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence, Tuple, Union

from ._lib import ErrorException


class DynamicSize:
    """reference src/nditeration.jl:13"""

    def __repr__(self):
        return "DynamicSize"


@dataclass(frozen=True)
class StaticSize:
    """reference src/nditeration.jl:14-29"""

    S: Tuple[int, ...]

    def __init__(self, *s):
        if len(s) == 1 and isinstance(s[0], (tuple, list)):
            s = tuple(s[0])
        object.__setattr__(self, "S", tuple(int(x) for x in s))

    def __getitem__(self, i: int) -> int:  # 1-based; beyond the rank -> 1 (:26)
        return self.S[i - 1] if i <= len(self.S) else 1

    def __len__(self) -> int:
        p = 1
        for x in self.S:
            p *= x
        return p

    def ndims(self) -> int:
        return len(self.S)


DYNAMIC = DynamicSize()


class DynamicCheck:
    pass


class NoDynamicCheck:
    pass


def _fld1(x: int, y: int) -> int:
    return (x - 1) // y + 1


def nd_partition(ndrange: Sequence[int], workgroupsize: Sequence[int]):
    """NDIteration.partition (reference src/nditeration.jl:143-164) -> (blocks, workgroupsize, dynamic)."""
    assert len(workgroupsize) <= len(ndrange)
    wgs = tuple(
        1 if (i >= len(workgroupsize) or workgroupsize[i] == 0) else int(workgroupsize[i]) for i in range(len(ndrange)))
    dynamic = False
    blocks = []
    for n, w in zip(ndrange, wgs):
        dynamic |= (n % w) != 0
        blocks.append(_fld1(int(n), w))
    return tuple(blocks), wgs, (DynamicCheck() if dynamic else NoDynamicCheck())


@dataclass
class NDRange:
    """reference src/nditeration.jl:49-72: a blocked iteration space (blocks x workitems)."""

    N: int
    blocks: Tuple[int, ...]
    workitems: Tuple[int, ...]
    static_blocks: bool = False
    static_workitems: bool = False

    def __len__(self) -> int:
        p = 1
        for b in self.blocks:
            p *= b
        return p

    def nitems(self) -> int:
        p = 1
        for w in self.workitems:
            p *= w
        return p

    def ndims(self) -> int:
        return self.N


def _unravel1(lin: int, dims: Sequence[int]) -> Tuple[int, ...]:
    r = lin - 1
    out = []
    for s in dims:
        out.append(r % s + 1)
        r //= s
    return tuple(out)


def expand(nd: NDRange, groupidx: Union[int, Sequence[int]], idx: Union[int, Sequence[int]]) -> Tuple[int, ...]:
    """expand (reference src/nditeration.jl:74-82,115-134): I_d = (g_d - 1) * wgs_d + l_d."""
    g = _unravel1(groupidx, nd.blocks) if isinstance(groupidx, int) else tuple(groupidx)
    l = _unravel1(idx, nd.workitems) if isinstance(idx, int) else tuple(idx)
    return tuple((g[d] - 1) * nd.workitems[d] + l[d] for d in range(nd.N))


def _as_tuple(x) -> Optional[Tuple[int, ...]]:
    if x is None:
        return None
    if isinstance(x, (int,)) or hasattr(x, "__index__"):
        return (int(x),)
    return tuple(int(v) for v in x)


def ka_partition(static_ndrange, static_workgroupsize, ndrange, workgroupsize):
    """KA.partition(kernel, ndrange, workgroupsize) (reference src/KernelAbstractions.jl:751-806).

    `static_*` are the kernel's type parameters (StaticSize or DynamicSize); returns (NDRange, dynamic).
    Raises ErrorException exactly where the reference calls `error(...)` (:768,773,780)."""
    ndrange = _as_tuple(ndrange)
    workgroupsize = _as_tuple(workgroupsize)
    if (ndrange is None and isinstance(static_ndrange, DynamicSize)) or \
            (workgroupsize is None and isinstance(static_workgroupsize, DynamicSize)):
        raise ErrorException(
            "Can not partition kernel!\n\nYou created a dynamically sized kernel, but forgot to provide runtime\n"
            "parameters for the kernel. Either provide them statically if known\nor dynamically.\n"
            f"NDRange(Static):  {static_ndrange}\nNDRange(Dynamic): {ndrange}\n"
            f"Workgroupsize(Static):  {static_workgroupsize}\nWorkgroupsize(Dynamic): {workgroupsize}")
    if isinstance(static_ndrange, StaticSize):
        if ndrange is not None and ndrange != static_ndrange.S:
            raise ErrorException(f"Static NDRange ({static_ndrange}) and launch NDRange ({ndrange}) differ")
        ndrange = static_ndrange.S
    if isinstance(static_workgroupsize, StaticSize):
        if workgroupsize is not None and workgroupsize != static_workgroupsize.S:
            raise ErrorException(
                f"Static WorkgroupSize ({static_workgroupsize}) and launch WorkgroupSize {workgroupsize} differ")
        workgroupsize = static_workgroupsize.S
    assert workgroupsize is not None and ndrange is not None
    blocks, wgs, dynamic = nd_partition(ndrange, workgroupsize)
    nd = NDRange(len(ndrange), blocks, wgs, isinstance(static_ndrange, StaticSize),
                 isinstance(static_workgroupsize, StaticSize))
    return nd, dynamic


def threads_to_workgroupsize(threads: int, ndrange: Sequence[int]) -> Tuple[int, ...]:
    """reference src/pocl/backend.jl:108-115"""
    total = 1
    out = []
    for n in ndrange:
        x = min(threads // total, int(n))
        total *= x
        out.append(x)
    return tuple(out)
