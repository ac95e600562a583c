"""NVTX ranges around the phases of a step (tracing).

`GPB_NVTX=1` in the environment turns them on; without it `annotate` returns the function it was given, so the
default path carries no wrapper at all.  The ranges nest as  objective_function > layer phase
(pre_tail / fwd_* / bwd_* / tail_*)  and show up in any NVTX-aware tool (Nsight Systems / Compute:
`ncu --nvtx --nvtx-include "geepee/bwd_mm/"` restricts a capture to one phase)."""
import functools
import os

ENABLED = os.environ.get('GPB_NVTX', '0') not in ('', '0')


def annotate(name):
    def deco(fn):
        if not ENABLED:
            return fn
        import torch

        @functools.wraps(fn)
        def wrapped(*args, **kwargs):
            torch.cuda.nvtx.range_push('geepee/' + name)
            try:
                return fn(*args, **kwargs)
            finally:
                torch.cuda.nvtx.range_pop()
        return wrapped
    return deco
