"""Side CUDA streams for the replicated O(Dout M^3) tails.

The per-row kernels of one objective call run on the caller's (current) stream.  The tails of
the sparse-GP layers -- q(u)/cavity algebra before the forward, chain rules after the backward,
and (world size > 1) the all-reduce of that layer's statistics -- are small, launch-bound kernels
that depend on nothing but the parameters resp. that layer's statistics.  Each layer gets its
own side stream:

    fork(i)   side stream i waits for everything queued so far on the main stream
    on(i)     context manager: work issued inside goes to side stream i
    join(i)   the main stream waits for everything queued so far on side stream i

so layer i+1's pre-tail overlaps layer i's forward kernels, and layer i's all-reduce + post-tail
overlap the backward kernels of the layers below it.  Memory discipline (torch caching
allocator): every step starts with fork() of all side streams and ends with join() of all of
them, so a block freed by one stream's tensors is never handed out while the other stream still
reads it.  On the CPU (emulator tests) everything runs inline; GPB_NO_SIDE_STREAMS=1 in the environment does the
same on the GPU (diagnostic: it is how the TMEM co-residency stall of profiles/r2_mm_pairs_tc.txt was found).
"""
import contextlib
import os

import torch


class TailStreams(object):
    def __init__(self, device, n):
        self.cuda = (device.type == 'cuda') and not os.environ.get('GPB_NO_SIDE_STREAMS')
        self.streams = [torch.cuda.Stream(device) for _ in range(n)] if self.cuda else [None] * n

    def mark(self):
        """Event on the current stream (e.g. 'parameters uploaded')."""
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            return ev
        return None

    def fork(self, i, after=None):
        """Side stream i waits for the main stream -- or only for the event `after` (used at the
        start of a step, when everything older than the upload is already complete)."""
        if self.cuda:
            if after is not None:
                self.streams[i].wait_event(after)
            else:
                self.streams[i].wait_stream(torch.cuda.current_stream())

    def on(self, i):
        if self.cuda:
            return torch.cuda.stream(self.streams[i])
        return contextlib.nullcontext()

    def keep(self, i, tensors):
        """Tensors allocated on the main stream that side stream i is about to read: tell the caching
        allocator, so that a free on the main stream cannot hand their memory out while stream i still
        reads them."""
        if self.cuda:
            for t in tensors:
                if torch.is_tensor(t) and t.is_cuda:
                    t.record_stream(self.streams[i])

    def join(self, i):
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.streams[i])

    def join_all(self):
        for i in range(len(self.streams)):
            self.join(i)
