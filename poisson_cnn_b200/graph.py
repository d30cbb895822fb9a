"""CUDA-graph replay of a forward pass for latency-bound (small-batch) use.

A forward pass of Poisson_CNN_Legacy is ~165 kernel launches; at batch 1-8 the host needs longer to enqueue them
than the GPU needs to run them.  GraphedCall captures the launches of one (batch, grid) shape once and replays
them with a single cudaGraphLaunch.  The reference reaches for @tf.function for the same reason
(poisson_CNN/losses/physics_informed_loss.py:34, dataset/utils/image_resize.py:4); here it is a plain CUDA graph,
no tracing compiler.

What makes the capture valid:
  * every libpcnn entry point only enqueues work on the caller's stream (no allocation, no synchronisation);
  * torch allocations inside the capture come from the graph's private memory pool;
  * BLK8 activation buffers carry state between passes (their materialised halo ring): the graph gets a PRIVATE
    buffer pool whose free lists are put in address order before every pass (so a buffer always serves the same
    tensor), and the capture starts only after two consecutive warm-up passes ended in the same pool state (same
    buffers, same halo states), so every replay starts from exactly what the captured pass started from.
"""
import torch

from . import ops


class GraphedCall:
    def __init__(self, fn, example_inputs, max_warmup=6):
        if not all(t.is_cuda for t in example_inputs):
            raise ValueError("capture needs CUDA example inputs")
        self.device = example_inputs[0].device
        self._in = [t.detach().clone().contiguous() for t in example_inputs]
        self._pool = {}
        with torch.cuda.device(self.device):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            steady, prev = False, None
            with torch.cuda.stream(side), ops.blk8_pool_scope(self._pool):
                for _ in range(max_warmup):
                    fn(self._in)
                    ops.blk8_pool_canonicalize(self._pool)       # fixed buffer -> tensor assignment from pass to pass
                    snap = ops.blk8_pool_snapshot(self._pool)
                    if prev is not None and snap == prev:
                        steady = True
                        break
                    prev = snap
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(self.device)
            if not steady:
                raise RuntimeError("the activation-buffer pool did not reach a periodic state in %d passes; not capturing" % max_warmup)
            self._graph = torch.cuda.CUDAGraph()
            with ops.blk8_pool_scope(self._pool), torch.cuda.graph(self._graph):
                self._out = fn(self._in)
            ops.blk8_pool_canonicalize(self._pool)
            if ops.blk8_pool_snapshot(self._pool) != prev:
                raise RuntimeError("the captured pass left the buffer pool in a different state than the warm-up passes")

    def __call__(self, inputs, out=None):
        """Copies `inputs` into the graph's static input buffers, replays, and returns the result: `out` if given
        (copied on the current stream), else the graph's own output tensor, which the next call overwrites."""
        if len(inputs) != len(self._in):
            raise ValueError("expected %d inputs" % len(self._in))
        for dst, src in zip(self._in, inputs):
            if tuple(src.shape) != tuple(dst.shape):
                raise ValueError("graph captured for input shape %s, got %s" % (tuple(dst.shape), tuple(src.shape)))
            dst.copy_(src, non_blocking=True)
        self._graph.replay()
        if out is not None:
            out.copy_(self._out, non_blocking=True)
            return out
        return self._out
