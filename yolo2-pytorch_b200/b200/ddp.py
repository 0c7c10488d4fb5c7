"""Data-parallel gradient exchange for the training step: one process per GPU, images sharded by rank, parameter
gradients summed over ranks with NCCL all-reduce (NVLink 5 / NVSwitch) and divided by the world size.

Replaces the reference's single-process nn.DataParallel (train.py:65-71: scatter / replicate / gather on GPU 0 and
`reduce_add_coalesced` of the gradients).  The backward chain (b200.train_engine) reports every parameter gradient
as soon as its kernels are enqueued (`on_grad`), deepest layers first -- layers3.0 (47 MB) and layers2.6/2.7
(38 MB each) come out at the very start -- so each bucket's all-reduce is launched on a dedicated communication
stream while the remaining backward kernels are still running.  Semantics match torch DDP: average of per-rank
gradients; BatchNorm statistics stay per rank (the per-replica statistics of DataParallel).

Works with any torch.distributed backend (NCCL on GPUs; gloo on CPU tensors for the host-logic tests).
"""
import torch
import torch.distributed as dist


class GradientAllReducer(object):
    def __init__(self, bucket_bytes=32 << 20, process_group=None):
        self.bucket_bytes = int(bucket_bytes)
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.comm_stream = None
        self._pending = []      # (flat buffer, [(name, tensor, numel)], work handle, ready event)
        self._bucket = []
        self._bucket_size = 0
        self.bytes_reduced = 0
        self.steps = 0          # finish() calls executed by Python (CUDA-graph replays do not come through here)

    # ---- called by the backward chain -----------------------------------------------------------------
    def on_grad(self, name, grad):
        """Register a freshly produced gradient (enqueued on the current stream)."""
        if self.world == 1:
            return
        nbytes = grad.numel() * grad.element_size()
        self._bucket.append((name, grad))
        self._bucket_size += nbytes
        if self._bucket_size >= self.bucket_bytes:
            self._flush()

    def _flush(self):
        if not self._bucket:
            return
        items = self._bucket
        self._bucket, self._bucket_size = [], 0
        cuda = items[0][1].is_cuda
        if cuda:
            if self.comm_stream is None:
                self.comm_stream = torch.cuda.Stream(device=items[0][1].device)
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(items[0][1].device))
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ready)
                flat = torch.cat([g.reshape(-1) for _, g in items]) if len(items) > 1 else items[0][1].reshape(-1)
                work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            flat = torch.cat([g.reshape(-1) for _, g in items]) if len(items) > 1 else items[0][1].reshape(-1)
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.bytes_reduced += flat.numel() * flat.element_size()
        self._pending.append((flat, items, work, cuda))

    def finish(self):
        """Flush the last bucket, wait for every all-reduce and write the averaged gradients back in place."""
        if self.world == 1:
            return
        self.steps += 1
        self._flush()
        scale = 1.0 / self.world
        for flat, items, work, cuda in self._pending:
            work.wait()
            if cuda:
                with torch.cuda.stream(self.comm_stream):
                    self._scatter_back(flat, items, scale)
            else:
                self._scatter_back(flat, items, scale)
        if self._pending and self._pending[0][3]:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self._pending = []

    @staticmethod
    def _scatter_back(flat, items, scale):
        off = 0
        for _, g in items:
            n = g.numel()
            if flat.data_ptr() != g.data_ptr():
                g.copy_(flat[off:off + n].view_as(g))
            g.mul_(scale)
            off += n


def global_mean_factor(local_count, process_group=None):
    """Weight that turns a per-rank MEAN over `local_count` items into this rank's share of the mean over the items of
    ALL ranks: local_count / sum_over_ranks(local_count) (one scalar all-reduce).

    Used for the class term of the region loss (`train/cross_entropy = 1`): the reference computes the loss on the
    batch gathered from all replicas (train.py:344-347), where that term is a mean over the positives of the whole
    batch and is then divided by cnt = B_global * cells * A (model/__init__.py:162-166).  With images sharded over
    ranks, (sum_r CE_r / N_r) * f_r / cnt_local averaged over ranks equals (sum CE / N_total) / cnt_global exactly
    when f_r = N_r / N_total, because cnt_global = world * cnt_local.  The sum-type terms need no correction.
    `local_count`: 0-dim tensor (any numeric dtype; CUDA with NCCL, CPU with gloo).  Returns a float32 0-dim tensor."""
    n = local_count.detach().to(torch.float32).reshape(())
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(process_group) == 1:
        return torch.ones_like(n)
    total = n.clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=process_group)
    return n / total.clamp(min=1.0)


def shard_range(total, rank, world):
    """Contiguous [start, end) of `total` images owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)
