"""Data-parallel gradient exchange: the one collective of the path.

The reference wraps the model in torch DDP (train.py:304) whose only data-path collective is the
bucketed gradient all-reduce under loss.backward() (train.py:528).  Here every parameter's .grad is a
view into ONE flat fp32 buffer (46.8 MB for R18), all-reduced (SUM) in a few contiguous buckets over
NCCL/NVLink while the backward is still running, then scaled by 1/world in place — no copies in or out.
Kurtosis / KD-layer gradients depend only on the replicated weights, so they are identical on every
rank and the averaging leaves them unchanged (same as DDP).  BatchNorm statistics stay per-rank (the
reference does not use SyncBN)."""
import torch
import torch.distributed as dist


class GradAllReduce:
    """Flat-buffer gradient all-reduce, overlapped with the backward.

    Parameters are laid out in the flat buffer in REVERSE registration order (the order their gradients
    become final during backward: classifier and layer4 — 3/4 of ResNet-18's parameters — first) and cut
    into `n_buckets` contiguous buckets.  A post-accumulate hook on every parameter counts its bucket
    down; the bucket's `all_reduce(SUM, async)` is issued the moment its last gradient is written, so the
    NCCL transfer of the big late-layer buckets runs under the backward of the early layers.  `__call__`
    (after backward) issues whatever is left, waits, and applies 1/world.

    A bucket is issued one event AFTER its last gradient was accounted for (the next hook / side-stream launch, or
    `__call__`): autograd runs a node's AccumulateGrad hooks right after the node, so by then every accumulation
    into the bucket is enqueued.  That matters for the side-stream weight gradients (functional.wgrad_side): there
    a conv weight's gradient is written by a GEMM on the second stream into `wflat` (a second flat buffer of the
    same layout) and the parameter is accounted for when that GEMM is launched; a kurtosis term's AccumulateGrad
    on the same weight follows immediately on the main stream.  The bucket is then issued FROM the side stream
    after it has waited for the main stream: flat[bucket] += wflat[bucket], all_reduce(flat[bucket]).  A gradient
    that arrives after its bucket was issued raises."""

    def __init__(self, model, process_group=None, broadcast_params=True, n_buckets=2, overlap=True, scale=True):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.params = [p for p in model.parameters() if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.overlap = bool(overlap) and self.world > 1
        self.scale = bool(scale)      # False: the optimizer applies 1/world (optim.FusedAdam/FusedSGD grad_scale)
        order = list(reversed(self.params))
        self._view = {}                        # id(p) -> the parameter's view into the flat buffer
        target = max(1, (total + n_buckets - 1) // max(1, n_buckets))
        self.buckets = []                      # [start, end, n_params]
        self._bucket_of = {}
        off, b_start, b_count = 0, 0, 0
        for p in order:
            n = p.numel()
            # autograd accumulates in place into the view; keep the parameter's own strides
            # (channels_last conv weights) so fused optimizers see matching layouts
            p.grad = torch.as_strided(self.flat, p.size(), p.stride(), storage_offset=off)
            self._view[id(p)] = p.grad
            self._bucket_of[id(p)] = len(self.buckets)
            off += n
            b_count += 1
            if off - b_start >= target:
                self.buckets.append([b_start, off, b_count])
                b_start, b_count = off, 0
        if b_count:
            self.buckets.append([b_start, off, b_count])
        self._pending = [b[2] for b in self.buckets]
        self._works = [None] * len(self.buckets)
        self._done = set()                     # id(p) accounted for in this step
        self._armed = []                       # complete buckets, issued at the next event
        self.wflat = None                      # side-stream weight gradients (same layout as flat)
        self._wview = {}
        self._side_stream = None
        self._side_now = set()                 # id(p) whose wgrad ran on the side stream in this step
        self._side_ever = {}                   # id(p) -> p: wflat entries that hold (possibly stale) values
        if self.overlap:
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._hook)
        if broadcast_params and self.world > 1:
            # DDP broadcasts rank-0 parameters and buffers at construction (train.py:304)
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t.data, src=0, group=self.group)

    def _launch(self, b):
        a, e, _ = self.buckets[b]
        if self._side_now or self._side_ever:
            side = self._side_stream
            side.wait_stream(torch.cuda.current_stream())     # every accumulation enqueued on the main stream so far
            with torch.cuda.stream(side):
                for i, q in self._side_ever.items():          # entries not rewritten in this step are stale
                    if i not in self._side_now and self._bucket_of[i] == b:
                        self._wview[i].zero_()
                self.flat[a:e].add_(self.wflat[a:e])
                self._works[b] = dist.all_reduce(self.flat[a:e], op=dist.ReduceOp.SUM, group=self.group,
                                                 async_op=True)
            return
        self._works[b] = dist.all_reduce(self.flat[a:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _flush_armed(self):
        for b in self._armed:
            if self._works[b] is None:
                self._launch(b)
        self._armed = []

    def _check_open(self, p):
        if self._works[self._bucket_of[id(p)]] is not None:
            raise RuntimeError("GradAllReduce: a gradient was written after its bucket's all-reduce was issued "
                               "(unexpected backward order; set BDBNN_WGRAD_SIDE=0 or overlap=False)")

    def _account(self, p):
        b = self._bucket_of[id(p)]
        if id(p) in self._done:       # second event of a weight with a side-stream wgrad AND an autograd term
            return
        self._done.add(id(p))
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._armed.append(b)

    # ---- functional.wgrad_side sink ----------------------------------------------------------------------
    def side_target(self, p, stream):
        """Buffer the side-stream wgrad of `p` writes (its view into `wflat`), or None: keep it on autograd's path."""
        if not self.overlap or id(p) not in self._bucket_of or not p.is_contiguous():
            return None
        if self.wflat is None:
            self.wflat = torch.zeros_like(self.flat)
            off = 0
            for q in reversed(self.params):
                self._wview[id(q)] = torch.as_strided(self.wflat, q.size(), q.stride(), storage_offset=off)
                off += q.numel()
        self._side_stream = stream
        self._check_open(p)
        self._flush_armed()
        self._side_now.add(id(p))
        self._side_ever[id(p)] = p
        self._account(p)          # the launch follows immediately; the bucket is issued at the NEXT event
        return self._wview[id(p)]

    def _rebind(self, p):
        """Fail-safe: `optimizer.zero_grad()` (torch's default set_to_none=True) or user code dropped /
        replaced the .grad view, so autograd wrote this step's gradient into a fresh tensor.  Copy it into
        the flat buffer and re-point .grad at the view — never all-reduce a stale buffer."""
        v = self._view[id(p)]
        g = p.grad
        if g is None:
            v.zero_()
        elif g.data_ptr() != v.data_ptr() or g.stride() != v.stride():
            v.copy_(g)
        else:
            return
        p.grad = v

    def _hook(self, p):
        self._check_open(p)           # (this gradient's accumulation is already enqueued: a flush below covers it)
        self._flush_armed()
        self._rebind(p)
        self._account(p)

    def zero(self):
        self.flat.zero_()
        for p in self.params:                  # re-point views a set_to_none zero_grad() dropped
            v = self._view[id(p)]
            if p.grad is not v:
                p.grad = v
        self._reset()

    def _reset(self):
        self._pending = [b[2] for b in self.buckets]
        self._works = [None] * len(self.buckets)
        self._done, self._armed, self._side_now = set(), [], set()

    def __call__(self):
        if not self.overlap:
            for p in self.params:              # no hooks registered: verify the views here
                self._rebind(p)
        if self.world > 1:
            self._armed = []
            for b in range(len(self.buckets)):
                if self._works[b] is None:
                    self._launch(b)
            for w in self._works:
                w.wait()
            if self.scale:
                self.flat.mul_(1.0 / self.world)
        self._reset()


class FlatGradOptimizerShim:
    """optimizer.zero_grad() replacement that keeps .grad views alive (zeroes the flat buffer) — the fast
    path.  Without it GradAllReduce still gives correct results (see `_rebind`) at the cost of one copy
    per parameter per step; TrainStep calls `grad_sync.zero()` itself when it is handed a GradAllReduce."""

    def __init__(self, optimizer, reducer: GradAllReduce):
        self.optimizer, self.reducer = optimizer, reducer
        self.param_groups = optimizer.param_groups

    def zero_grad(self, set_to_none=False):
        self.reducer.zero()

    def step(self):
        return self.optimizer.step()

    def state_dict(self):
        return self.optimizer.state_dict()
