"""Data-parallel gradient exchange: the one collective of the path.

The reference wraps the model in torch DDP (train.py:304) whose only data-path collective is the
bucketed gradient all-reduce under loss.backward() (train.py:528).  Here every parameter's .grad is a
view into ONE flat fp32 buffer, so the exchange is a single `all_reduce(SUM)` over 46.8 MB (R18) on
NCCL/NVLink followed by an in-place 1/world scale — no per-bucket Python hooks, no copies in or out.
Kurtosis / KD-layer gradients depend only on the replicated weights, so they are identical on every
rank and the averaging leaves them unchanged (same as DDP).  BatchNorm statistics stay per-rank (the
reference does not use SyncBN)."""
import torch
import torch.distributed as dist


class GradAllReduce:
    def __init__(self, model, process_group=None, broadcast_params=True):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.params = [p for p in model.parameters() if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            n = p.numel()
            # autograd accumulates in place into the view; keep the parameter's own strides
            # (channels_last conv weights) so fused optimizers see matching layouts
            p.grad = torch.as_strided(self.flat, p.size(), p.stride(), storage_offset=off)
            off += n
        if broadcast_params and self.world > 1:
            # DDP broadcasts rank-0 parameters and buffers at construction (train.py:304)
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t.data, src=0, group=self.group)

    def zero(self):
        self.flat.zero_()

    def __call__(self):
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.mul_(1.0 / self.world)


class FlatGradOptimizerShim:
    """optimizer.zero_grad() replacement that keeps .grad views alive (zeroes the flat buffer)."""

    def __init__(self, optimizer, reducer: GradAllReduce):
        self.optimizer, self.reducer = optimizer, reducer
        self.param_groups = optimizer.param_groups

    def zero_grad(self, set_to_none=False):
        self.reducer.zero()

    def step(self):
        return self.optimizer.step()

    def state_dict(self):
        return self.optimizer.state_dict()
