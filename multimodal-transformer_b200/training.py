"""Training-step plumbing around the drop-in modules: fused MSE loss (+ its gradient), one-launch-per-arena Adam with
L2 weight decay (torch.optim.Adam semantics, MFT/train.py:557), and the data-parallel gradient all-reduce over the
flat gradient buffers (one process per GPU, torch.distributed; the reference has no distributed code at all).

    opt = FlatAdam(model, lr=1e-4, weight_decay=1e-4)
    pred = model(inputs, mask, lengths)
    loss = train_step_loss(pred, target, global_sum_lengths)      # backward() included
    opt.step()                                                    # all-reduces first when torch.distributed is up
"""
import torch
import torch.distributed as dist

from . import functional as K


def _arenas_of(model):
    seen, out = set(), []
    for m in model.modules():
        for getter in ('arena',):
            if hasattr(m, getter) and callable(getattr(m, getter)):
                a = getattr(m, getter)()
                if id(a) not in seen:
                    seen.add(id(a)); out.append(a)
        a = getattr(m, '_dec_arena', None)
        if a is not None and id(a) not in seen:
            seen.add(id(a)); out.append(a)
    return out


def train_step_loss(pred, target, norm):
    """loss = sum((pred - target)^2) / norm (MFT/train.py:135-139) computed and differentiated by one kernel, then
    pred.backward(dloss/dpred).  `norm` is the GLOBAL sum of lengths under data parallelism.  Returns loss [1]."""
    loss, dpred = K.mse_loss_sum_normalised(pred.detach(), target, norm)
    pred.backward(dpred.view_as(pred))
    return loss


class FlatAdam:
    """Adam over the models' flat parameter arenas: one mt_adam_step launch per arena, gradients read in place from
    the flat buffers backward produced.  Parameters outside any arena (embed / fusion Linears) are gathered into one
    more arena the first time they are seen with a gradient.  Parameters that never receive a gradient (the orphan
    attn{mod}/ff{mod} templates) are left untouched, like torch.optim.Adam does."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, process_group=None):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.group = process_group
        self.step_count = 0
        self.state = {}          # id(arena) -> (m, v)
        self.misc = None
        self.param_groups = [dict(lr=lr)]      # ReduceLROnPlateau-style schedulers poke this

    def _all_arenas(self):
        arenas = _arenas_of(self.model)
        in_arena = {id(p) for a in arenas for p in a.params}
        if self.misc is None:
            rest = [p for p in self.model.parameters() if id(p) not in in_arena and p.grad is not None]
            if rest:
                self.misc = K.Arena(rest)
                self.misc.bind()
        return arenas + ([self.misc] if self.misc is not None else [])

    def _flat_grad(self, arena):
        g = arena.flat_grad()
        if g is None:
            if any(p.grad is None for p in arena.params):
                return None
            g = torch.cat([p.grad.reshape(-1) for p in arena.params])
        return g

    def all_reduce_grads(self):
        """SUM the gradients over ranks (the loss is already normalised by the global sum of lengths)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        for a in self._all_arenas():
            g = self._flat_grad(a)
            if g is None:
                continue
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
            if a.flat_grad() is None:          # gathered copy: scatter back
                for p, v in zip(a.params, a.grad_views(g)):
                    p.grad.copy_(v)

    @torch.no_grad()
    def step(self):
        self.all_reduce_grads()
        self.step_count += 1
        lr = self.param_groups[0]['lr']
        for a in self._all_arenas():
            g = self._flat_grad(a)
            if g is None:
                continue
            flat = a.bind()
            st = self.state.get(id(a))
            if st is None:
                st = (torch.zeros_like(flat), torch.zeros_like(flat))
                self.state[id(a)] = st
            K.adam_step_flat(flat, g, st[0], st[1], self.step_count, lr, self.betas, self.eps, self.weight_decay)
            a._lp_key = None             # the kernel wrote the arena behind autograd's back: refresh the bf16 shadow

    def zero_grad(self):
        for p in self.model.parameters():
            p.grad = None
