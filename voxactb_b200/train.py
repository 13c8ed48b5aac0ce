"""Training-tail building blocks behind the C ABI: the agent's per-sample
cross-entropy losses with their logit gradients, fused multi-tensor LAMB / Adam steps with the reference's
semantics, and the DDP-equivalent gradient all-reduce (NCCL through torch.distributed: sum, then / world).
The backward pass of the Q-network is NOT built yet, so these are not wired into an ``update()``."""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


def cross_entropy(logits, labels, grad_scale=None):
    """Per-sample CE of ``logits`` [B, N] (any row-strided 2-D CUDA fp32 view) against int label indices --
    ``nn.CrossEntropyLoss(reduction='none')(pred, onehot.argmax(-1))`` of reference agent:391-392.
    Returns (loss [B], grad [B, N] or None); grad = grad_scale * (softmax - onehot)."""
    if logits.dim() != 2 or logits.stride(1) != 1 or logits.dtype != torch.float32 or not logits.is_cuda:
        raise ValueError('cross_entropy: logits must be a CUDA fp32 [B, N] view with unit column stride')
    B, N = logits.shape
    if not labels.is_cuda and labels.numel() and (int(labels.min()) < 0 or int(labels.max()) >= N):
        raise IndexError('cross_entropy: Target %d is out of bounds for %d classes'
                         % (int(labels.max()) if int(labels.max()) >= N else int(labels.min()), N))
    lab = labels.to(device=logits.device, dtype=torch.int32).contiguous()
    if lab.shape != (B,):
        raise ValueError('cross_entropy: labels must be [B]')
    L = _lib.lib()
    ws_bytes = L.vxb_ce_loss_workspace_bytes(B, N)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=logits.device)
    loss = torch.empty(B, dtype=torch.float32, device=logits.device)
    grad = torch.empty(B, N, dtype=torch.float32, device=logits.device) if grad_scale is not None else None
    rc = L.vxb_ce_loss_f32(ctypes.c_void_p(logits.data_ptr()), logits.stride(0), _lib.ptr(lab), B, N,
                           float(grad_scale or 0.0), _lib.ptr(loss), _lib.ptr(grad), N, _lib.ptr(ws), ws_bytes,
                           _lib.stream())
    _lib.check(rc, 'vxb_ce_loss_f32')
    return loss, grad


def peract_losses(q_trans, q_rot_grip, q_collision, action_trans, action_rot_grip, action_ignore_collisions,
                  num_rotation_classes=72, weights=(1.0, 1.0, 1.0, 1.0), with_grad=False):
    """The loss of QAttentionPerActBCAgent.update (reference agent:517-578) from label INDICES:
    total = mean_b( w_t CE(V^3) + w_r (CE_x + CE_y + CE_z) + w_g CE(grip) + w_c CE(collision) ).
    Returns (total, dict of per-sample terms, dict of logit gradients or None)."""
    B = q_trans.shape[0]
    R = num_rotation_classes
    gs = (lambda w: w / B) if with_grad else (lambda w: None)
    flat = q_trans.reshape(B, -1)
    V = q_trans.shape[-1]
    t_idx = (action_trans[:, 0].long() * V + action_trans[:, 1].long()) * V + action_trans[:, 2].long()
    lt, gt = cross_entropy(flat, t_idx, gs(weights[0]))
    terms, grads = {'trans': lt}, {}
    rot = torch.zeros_like(lt)
    g_rg = torch.zeros_like(q_rot_grip) if with_grad else None
    for a in range(3):
        l, g = cross_entropy(q_rot_grip[:, a * R:(a + 1) * R], action_rot_grip[:, a], gs(weights[1]))
        rot = rot + l
        if with_grad:
            g_rg[:, a * R:(a + 1) * R] = g
    lg, gg = cross_entropy(q_rot_grip[:, 3 * R:], action_rot_grip[:, 3], gs(weights[2]))
    lc, gc = cross_entropy(q_collision, action_ignore_collisions.reshape(B), gs(weights[3]))
    terms.update(rot=rot, grip=lg, collision=lc)
    total = (lt * weights[0] + rot * weights[1] + lg * weights[2] + lc * weights[3]).mean()
    if with_grad:
        g_rg[:, 3 * R:] = gg
        grads = {'q_trans': gt.reshape(q_trans.shape), 'q_rot_grip': g_rg, 'q_collision': gc}
    return total, terms, (grads if with_grad else None)


class _FusedOptimizer(torch.optim.Optimizer):
    def _tables(self, group):
        ps = [p for p in group['params'] if p.grad is not None]
        for p in ps:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()):
                raise RuntimeError('fused optimizers need contiguous CUDA fp32 parameters and gradients')
            st = self.state[p]
            if not st:
                st['step'] = 0
                st['exp_avg'] = torch.zeros_like(p)
                st['exp_avg_sq'] = torch.zeros_like(p)
            st['step'] = int(st['step']) + 1
        n = len(ps)
        arr = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])
        sizes = (ctypes.c_longlong * n)(*[p.numel() for p in ps])
        return ps, n, arr(ps), arr([p.grad for p in ps]), arr([self.state[p]['exp_avg'] for p in ps]), \
            arr([self.state[p]['exp_avg_sq'] for p in ps]), sizes

    def _ws(self, n, sizes, device):
        nbytes = _lib.lib().vxb_optimizer_workspace_bytes(n, sizes)
        return torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes


class Lamb(_FusedOptimizer):
    """Fused multi-tensor LAMB with the semantics of reference peract/helpers/optim/lamb.py:27-122."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0, adam=False):
        if adam:
            raise NotImplementedError('adam=True (trust ratio forced to 1) is not built; use voxactb_b200.train.Adam')
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            ps, n, w, g, m, v, sizes = self._tables(group)
            if not n:
                continue
            ws, nbytes = self._ws(n, sizes, ps[0].device)
            rc = _lib.lib().vxb_lamb_step_f32(n, w, g, m, v, sizes, group['lr'], group['betas'][0], group['betas'][1],
                                              group['eps'], group['weight_decay'], _lib.ptr(ws), nbytes, _lib.stream())
            _lib.check(rc, 'vxb_lamb_step_f32')
            torch._C._increment_version(ps)   # the kernel wrote through raw pointers: invalidate prepared-weight caches
        return loss


class Adam(_FusedOptimizer):
    """Fused multi-tensor torch.optim.Adam (L2 weight decay), as the agent builds it (reference agent:263-268)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            ps, n, w, g, m, v, sizes = self._tables(group)
            if not n:
                continue
            steps = {int(self.state[p]['step']) for p in ps}   # a torch.optim.Adam checkpoint stores tensors
            if len(steps) != 1:
                raise RuntimeError('fused Adam needs all parameters of a group at the same step')
            ws, nbytes = self._ws(n, sizes, ps[0].device)
            rc = _lib.lib().vxb_adam_step_f32(n, w, g, m, v, sizes, steps.pop(), group['lr'], group['betas'][0],
                                              group['betas'][1], group['eps'], group['weight_decay'], _lib.ptr(ws),
                                              nbytes, _lib.stream())
            _lib.check(rc, 'vxb_adam_step_f32')
            torch._C._increment_version(ps)
        return loss


def allreduce_gradients(params, bucket_bytes=25 << 20):
    """DDP-equivalent gradient averaging (reference agent:50-54 wraps the Q-network in DDP over gloo): bucketed
    all-reduce (sum) of the gradients over the default process group -- NCCL over NVLink on GPUs -- then / world."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    world = dist.get_world_size()
    grads = [p.grad for p in params if p.grad is not None]
    bucket, size = [], 0
    def flush():
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        o = 0
        for g in bucket:
            g.copy_(flat[o:o + g.numel()].view_as(g))
            o += g.numel()
    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
            bucket, size = [], 0
    flush()
