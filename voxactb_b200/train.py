"""Training step behind the C ABI: the agent's per-sample cross-entropy losses with their
logit gradients, the Q-network backward (vxb_qnet_backward_f32, reached either through torch autograd -- the encoder's
training forward is one autograd node, so the reference's own ``agent.update`` works unchanged -- or directly by
``PerActTrainer.update`` below), the DDP-equivalent gradient all-reduce over NCCL and fused multi-tensor LAMB / Adam
steps with the reference's semantics."""
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


class PerActTrainer:
    """Host-side mirror of the training half of QAttentionPerActBCAgent (reference qattention_peract_bc_agent.py:418-641)
    for one QFunction: ``update`` = forward (training mode) -> the four (five with the arm head) cross-entropy losses on
    label indices (:517-578) -> Q-network backward -> gradient all-reduce across ranks (DDP semantics, :50-54) ->
    optimizer step (LAMB by default, lamb.py:60-122).  No autograd graph is built: the loss gradients w.r.t. the logits
    come from the fused CE kernel and go straight into vxb_qnet_backward_f32."""

    def __init__(self, q, lr=5e-4, weight_decay=1e-6, optimizer='lamb', loss_weights=(1.0, 1.0, 1.0, 1.0), arm_loss_weight=1.0,
                 num_rotation_classes=72):
        self.q = q
        self.enc = q._qnet
        self.params = [p for p in self.enc._param_table()[2] if p is not None]
        if optimizer == 'lamb':
            self.optimizer = Lamb(self.params, lr=lr, weight_decay=weight_decay, betas=(0.9, 0.999))
        elif optimizer == 'adam':
            self.optimizer = Adam(self.params, lr=lr, weight_decay=weight_decay)
        else:
            raise Exception('Unknown optimizer type')
        self.loss_weights = tuple(loss_weights)
        self.arm_loss_weight = arm_loss_weight
        self.R = num_rotation_classes
        self.reducer = GradientReducer(self.params)   # flat fp32 gradient arena; the backward writes straight into its views
        self.last = {}
        self.profile = False          # True: CUDA-event stage times of the next update in last_stage_ms
        self.last_stage_ms = {}

    def losses_and_logit_grads(self, q_trans, q_rot_grip, q_collision, arm_out, labels):
        B = q_trans.shape[0]
        total, terms, grads = peract_losses(q_trans, q_rot_grip, q_collision, labels['trans'], labels['rot_grip'],
                                            labels['collision'], self.R, self.loss_weights, with_grad=True)
        g_arm = None
        if arm_out is not None and 'arm' in labels:
            la, g_arm = cross_entropy(arm_out, labels['arm'].reshape(B), self.arm_loss_weight / B)
            terms['arm'] = la
            total = total + (la * self.arm_loss_weight).mean()
        return total, terms, grads, g_arm

    def update(self, rgb_pcd, proprio, pcd, lang_goal_emb, lang_token_embs, bounds, labels):
        """One training step on a replay batch already on the device.  labels: dict(trans [B,3], rot_grip [B,4],
        collision [B,1][, arm [B,1]]) of integer indices (agent:419-423).  Returns dict(total_loss, terms)."""
        enc = self.enc
        if not enc.training:
            raise RuntimeError('PerActTrainer.update needs the encoder in training mode (QFunction.train())')
        evs = []

        def mark(name):
            if self.profile:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                evs.append((name, e))

        with torch.no_grad():
            mark('start')
            b = rgb_pcd[0][0].shape[0]
            pcd_flat = torch.cat([p.permute(0, 2, 3, 1).reshape(b, -1, 3) for p in pcd], 1)
            feats = torch.cat([rp[0].permute(0, 2, 3, 1).reshape(b, -1, rp[0].shape[1]) for rp in rgb_pcd], 1)
            grid = self.q._voxelizer.coords_to_bounding_voxel_grid(pcd_flat, coord_features=feats, coord_bounds=bounds)
            inputs = (_lib.f32(grid), _lib.f32(proprio), _lib.f32(lang_token_embs))
            mark('voxelize')
            outs, gen = enc._forward_train(*inputs)
            mark('forward')
            arm = outs[3] if enc.arm_pred_loss else None
            total, terms, g, g_arm = self.losses_and_logit_grads(outs[0], outs[1], outs[2], arm, labels)
            mark('losses')
            enc._backward_train(gen, inputs, (g['q_trans'], g['q_rot_grip'], g['q_collision'], g_arm), out=self.reducer.views)
            for p, v in zip(self.params, self.reducer.views):
                p.grad = v
            mark('backward')
            self.reducer.allreduce()
            mark('allreduce')
            self.optimizer.step()
            mark('optimizer')
        if self.profile:
            torch.cuda.synchronize()
            self.last_stage_ms = {evs[i][0]: evs[i - 1][1].elapsed_time(evs[i][1]) for i in range(1, len(evs))}
        self.last = {'total_loss': total, 'terms': terms, 'voxel_grid': grid}
        return self.last


class GradientReducer:
    """DDP-equivalent gradient averaging (reference agent:50-54) on a flat, pre-allocated gradient arena: the per-parameter
    `.grad` tensors are views into ONE contiguous fp32 buffer (133 MB for the PerAct Q-network), all-reduced in
    `bucket_bytes` slices over NCCL (NVLink / NVSwitch) on a side stream and scaled by 1 / world in the same pass."""

    def __init__(self, params, bucket_bytes=32 << 20):
        self.params = list(params)
        self.comm = None
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=self.params[0].device)
        self.views = []
        o = 0
        for p in self.params:
            self.views.append(self.flat[o:o + p.numel()].view_as(p))
            o += p.numel()
        self.bucket_elems = max(1, bucket_bytes // 4)

    def allreduce(self):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        world = dist.get_world_size()
        for p, v in zip(self.params, self.views):
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
                p.grad = v
        n = self.flat.numel()
        if self.flat.is_cuda:
            # the library's own NCCL communicator (vxb_allreduce_grads): one bucketed all-reduce group + 1/world scaling on the
            # caller's stream.  The 128-byte unique id travels through the existing torch.distributed group once.
            L = _lib.lib()
            if self.comm is None:
                idbuf = ctypes.create_string_buffer(128)
                if dist.get_rank() == 0:
                    _lib.check(L.vxb_nccl_unique_id(idbuf), 'vxb_nccl_unique_id')
                t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).to(self.flat.device)
                dist.broadcast(t, src=0)
                idbytes = bytes(t.cpu().tolist())
                comm = ctypes.c_void_p()
                torch.cuda.synchronize()
                _lib.check(L.vxb_nccl_init(idbytes, dist.get_rank(), world, ctypes.byref(comm)), 'vxb_nccl_init')
                self.comm = comm
            _lib.check(L.vxb_allreduce_grads(self.comm, _lib.ptr(self.flat), n, 1.0 / world, self.bucket_elems, _lib.stream()),
                       'vxb_allreduce_grads')
            return
        works = []
        for o in range(0, n, self.bucket_elems):
            works.append(dist.all_reduce(self.flat[o:min(n, o + self.bucket_elems)], op=dist.ReduceOp.SUM, async_op=True))
        for w in works:
            w.wait()
        self.flat.div_(world)
