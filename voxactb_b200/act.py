"""act() tail and per-episode language caching.

``FusedActor`` mirrors the inference half of QAttentionPerActBCAgent.act + QAttentionStackAgent.act (reference
qattention_peract_bc_agent.py:643-787, qattention_stack_agent.py:46-98) for one QFunction: forward, fused arg-max selection
(vxb_select_action_f32) and the 9-D continuous action (vxb_act_tail_f32) stay on the device; ONE 36-byte device->host copy per
step replaces the reference's chain of `.cpu().numpy()` conversions.  ``CachedLanguageEncoder`` removes the CLIP text
transformer from the step loop: the reference re-encodes the (unchanged) instruction on every act() (agent:663-665)."""
import ctypes

import torch

from . import _lib


class CachedLanguageEncoder:
    """Wraps an ``encode(tokens) -> (lang_goal_emb, lang_token_embs)`` callable (e.g. CLIP's
    ``encode_text_with_embeddings``) with a cache keyed on the token ids: an instruction is encoded once per episode."""

    def __init__(self, encode, max_entries=64):
        self._encode = encode
        self._cache = {}
        self._max = max_entries
        self.hits = self.misses = 0

    def __call__(self, tokens):
        key = tokens.detach().cpu().numpy().tobytes()
        hit = self._cache.get(key)
        if hit is not None:
            self.hits += 1
            return hit
        self.misses += 1
        with torch.no_grad():
            emb, tok = self._encode(tokens)
        if len(self._cache) >= self._max:
            self._cache.pop(next(iter(self._cache)))
        self._cache[key] = (emb.detach(), tok.detach())
        return self._cache[key]

    def reset(self):
        self._cache.clear()


class FusedActor:
    def __init__(self, q, rotation_resolution=5):
        self.q = q
        self.rotation_resolution = float(rotation_resolution)
        self._host = None

    @torch.no_grad()
    def act(self, rgb_pcd, proprio, pcd, lang_goal_emb, lang_token_embs, bounds):
        """Returns (continuous_action [B,9] on the HOST (pinned), dict of device tensors: coords, rot_grip, collision, xyz)."""
        out = self.q(rgb_pcd, proprio, pcd, lang_goal_emb, lang_token_embs, bounds, None, None)
        coords, rg, coll, xyz = self.q.select_action(out[0], out[1], out[2], bounds)
        B = coords.shape[0]
        action = torch.empty(B, 9, dtype=torch.float32, device=coords.device)
        rc = _lib.lib().vxb_act_tail_f32(_lib.ptr(rg), _lib.ptr(coll), _lib.ptr(xyz), self.rotation_resolution, _lib.ptr(action),
                                         B, _lib.stream())
        _lib.check(rc, 'vxb_act_tail_f32')
        if self._host is None or self._host.shape[0] != B:
            self._host = torch.empty(B, 9, dtype=torch.float32).pin_memory()
        self._host.copy_(action, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._host, {'coords': coords, 'rot_grip': rg, 'collision': coll, 'attention_xyz': xyz, 'voxel_grid': out[3]}


class HostStager:
    """Double-buffered host -> device staging of observation batches (the device side of the
    replay / observation pipeline): ``stage(host)`` enqueues the copies of the NEXT batch on a private copy stream, so they
    overlap the Q-network of the current one; ``acquire()`` makes the compute stream wait for the oldest staged batch and
    returns its device tensors.  ``host`` is a dict of pinned CPU tensors or lists of them; the device buffers are allocated
    once per slot and reused, guarded by events in both directions (a slot is not overwritten before the step that read it
    has finished)."""

    def __init__(self, device, slots=2):
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.slots = [dict(buf=None, ready=torch.cuda.Event(), free=torch.cuda.Event(), staged=False) for _ in range(slots)]
        self._w = self._r = 0

    def _like(self, v):
        if isinstance(v, (list, tuple)):
            return [self._like(t) for t in v]
        return torch.empty(v.shape, dtype=v.dtype, device=self.device)

    @staticmethod
    def _copy(dst, src):
        if isinstance(src, (list, tuple)):
            for d, t in zip(dst, src):
                HostStager._copy(d, t)
        else:
            dst.copy_(src, non_blocking=True)

    def stage(self, host):
        slot = self.slots[self._w]
        if slot['staged']:
            raise RuntimeError('HostStager: every slot holds a batch that has not been acquired yet')
        if slot['buf'] is None:
            slot['buf'] = {k: self._like(v) for k, v in host.items()}
        else:
            self.copy_stream.wait_event(slot['free'])             # the step that read this slot has finished
        with torch.cuda.stream(self.copy_stream):
            for k, v in host.items():
                self._copy(slot['buf'][k], v)
            slot['ready'].record(self.copy_stream)
        slot['staged'] = True
        self._w = (self._w + 1) % len(self.slots)

    def acquire(self):
        slot = self.slots[self._r]
        if not slot['staged']:
            raise RuntimeError('HostStager: acquire() without a staged batch')
        torch.cuda.current_stream(self.device).wait_event(slot['ready'])
        slot['staged'] = False
        self._cur = slot
        self._r = (self._r + 1) % len(self.slots)
        return slot['buf']

    def release(self):
        """Call after the last kernel that reads the acquired batch has been enqueued on the compute stream."""
        self._cur['free'].record(torch.cuda.current_stream(self.device))


class GraphedActor:
    """Closed-loop acting (eval.py runs batch 1, reference qattention_peract_bc_agent.py:231): the whole step -- flatten,
    voxelize, Q-network, fused arg-max selection, 9-D act tail -- is captured ONCE per input geometry in a CUDA graph and
    replayed, which removes the ~120 kernel launches, the ctypes calls and the parameter-table rebuild from every step.
    Inputs are copied into the graph's static buffers, the 9-D action comes back through one pinned 36-byte-per-sample copy.
    The graph is re-captured when the geometry changes or a parameter has been modified (its ``_version`` moved)."""

    def __init__(self, q, rotation_resolution=5):
        self.actor = FusedActor(q, rotation_resolution)
        self.q = q
        self._key = None
        self._graph = None
        self._static = None
        self._out = None
        self._host = None
        self.captures = 0

    @staticmethod
    def _flat(rgb_pcd, proprio, pcd, lang_goal_emb, lang_token_embs, bounds):
        return [t for rp in rgb_pcd for t in rp] + list(pcd) + [proprio, lang_goal_emb, lang_token_embs, bounds]

    def _signature(self, tensors):
        return (tuple((tuple(t.shape), t.dtype, t.device) for t in tensors),
                tuple(p._version for p in self.q.parameters()), int(self.q._qnet.math_mode))

    def _run(self, st):
        ncam = self._ncam
        rgb_pcd = [[st[2 * i], st[2 * i + 1]] for i in range(ncam)]
        pcd = st[2 * ncam:3 * ncam]
        proprio, goal, tok, bounds = st[3 * ncam:]
        q = self.q
        out = q(rgb_pcd, proprio, pcd, goal, tok, bounds, None, None)
        coords, rg, coll, xyz = q.select_action(out[0], out[1], out[2], bounds)
        B = coords.shape[0]
        action = torch.empty(B, 9, dtype=torch.float32, device=coords.device)
        rc = _lib.lib().vxb_act_tail_f32(_lib.ptr(rg), _lib.ptr(coll), _lib.ptr(xyz), self.actor.rotation_resolution,
                                         _lib.ptr(action), B, _lib.stream())
        _lib.check(rc, 'vxb_act_tail_f32')
        return action, {'coords': coords, 'rot_grip': rg, 'collision': coll, 'attention_xyz': xyz, 'q_trans': out[0],
                        'voxel_grid': out[3]}

    @torch.no_grad()
    def act(self, rgb_pcd, proprio, pcd, lang_goal_emb, lang_token_embs, bounds):
        """Same arguments as QFunction.forward (device tensors).  Returns (action [B,9] pinned host tensor, dict of the
        graph's static device outputs -- valid until the next act())."""
        tensors = self._flat(rgb_pcd, proprio, pcd, lang_goal_emb, lang_token_embs, bounds)
        sig = self._signature(tensors)
        if sig != self._key:
            self._ncam = len(rgb_pcd)
            self._static = [t.detach().clone() for t in tensors]
            side = torch.cuda.Stream(tensors[0].device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):          # warm-up outside the capture: workspaces, prepared weights, kernel attributes
                self._run(self._static)
                self._run(self._static)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._out = self._run(self._static)
            self._graph, self._key = g, sig
            self.captures += 1
            B = self._out[0].shape[0]
            self._host = torch.empty(B, 9, dtype=torch.float32).pin_memory()
        for s, t in zip(self._static, tensors):
            if s.data_ptr() != t.data_ptr():
                s.copy_(t, non_blocking=True)
        self._graph.replay()
        self._host.copy_(self._out[0], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._host, self._out[1]
