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
