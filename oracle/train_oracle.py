"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's training step.

Follows /root/reference/peract/agents/peract_bc/qattention_peract_bc_agent.py:
  * losses            :391-392 (`_celoss` = CrossEntropyLoss(reduction='none') on `labels.argmax(-1)`),
                      :517-578 (one-hots from label indices; trans CE over V^3, three rotation CEs, grip, collision,
                                optional arm CE; weighted sum; mean over the batch)
  * backward          :581 (`total_loss.backward()`): autograd over the oracle forward (oracle/qnet_oracle.py), the
                      voxel grid is detached (`voxel_grid.detach()`, :107-108), dropout off for gradient parity
  * optimizer         :263-268 / helpers/optim/lamb.py:60-122 (oracle/optim_oracle.py), one step from zero state

Pinned by tests/test_oracle.py::test_training_step_oracle_matches_reference_golden against
tests/golden/train_v20.npz, which tests/golden/make_golden.py generates by running the reference modules
(PerceiverVoxelLangEncoder in train mode with zero dropout, torch autograd, the reference Lamb class).
Nothing in voxactb_b200/ imports this module.
"""
import torch
import torch.nn.functional as F

from . import optim_oracle, qnet_oracle, voxel_oracle


def peract_losses(q_trans, q_rot_grip, q_collision, action_trans, action_rot_grip, action_ignore_collisions,
                  arm_out=None, action_label=None, num_rotation_classes=72, weights=(1.0, 1.0, 1.0, 1.0, 1.0)):
    """agent:517-578 from label indices.  Returns (total, dict of per-sample loss terms)."""
    bs = q_trans.shape[0]
    V = q_trans.shape[-1]
    R = num_rotation_classes

    def ce(logits, idx):                                       # :391-392, labels.argmax(-1) == idx for a one-hot
        return F.cross_entropy(logits, idx.long(), reduction='none')

    flat = q_trans.reshape(bs, -1)                             # :525
    t_idx = (action_trans[:, 0].long() * V + action_trans[:, 1].long()) * V + action_trans[:, 2].long()   # :519-522
    terms = {'trans': ce(flat, t_idx)}                         # :527
    rot = 0.
    for a in range(3):                                         # :549-557
        rot = rot + ce(q_rot_grip[:, a * R:(a + 1) * R], action_rot_grip[:, a])
    terms['rot'] = rot
    terms['grip'] = ce(q_rot_grip[:, 3 * R:], action_rot_grip[:, 3])           # :552,560
    terms['collision'] = ce(q_collision, action_ignore_collisions.reshape(bs))  # :553,563
    combined = terms['trans'] * weights[0] + terms['rot'] * weights[1] + terms['grip'] * weights[2] + \
        terms['collision'] * weights[3]                        # :572-576
    if arm_out is not None:
        terms['arm'] = ce(arm_out, action_label.reshape(bs))   # :565-570
        combined = combined + terms['arm'] * weights[4]
    return combined.mean(), terms                              # :578


def training_step(sd, cfg, rgb, pcd, proprio, lang_token_embs, bounds, voxel_size, labels, lr=5e-4, weight_decay=1e-6,
                  betas=(0.9, 0.999), eps=1e-6, weights=(1.0, 1.0, 1.0, 1.0, 1.0)):
    """One `update`: forward, loss, backward, LAMB step from zero optimizer state.
    sd: state dict (fp32 tensors; not modified).  labels: dict(trans [B,3], rot_grip [B,4], collision [B,1][, arm [B,1]]).
    Returns dict(total, terms, grads {name: tensor}, params {name: tensor after the step})."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    with torch.enable_grad():
        out = qnet_oracle.qfunction_forward(params, cfg, voxel_oracle.voxelize, rgb, pcd, proprio, lang_token_embs, bounds,
                                            voxel_size)
        total, terms = peract_losses(out['trans'], out['rot_grip'], out['collision'], labels['trans'], labels['rot_grip'],
                                     labels['collision'], out.get('arm') if 'arm' in labels else None, labels.get('arm'),
                                     weights=weights)
        total.backward()
    grads, new = {}, {}
    for k, p in params.items():
        if p.grad is None:                                     # parameters the forward never touches (e.g. unused heads)
            continue
        grads[k] = p.grad.detach().clone()
        q = p.detach().clone()
        optim_oracle.lamb_step(q, grads[k], torch.zeros_like(q), torch.zeros_like(q), lr, betas[0], betas[1], eps, weight_decay)
        new[k] = q
    return {'total': total.detach(), 'terms': {k: v.detach() for k, v in terms.items()}, 'grads': grads, 'params': new}
