"""Generate tests/golden/*.npz by running the REFERENCE's own modules (authoring container only).

    python tests/golden/make_golden.py

Imports VoxelGrid / PerceiverVoxelLangEncoder from /root/reference (unmodified), feeds them the
seeded synthetic inputs of voxactb_b200.synth and stores the outputs.  Inputs are regenerated from
the seeds at test time (a checksum of every input is stored to detect RNG drift); small inputs are
stored verbatim.  The GPU box has no /root/reference: these files are what travels.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import refimport  # noqa: E402
from voxactb_b200 import synth  # noqa: E402

QNET_CASES = {
    # name: dict(V, k, s, L, depth, B, cameras, H, W, low_dim, arm, crop, seed)
    'qnet_v20': dict(V=20, k=5, s=5, L=64, depth=2, B=2, cameras=2, H=32, W=32, low_dim=4, arm=False, crop=False, seed=11),
    'qnet_v20_arm_crop': dict(V=20, k=5, s=5, L=96, depth=1, B=3, cameras=2, H=32, W=32, low_dim=7, arm=True, crop=True, seed=12),
    'qnet_v32_config1': dict(V=32, k=5, s=4, L=2048, depth=6, B=1, cameras=1, H=128, W=128, low_dim=4, arm=False, crop=False, seed=1235),
    'qnet_v100_b1': dict(V=100, k=5, s=5, L=2048, depth=6, B=1, cameras=4, H=128, W=128, low_dim=4, arm=False, crop=False, seed=1236),
    # BASELINE.json configurations at their stated size (the reference runs them in chunks of `chunk` samples: samples
    # are independent, there is no batch coupling anywhere on the path)
    'qnet_v100_b16': dict(V=100, k=5, s=5, L=2048, depth=6, B=16, cameras=4, H=128, W=128, low_dim=4, arm=False, crop=False,
                          seed=1240, chunk=2, tstride=997),
    # config 3: acting + stabilizing agents (low_dim 7, arm head present) on the same observations, different weights
    'qnet_v100_acting': dict(V=100, k=5, s=5, L=2048, depth=6, B=2, cameras=4, H=128, W=128, low_dim=7, arm=True, crop=False,
                             seed=1241, wseed=2241),
    'qnet_v100_stabilizing': dict(V=100, k=5, s=5, L=2048, depth=6, B=2, cameras=4, H=128, W=128, low_dim=7, arm=True,
                                  crop=False, seed=1241, wseed=3241),
    # config 4: per-sample VLM-crop bounds
    'qnet_v100_crop': dict(V=100, k=5, s=5, L=2048, depth=6, B=2, cameras=4, H=128, W=128, low_dim=4, arm=False, crop=True,
                           seed=1242),
}
# 2-robot encoder (PerceiverVoxelLang2RobotsEncoder, C = 192, two head sets)
QNET2_CASES = {
    'qnet2_v20': dict(V=20, k=5, s=5, L=64, depth=2, B=2, cameras=2, H=32, W=32, low_dim=4, arm=False, crop=False, seed=31),
    'qnet2_v100_b1': dict(V=100, k=5, s=5, L=2048, depth=6, B=1, cameras=4, H=128, W=128, low_dim=4, arm=False, crop=False,
                          seed=1243),
}
# one training step (agent.update): forward in train mode with zero dropout, reference losses, autograd, reference LAMB
TRAIN_CASES = {
    'train_v20': dict(V=20, k=5, s=5, L=64, depth=2, B=2, cameras=2, H=32, W=32, low_dim=4, arm=False, crop=False, seed=41),
    'train_v20_arm': dict(V=20, k=5, s=5, L=48, depth=1, B=3, cameras=2, H=32, W=32, low_dim=7, arm=True, crop=True, seed=42),
}
VOXEL_CASES = {
    'voxel_v20': dict(V=20, B=2, cameras=2, H=32, W=32, crop=False, seed=21),
    'voxel_v32_crop': dict(V=32, B=3, cameras=1, H=48, W=40, crop=True, seed=22),
    'voxel_v100': dict(V=100, B=1, cameras=4, H=128, W=128, crop=False, seed=23),
}


def checksum(t):
    return float(t.double().abs().sum())


def ref_indices(vg, coords, bounds):
    """voxel_grid.py:152-163 evaluated with the reference module's own buffers."""
    bb_mins = bounds[..., 0:3]
    bb_maxs = bounds[..., 3:6]
    res = (bb_maxs - bb_mins) / (vg._dims_orig.float() + 1e-12)
    denom = res + 1e-12
    shifted = bb_mins - res
    fl = torch.floor((coords - shifted.unsqueeze(1)) / denom.unsqueeze(1)).int()
    return torch.max(torch.min(fl, vg._dims_m_one), vg._dims_m_one_zeros)


def encoder_kwargs(c):
    return dict(depth=c['depth'], iterations=1, voxel_size=c['V'], initial_dim=10, low_dim_size=c['low_dim'],
                layer=0, num_rotation_classes=72, num_grip_classes=2, num_collision_classes=2, input_axis=3,
                num_latents=c['L'], latent_dim=512, cross_heads=1, latent_heads=8, cross_dim_head=64,
                latent_dim_head=64, weight_tie_layers=False, activation='lrelu', pos_encoding_with_lang=True,
                input_dropout=0.1, attn_dropout=0.1, decoder_dropout=0.0, lang_fusion_type='seq',
                voxel_patch_size=c['k'], voxel_patch_stride=c['s'], no_skip_connection=False,
                no_perceiver=False, no_language=False, final_dim=64, arm_pred_loss=c['arm'])


def weight_seed(c):
    return c.get('wseed', c['seed'] + 1000)


def chunked(fn, B, chunk):
    """Run fn(slice) over the batch in chunks and concatenate each output along dim 0."""
    parts = [fn(slice(i, min(B, i + chunk))) for i in range(0, B, chunk)]
    return [torch.cat([p[j] for p in parts], 0) for j in range(len(parts[0]))]


def proprio_left(c):
    return torch.rand(c['B'], c['low_dim'], generator=torch.Generator().manual_seed(c['seed'] + 7))


def main_two_robots():
    RefVG, _ = refimport.load()
    RefEnc2 = refimport.load2()
    only = sys.argv[2:]
    for name, c in QNET2_CASES.items():
        if only and name not in only:
            continue
        obs = synth.make_observation(c['seed'], c['B'], c['cameras'], c['H'], c['W'], low_dim=c['low_dim'])
        coords, feats = synth.flatten_cameras(obs)
        vg = RefVG(synth.SCENE_BOUNDS, c['V'], 'cpu', c['B'], 3, coords.shape[1])
        grid = vg.coords_to_bounding_voxel_grid(coords, feats, obs['bounds']).permute(0, 4, 1, 2, 3)
        kw = encoder_kwargs(c)
        kw.pop('arm_pred_loss')
        net = RefEnc2(**kw).eval()
        sd = synth.random_state_dict(net, c['seed'] + 1000)
        missing = net.load_state_dict(sd, strict=False)
        assert not missing.unexpected_keys, missing
        with torch.no_grad():
            outs = net(grid, obs['proprio'], proprio_left(c), obs['lang_goal_emb'], obs['lang_token_embs'], None,
                       obs['bounds'], None)
        out = dict(keys=np.array(sorted(net.state_dict().keys())),
                   rot_grip=outs[1].numpy(), collision=outs[2].numpy(),
                   rot_grip_left=outs[4].numpy(), collision_left=outs[5].numpy())
        for key, t in (('trans', outs[0]), ('trans_left', outs[3])):
            if c['V'] <= 32:
                out[key] = t.numpy()
            else:
                out[key + '_strided'] = t.reshape(c['B'], -1)[:, ::97].numpy()
                out[key + '_argmax'] = t.reshape(c['B'], -1).argmax(-1).numpy()
                out[key + '_stats'] = np.array([float(t.double().sum()), float(t.double().abs().sum()), float(t.max()),
                                                float(t.min())])
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
        print(name, 'rot_grip_left absmax', float(np.abs(out['rot_grip_left']).max()))


def train_labels(c):
    """Seeded label indices of one replay batch."""
    g = torch.Generator().manual_seed(c['seed'] + 5000)
    lab = dict(trans=torch.randint(0, c['V'], (c['B'], 3), generator=g, dtype=torch.int32),
               rot_grip=torch.cat([torch.randint(0, 72, (c['B'], 3), generator=g, dtype=torch.int32),
                                   torch.randint(0, 2, (c['B'], 1), generator=g, dtype=torch.int32)], 1),
               collision=torch.randint(0, 2, (c['B'], 1), generator=g, dtype=torch.int32))
    if c['arm']:
        lab['arm'] = torch.randint(0, 2, (c['B'], 1), generator=g, dtype=torch.int32)
    return lab


def train_encoder_kwargs(c):
    kw = encoder_kwargs(c)
    kw.update(input_dropout=0.0, attn_dropout=0.0, decoder_dropout=0.0)   # dropout off for gradient parity
    return kw


def main_train():
    """agent.update (qattention_peract_bc_agent.py:484-582) re-enacted with the reference's own modules."""
    import importlib.util
    import torch.nn as nn
    RefVG, RefEnc = refimport.load()
    spec = importlib.util.spec_from_file_location('ref_lamb', os.path.join(refimport.REF, 'helpers', 'optim', 'lamb.py'))
    lamb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(lamb)
    torch.set_num_threads(os.cpu_count())
    celoss_fn = nn.CrossEntropyLoss(reduction='none')

    def celoss(pred, labels):                                  # agent:391-392
        return celoss_fn(pred, labels.argmax(-1))

    for name, c in TRAIN_CASES.items():
        obs = synth.make_observation(c['seed'], c['B'], c['cameras'], c['H'], c['W'], low_dim=c['low_dim'],
                                     per_sample_crop=c['crop'])
        lab = train_labels(c)
        coords, feats = synth.flatten_cameras(obs)
        vg = RefVG(synth.SCENE_BOUNDS, c['V'], 'cpu', c['B'], 3, coords.shape[1])
        grid = vg.coords_to_bounding_voxel_grid(coords, feats, obs['bounds']).permute(0, 4, 1, 2, 3).detach()
        net = RefEnc(**train_encoder_kwargs(c)).train()
        sd = synth.random_state_dict(net, c['seed'] + 1000)
        assert not net.load_state_dict(sd, strict=False).unexpected_keys
        outs = net(grid, obs['proprio'], obs['lang_goal_emb'], obs['lang_token_embs'], None, obs['bounds'], None)
        q_trans, q_rot_grip, q_collision = outs[0], outs[1], outs[2]
        bs, V, R = c['B'], c['V'], 72
        onehot = torch.zeros(bs, 1, V, V, V)
        for b in range(bs):                                    # agent:517-522
            gt = lab['trans'][b].int()
            onehot[b, :, gt[0], gt[1], gt[2]] = 1
        terms = {'trans': celoss(q_trans.view(bs, -1), onehot.view(bs, -1))}
        rx, ry, rz, gr, ic = (torch.zeros(bs, R), torch.zeros(bs, R), torch.zeros(bs, R), torch.zeros(bs, 2), torch.zeros(bs, 2))
        for b in range(bs):                                    # agent:538-546
            g = lab['rot_grip'][b].int()
            rx[b, g[0]] = 1; ry[b, g[1]] = 1; rz[b, g[2]] = 1; gr[b, g[3]] = 1
            ic[b, lab['collision'][b].int()[0]] = 1
        terms['rot'] = celoss(q_rot_grip[:, 0:R], rx) + celoss(q_rot_grip[:, R:2 * R], ry) + celoss(q_rot_grip[:, 2 * R:3 * R], rz)
        terms['grip'] = celoss(q_rot_grip[:, 3 * R:], gr)
        terms['collision'] = celoss(q_collision, ic)
        combined = terms['trans'] + terms['rot'] + terms['grip'] + terms['collision']      # all loss weights 1.0
        if c['arm']:
            arm1h = torch.zeros(bs, 2)
            for b in range(bs):                                # agent:565-570
                arm1h[b, lab['arm'][b].long()] = 1
            terms['arm'] = celoss(outs[3], arm1h)
            combined = combined + terms['arm']
        total = combined.mean()
        opt = lamb.Lamb(net.parameters(), lr=5e-4, weight_decay=1e-6, betas=(0.9, 0.999), adam=False)
        opt.zero_grad()
        total.backward()
        grads = {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}
        opt.step()
        new = {k: p.detach().clone() for k, p in net.named_parameters() if k in grads}
        keys = sorted(grads.keys())
        out = dict(total=np.array([float(total)]), keys=np.array(keys),
                   label_checksum=np.array([int(sum(int(v.long().sum()) for v in lab.values()))]),
                   grad_sum=np.array([float(grads[k].double().sum()) for k in keys]),
                   grad_abs=np.array([float(grads[k].double().abs().sum()) for k in keys]),
                   grad_max=np.array([float(grads[k].abs().max()) for k in keys]),
                   param_sum=np.array([float(new[k].double().sum()) for k in keys]),
                   param_delta_abs=np.array([float((new[k].double() - sd[k].double()).abs().sum()) for k in keys]),
                   no_grad_keys=np.array(sorted(k for k, p in net.named_parameters() if p.grad is None)))
        for t, v in terms.items():
            out['loss_' + t] = v.detach().numpy()
        for k in keys:                                         # full tensors for the small ones, strided samples otherwise
            g = grads[k].reshape(-1)
            out['g:' + k] = (g if g.numel() <= 4096 else g[:: max(1, g.numel() // 2048)]).numpy()
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
        print(name, 'total', float(total), 'params with grad', len(keys), 'without', len(out['no_grad_keys']))


DEPTH_CASES = {
    'depth_v20': dict(V=20, B=2, cameras=2, H=24, W=32, seed=51),
}


def main_depth():
    """Raw-depth fixtures (row f1): the reference's OWN back-projection (PyRep vision_sensor.py, loaded from the unmodified
    file with the simulator bindings stubbed) followed by the reference VoxelGrid."""
    import importlib.util
    from unittest import mock
    for name in ('pyrep', 'pyrep.backend', 'pyrep.backend.sim', 'pyrep.objects', 'pyrep.objects.object', 'pyrep.const'):
        sys.modules.setdefault(name, mock.MagicMock())
    sys.modules['pyrep.objects.object'].Object = object
    spec = importlib.util.spec_from_file_location('ref_vision_sensor', '/root/reference/PyRep/pyrep/objects/vision_sensor.py')
    vs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(vs)
    RefVG, _ = refimport.load()
    for name, c in DEPTH_CASES.items():
        o = synth.make_depth_observation(c['seed'], c['B'], c['cameras'], c['H'], c['W'])
        B, cams, H, W = o['depth'].shape
        pts = np.empty((B, cams * H * W, 3), dtype=np.float32)
        for b in range(B):
            for cam in range(cams):
                wc = vs.VisionSensor.pointcloud_from_depth_and_camera_params(o['depth'][b, cam].numpy(), o['extrinsics'][b, cam],
                                                                             o['intrinsics'][b, cam])
                pts[b, cam * H * W:(cam + 1) * H * W] = wc.reshape(-1, 3)
        coords = torch.from_numpy(pts)
        feats = o['rgb'].permute(0, 1, 3, 4, 2).reshape(B, -1, 3)
        vg = RefVG(synth.SCENE_BOUNDS, c['V'], 'cpu', B, 3, coords.shape[1])
        grid = vg.coords_to_bounding_voxel_grid(coords, feats, o['bounds'])
        idx = ref_indices(vg, coords, o['bounds'])
        np.savez_compressed(os.path.join(HERE, name + '.npz'), points=pts, idx=idx.numpy().astype(np.int16), grid=grid.numpy(),
                            in_checksum=np.array([checksum(o['depth']), float(np.abs(o['intrinsics']).sum()),
                                                  float(np.abs(o['extrinsics']).sum())]))
        print(name, 'occupied', int((grid[..., -1] > 0).sum()), 'in-bounds points', int(((idx > 0) & (idx < c['V'] + 1)).all(-1).sum()))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == 'depth':
        return main_depth()
    if len(sys.argv) > 1 and sys.argv[1] == 'two_robots':
        return main_two_robots()
    if len(sys.argv) > 1 and sys.argv[1] == 'train':
        return main_train()
    only = sys.argv[2:] if len(sys.argv) > 2 and sys.argv[1] == 'qnet' else []
    RefVG, RefEnc = refimport.load()
    torch.set_num_threads(os.cpu_count())
    for name, c in VOXEL_CASES.items():
        if only:
            continue
        obs = synth.make_observation(c['seed'], c['B'], c['cameras'], c['H'], c['W'], per_sample_crop=c['crop'])
        coords, feats = synth.flatten_cameras(obs)
        vg = RefVG(synth.SCENE_BOUNDS, c['V'], 'cpu', c['B'], 3, coords.shape[1])
        grid = vg.coords_to_bounding_voxel_grid(coords, feats, obs['bounds'])
        idx = ref_indices(vg, coords, obs['bounds'])
        out = dict(cfg=np.array([c['V'], c['B'], c['cameras'], c['H'], c['W'], int(c['crop']), c['seed']]),
                   in_checksum=np.array([checksum(coords), checksum(feats), checksum(obs['bounds'])]),
                   idx_checksum=np.array([int(idx.long().sum()), int((idx.long() * torch.arange(1, 4)).sum())]),
                   occupied=np.array([int((grid[..., -1] > 0).sum())]),
                   grid_sum=np.array([float(grid.double().sum()), float(grid.double().abs().sum())]))
        if c['V'] <= 32:
            out['grid'] = grid.numpy()
            out['idx'] = idx.numpy().astype(np.int16)
        else:
            nz = torch.nonzero(grid[..., -1].reshape(-1) > 0).reshape(-1)
            sel = nz[:: max(1, nz.numel() // 4096)]
            out['sample_pos'] = sel.numpy().astype(np.int64)
            out['sample_val'] = grid.reshape(-1, grid.shape[-1])[sel].numpy()
            out['idx_head'] = idx[:, :4096].numpy().astype(np.int16)
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
        print(name, 'occupied', out['occupied'])
    for name, c in QNET_CASES.items():
        if only and name not in only:
            continue
        obs = synth.make_observation(c['seed'], c['B'], c['cameras'], c['H'], c['W'], low_dim=c['low_dim'],
                                     per_sample_crop=c['crop'])
        coords, feats = synth.flatten_cameras(obs)
        net = RefEnc(**encoder_kwargs(c)).eval()
        sd = synth.random_state_dict(net, weight_seed(c))
        missing = net.load_state_dict(sd, strict=False)
        assert not missing.unexpected_keys, missing

        def run(sl):
            n = sl.stop - sl.start
            vg = RefVG(synth.SCENE_BOUNDS, c['V'], 'cpu', n, 3, coords.shape[1])
            bnd = obs['bounds'] if obs['bounds'].shape[0] == 1 else obs['bounds'][sl]
            grid = vg.coords_to_bounding_voxel_grid(coords[sl], feats[sl], bnd).permute(0, 4, 1, 2, 3)
            with torch.no_grad():
                return list(net(grid, obs['proprio'][sl], obs['lang_goal_emb'][sl], obs['lang_token_embs'][sl], None, bnd,
                                None))

        outs = chunked(run, c['B'], c.get('chunk', c['B']))
        trans = outs[0]
        out = dict(in_checksum=np.array([checksum(coords), checksum(feats), checksum(obs['proprio']),
                                         checksum(obs['lang_token_embs']), checksum(obs['bounds'])]),
                   sd_checksum=np.array([sum(checksum(v) for v in sd.values())]),
                   keys=np.array(sorted(net.state_dict().keys())),
                   rot_grip=outs[1].numpy(), collision=outs[2].numpy(),
                   trans_argmax=trans.reshape(c['B'], -1).argmax(-1).numpy(),
                   trans_stats=np.array([float(trans.double().sum()), float(trans.double().abs().sum()),
                                         float(trans.max()), float(trans.min())]))
        if c['arm']:
            out['arm'] = outs[3].numpy()
        if c['V'] <= 32:
            out['trans'] = trans.numpy()
        else:
            out['trans_strided'] = trans.reshape(c['B'], -1)[:, ::c.get('tstride', 97)].numpy()
            out['trans_sums'] = trans.reshape(c['B'], -1).double().sum(-1).numpy()
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
        print(name, 'trans range', out['trans_stats'][2:], 'rot_grip absmax', float(np.abs(out['rot_grip']).max()))


if __name__ == '__main__':
    main()
