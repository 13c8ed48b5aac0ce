"""CPU oracle for the PerceiverActor Q-network forward -- TEST INFRASTRUCTURE ONLY.

A functional restatement (torch CPU fp32 ops on a plain ``state_dict``) of
``PerceiverVoxelLangEncoder.forward`` (reference
peract/agents/peract_bc/perceiver_lang_io.py:345-485) and of the blocks it uses
(peract/helpers/network_utils.py:128-170 Conv3DBlock, :237-254 Conv3DUpsampleBlock,
:257-289 DenseBlock, :773-809 SpatialSoftmax3D; perceiver_lang_io.py:56-132
PreNorm / GEGLU / FeedForward / Attention), plus the action-selection helpers of
``QFunction`` (peract/agents/peract_bc/qattention_peract_bc_agent.py:57-80) and the
``act`` tail (:709-724).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module; the product path
(``voxactb_b200``) never does.

Where the arithmetic lives: the reference executes these steps through PyTorch ATen
(pinned torch==1.7.1 / 1.13.1 in the reference's requirements; torch 2.11 here --
same operator semantics, summation order differs by backend, hence a float
tolerance rather than bit equality).

Parity pin: no reference test holds golden vectors for this path.
This restatement is pinned against the reference's own modules, imported live in the
authoring container (tests/test_oracle.py), and through
tests/golden/*.npz produced by tests/golden/make_golden.py from those modules.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.02  # network_utils.py:12


def _act(x, activation):
    if activation == 'lrelu':
        return F.leaky_relu(x, LRELU_SLOPE)
    if activation == 'relu':
        return F.relu(x)
    if activation is None:
        return x
    raise ValueError(activation)


def conv3d_block(x, w, b, stride, activation):
    """Conv3DBlock.forward, network_utils.py:128-170: replicate padding k//2, conv, activation."""
    k = w.shape[-1]
    p = k // 2
    if p > 0:
        x = F.pad(x, (p, p, p, p, p, p), mode='replicate')
    return _act(F.conv3d(x, w, b, stride=stride), activation)


def spatial_softmax3d(x, temperature=0.01):
    """SpatialSoftmax3D.forward, network_utils.py:797-809, including the meshgrid('xy') axis quirk
    of :782-792 (pos_x runs along tensor axis H, pos_y along D, pos_z along W)."""
    b, c, d, h, w = x.shape
    pos_x, pos_y, pos_z = np.meshgrid(np.linspace(-1., 1., d), np.linspace(-1., 1., h),
                                      np.linspace(-1., 1., w))
    pos_x = torch.from_numpy(pos_x.reshape(-1)).float().to(x.device)
    pos_y = torch.from_numpy(pos_y.reshape(-1)).float().to(x.device)
    pos_z = torch.from_numpy(pos_z.reshape(-1)).float().to(x.device)
    feat = x.contiguous().view(-1, h * w * d)
    att = F.softmax(feat / temperature, dim=-1)
    ex = torch.sum(pos_x * att, dim=1, keepdim=True)
    ey = torch.sum(pos_y * att, dim=1, keepdim=True)
    ez = torch.sum(pos_z * att, dim=1, keepdim=True)
    return torch.cat([ex, ey, ez], 1).view(-1, c * 3)


def _layernorm(x, sd, prefix):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + '.weight'], sd[prefix + '.bias'])


def _attention(x, context, sd, prefix, heads):
    """Attention.forward, perceiver_lang_io.py:107-132 (eval mode: dropout is the identity)."""
    q = F.linear(x, sd[prefix + '.to_q.weight'])
    kv = F.linear(context, sd[prefix + '.to_kv.weight'])
    k, v = kv.chunk(2, dim=-1)
    b, n, inner = q.shape
    dh = inner // heads

    def split(t):
        return t.view(t.shape[0], t.shape[1], heads, dh).permute(0, 2, 1, 3).reshape(-1, t.shape[1], dh)

    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum('bid,bjd->bij', q, k) * (dh ** -0.5)
    attn = sim.softmax(dim=-1)
    out = torch.einsum('bij,bjd->bid', attn, v)
    out = out.view(b, heads, n, dh).permute(0, 2, 1, 3).reshape(b, n, inner)
    return F.linear(out, sd[prefix + '.to_out.weight'], sd[prefix + '.to_out.bias'])


def _feedforward(x, sd, prefix):
    """FeedForward / GEGLU, perceiver_lang_io.py:74-90."""
    h = F.linear(x, sd[prefix + '.net.0.weight'], sd[prefix + '.net.0.bias'])
    a, g = h.chunk(2, dim=-1)
    return F.linear(a * F.gelu(g), sd[prefix + '.net.2.weight'], sd[prefix + '.net.2.bias'])


def qnet_forward(sd, cfg, ins, proprio, lang_token_embs):
    """PerceiverVoxelLangEncoder.forward, perceiver_lang_io.py:345-485, for lang_fusion_type='seq',
    pos_encoding_with_lang=True (the configuration launch_utils.py:744-775 builds).

    sd  : state_dict of the encoder (reference parameter names)
    cfg : dict(voxel_patch_size, voxel_patch_stride, depth, iterations, cross_heads, latent_heads,
               activation, num_collision_classes, arm_pred_loss, no_language)
    ins : [B,10,V,V,V]; proprio [B,low]; lang_token_embs [B,77,512]
    returns dict(trans, rot_grip, collision[, arm])
    """
    act = cfg.get('activation', 'lrelu')
    stride = cfg['voxel_patch_stride']
    bsz = ins.shape[0]
    d0 = conv3d_block(ins, sd['input_preprocess.conv3d.weight'], sd['input_preprocess.conv3d.bias'], 1, act)
    feats = [spatial_softmax3d(d0), d0.amax(dim=(2, 3, 4))]                      # :360
    x = conv3d_block(d0, sd['patchify.conv3d.weight'], sd['patchify.conv3d.bias'], stride, act)  # :363
    b, c, d, h, w = x.shape
    p = _act(F.linear(proprio, sd['proprio_preprocess.linear.weight'],
                      sd['proprio_preprocess.linear.bias']), act)               # :371
    p = p[:, :, None, None, None].expand(-1, -1, d, h, w)
    x = torch.cat([x, p], dim=1)                                                # :373
    if cfg.get('two_robots', False):
        # PerceiverVoxelLang2RobotsEncoder.forward, perceiver_lang_io.py:723-729: the SAME proprio block on the left arm
        p2 = _act(F.linear(cfg['proprio_left'], sd['proprio_preprocess.linear.weight'],
                           sd['proprio_preprocess.linear.bias']), act)
        x = torch.cat([x, p2[:, :, None, None, None].expand(-1, -1, d, h, w)], dim=1)
    if cfg.get('no_language', False):
        lang_token_embs = torch.zeros_like(lang_token_embs)                     # :376-378
    x = x.permute(0, 2, 3, 4, 1)                                                # :389
    seq_shape = x.shape
    x = x.reshape(b, -1, x.shape[-1])                                           # :412
    l = F.linear(lang_token_embs, sd['lang_preprocess.weight'], sd['lang_preprocess.bias'])  # :417
    if sd['pos_encoding'].dim() == 5:
        # pos_encoding_with_lang=False (:392-393): the encoding [1,S,S,S,C] goes on the voxel tokens only, the language tokens
        # stay a bag of words (the configuration of the PerAct paper, see the note at :395-406)
        seq = torch.cat((l, x + sd['pos_encoding'].reshape(1, -1, x.shape[-1])), dim=1)
    else:
        seq = torch.cat((l, x), dim=1) + sd['pos_encoding']                     # :418,422
    lat = sd['latents'].unsqueeze(0).expand(b, -1, -1)                          # :425
    for _ in range(cfg.get('iterations', 1)):
        ctx = _layernorm(seq, sd, 'cross_attend_blocks.0.norm_context')
        lat = _attention(_layernorm(lat, sd, 'cross_attend_blocks.0.norm'), ctx, sd,
                         'cross_attend_blocks.0.fn', cfg.get('cross_heads', 1)) + lat    # :431
        lat = _feedforward(_layernorm(lat, sd, 'cross_attend_blocks.1.norm'), sd,
                           'cross_attend_blocks.1.fn') + lat                            # :432
        for i in range(cfg['depth']):                                                   # :435-437
            xn = _layernorm(lat, sd, 'layers.%d.0.norm' % i)
            lat = _attention(xn, xn, sd, 'layers.%d.0.fn' % i, cfg.get('latent_heads', 8)) + lat
            lat = _feedforward(_layernorm(lat, sd, 'layers.%d.1.norm' % i), sd, 'layers.%d.1.fn' % i) + lat
    dec = _attention(_layernorm(seq, sd, 'decoder_cross_attn.norm'),
                     _layernorm(lat, sd, 'decoder_cross_attn.norm_context'), sd,
                     'decoder_cross_attn.fn', cfg.get('cross_heads', 1))                # :440
    dec = dec[:, l.shape[1]:]                                                           # :444
    dec = dec.view(b, *seq_shape[1:-1], dec.shape[-1]).permute(0, 4, 1, 2, 3)           # :447-448
    feats.extend([spatial_softmax3d(dec.contiguous()), dec.amax(dim=(2, 3, 4))])        # :451
    # up0 = Conv3DUpsampleBlock: conv -> trilinear upsample -> conv  (network_utils.py:237-254)
    u0 = low = conv3d_block(dec, sd['up0.conv_up.0.conv3d.weight'], sd['up0.conv_up.0.conv3d.bias'], 1, act)
    if stride > 1:
        u0 = F.interpolate(u0, scale_factor=stride, mode='trilinear', align_corners=False)
        u0 = conv3d_block(u0, sd['up0.conv_up.2.conv3d.weight'], sd['up0.conv_up.2.conv3d.bias'], 1, act)
    else:
        u0 = conv3d_block(u0, sd['up0.conv_up.1.conv3d.weight'], sd['up0.conv_up.1.conv3d.bias'], 1, act)
    if cfg.get('no_skip_connection', False):
        u = conv3d_block(u0, sd['final.conv3d.weight'], sd['final.conv3d.bias'], 1, act)                      # :457
    elif cfg.get('no_perceiver', False):
        u = conv3d_block(d0, sd['final.conv3d.weight'], sd['final.conv3d.bias'], 1, act)                      # :459
    else:
        u = conv3d_block(torch.cat([d0, u0], dim=1), sd['final.conv3d.weight'], sd['final.conv3d.bias'], 1, act)  # :462
    trans = conv3d_block(u, sd['trans_decoder.conv3d.weight'], sd['trans_decoder.conv3d.bias'], 1, None)     # :465
    feats.extend([spatial_softmax3d(u), u.amax(dim=(2, 3, 4))])                          # :470
    flat = torch.cat(feats, dim=1)
    if cfg.get('tap') is not None:
        # intermediate activations for the backward parity tests (tests call .retain_grad() on them)
        cfg['tap'].update(d0=d0, u0=u0, u=u, low=low, dec=dec, latents=lat, tokens=seq, feats=flat)
    h0 = _act(F.linear(flat, sd['dense0.linear.weight'], sd['dense0.linear.bias']), act)
    h1 = _act(F.linear(h0, sd['dense1.linear.weight'], sd['dense1.linear.bias']), act)
    rgc = F.linear(h1, sd['rot_grip_collision_ff.linear.weight'], sd['rot_grip_collision_ff.linear.bias'])
    ncol = cfg.get('num_collision_classes', 2)
    out = {'trans': trans, 'rot_grip': rgc[:, :-ncol], 'collision': rgc[:, -ncol:], 'feats': flat}
    if cfg.get('two_robots', False):                                                     # :842-858
        out['trans_left'] = conv3d_block(u, sd['trans_decoder_left_arm.conv3d.weight'],
                                         sd['trans_decoder_left_arm.conv3d.bias'], 1, None)
        g0 = _act(F.linear(flat, sd['dense0_left_arm.linear.weight'], sd['dense0_left_arm.linear.bias']), act)
        g1 = _act(F.linear(g0, sd['dense1_left_arm.linear.weight'], sd['dense1_left_arm.linear.bias']), act)
        rgc2 = F.linear(g1, sd['rot_grip_collision_ff_left_arm.linear.weight'],
                        sd['rot_grip_collision_ff_left_arm.linear.bias'])
        out['rot_grip_left'], out['collision_left'] = rgc2[:, :-ncol], rgc2[:, -ncol:]
    if cfg.get('arm_pred_loss', False):                                                  # :479-483
        h2 = _act(F.linear(flat, sd['dense2.linear.weight'], sd['dense2.linear.bias']), act)
        out['arm'] = F.linear(h2, sd['arm_ff.linear.weight'], sd['arm_ff.linear.bias'])
    assert bsz == out['trans'].shape[0]
    return out


def qfunction_forward(sd, cfg, voxelize_fn, rgb, pcd, proprio, lang_token_embs, bounds, voxel_size):
    """QFunction.forward, qattention_peract_bc_agent.py:82-135: flatten the per-camera images,
    voxelise, permute to channels-first, run the Q-network.  rgb/pcd: lists of [B,3,H,W]."""
    b = rgb[0].shape[0]
    pcd_flat = torch.cat([p.permute(0, 2, 3, 1).reshape(b, -1, 3) for p in pcd], 1)
    feat = torch.cat([p.permute(0, 2, 3, 1).reshape(b, -1, 3) for p in rgb], 1)
    grid = voxelize_fn(pcd_flat.numpy(), feat.numpy(), bounds.numpy(), voxel_size)
    grid = torch.from_numpy(grid).permute(0, 4, 1, 2, 3).to(proprio.dtype)   # float64 when the caller asks for a double-precision run
    out = qnet_forward(sd, cfg, grid, proprio, lang_token_embs)
    out['voxel_grid'] = grid
    return out


def argmax_3d(q_trans):
    """QFunction._argmax_3d, qattention_peract_bc_agent.py:57-63."""
    b, c, d, h, w = q_trans.shape
    idxs = q_trans.view(b, c, -1).argmax(-1)
    return torch.cat([(idxs // h) // d, (idxs // h) % w, idxs % w], 1)


def choose_highest_action(q_trans, q_rot_grip, q_collision, rotation_resolution=5):
    """QFunction.choose_highest_action, qattention_peract_bc_agent.py:65-80."""
    coords = argmax_3d(q_trans)
    n = int(360 // rotation_resolution)
    q_rot = torch.stack(torch.split(q_rot_grip[:, :-2], n, dim=1), dim=1)
    rg = torch.cat([q_rot[:, 0:1].argmax(-1), q_rot[:, 1:2].argmax(-1), q_rot[:, 2:3].argmax(-1),
                    q_rot_grip[:, -2:].argmax(-1, keepdim=True)], -1)
    coll = q_collision[:, -2:].argmax(-1, keepdim=True)
    return coords, rg, coll


def attention_coordinate(coords, bounds, voxel_size):
    """act() tail, qattention_peract_bc_agent.py:701,724: bounds_min + res*idx + res/2."""
    res = (bounds[:, 3:] - bounds[:, :3]) / voxel_size
    return bounds[:, :3] + res * coords.int() + res / 2
