"""Seeded synthetic observations for parity tests and bench.py.

CPU-generated (deterministic for a given torch version), shaped like what
``PreprocessAgent`` hands to ``QFunction.forward``: per camera a world-frame point cloud
``[B,3,H,W]`` and an RGB image already normalised to [-1,1] (reference
helpers/preprocess_agent.py:20-21), proprio, CLIP token-embedding stand-ins, scene bounds.
80 % of the pixels lie on a few planes/spheres inside the bounds (surface-like: exercises atomic
contention), 15 % are uniform in the bounds and 5 % uniform in bounds inflated by 20 %
(out-of-bounds points -> the clamped/cropped border path).
"""
import torch

SCENE_BOUNDS = [-0.3, -0.5, 0.6, 0.7, 0.5, 1.6]  # reference conf/config.yaml:15


def make_point_cloud(gen, B, H, W, bounds):
    mn = torch.tensor(bounds[:3])
    mx = torch.tensor(bounds[3:])
    ext = mx - mn
    n = H * W
    pts = torch.empty(B, n, 3)
    for b in range(B):
        kind = torch.rand(n, generator=gen)
        u = torch.rand(n, 3, generator=gen)
        uniform_in = mn + u * ext
        inflated = (mn - 0.1 * ext) + u * (1.2 * ext)
        nsurf = int(torch.randint(3, 6, (1,), generator=gen))
        which = torch.randint(0, nsurf, (n,), generator=gen)
        surf = torch.empty(n, 3)
        for s in range(nsurf):
            sel = which == s
            m = int(sel.sum())
            if m == 0:
                continue
            uv = torch.rand(m, 3, generator=gen)
            if s % 2 == 0:  # axis-aligned plane with a little thickness noise
                axis = int(torch.randint(0, 3, (1,), generator=gen))
                level = float(torch.rand(1, generator=gen))
                p = mn + uv * ext
                p[:, axis] = mn[axis] + ext[axis] * (level + 0.002 * (uv[:, axis] - 0.5))
            else:  # sphere shell
                centre = mn + (0.25 + 0.5 * torch.rand(3, generator=gen)) * ext
                radius = 0.1 + 0.15 * float(torch.rand(1, generator=gen))
                d = torch.randn(m, 3, generator=gen)
                d = d / d.norm(dim=1, keepdim=True).clamp_min(1e-6)
                p = centre + radius * d
            surf[sel] = p
        out = torch.where((kind < 0.8)[:, None], surf,
                          torch.where((kind < 0.95)[:, None], uniform_in, inflated))
        pts[b] = out
    return pts.view(B, H, W, 3).permute(0, 3, 1, 2).contiguous()


def make_observation(seed, B, cameras=4, H=128, W=128, low_dim=4, bounds=None, per_sample_crop=False,
                     crop_radius=0.3):
    """Returns dict(rgb=[cam x [B,3,H,W]], pcd=[...], proprio [B,low], lang_goal_emb [B,1024],
    lang_token_embs [B,77,512], bounds [1|B,6]) as CPU fp32 tensors."""
    gen = torch.Generator().manual_seed(int(seed))
    scene = list(SCENE_BOUNDS if bounds is None else bounds)
    pcd, rgb = [], []
    for _ in range(cameras):
        pcd.append(make_point_cloud(gen, B, H, W, scene))
        img = torch.randint(0, 256, (B, 3, H, W), generator=gen).float()
        rgb.append(img / 255.0 * 2.0 - 1.0)
    proprio = torch.rand(B, low_dim, generator=gen)
    lang_goal = torch.randn(B, 1024, generator=gen)
    lang_tok = torch.randn(B, 77, 512, generator=gen)
    if per_sample_crop:
        # VLM crop: bounds = round(target, 2) -/+ crop_radius  (reference helpers/utils.py:32-40)
        mn = torch.tensor(scene[:3])
        mx = torch.tensor(scene[3:])
        target = mn + torch.rand(B, 3, generator=gen) * (mx - mn)
        target = torch.round(target * 100) / 100
        bnd = torch.cat([target - crop_radius, target + crop_radius], dim=1)
    else:
        bnd = torch.tensor(scene).view(1, 6)
    return dict(rgb=rgb, pcd=pcd, proprio=proprio, lang_goal_emb=lang_goal, lang_token_embs=lang_tok,
                bounds=bnd.float())


def flatten_cameras(obs):
    """[B,N,3] coords and features exactly as QFunction.forward flattens them (agent:86-93)."""
    b = obs['pcd'][0].shape[0]
    coords = torch.cat([p.permute(0, 2, 3, 1).reshape(b, -1, 3) for p in obs['pcd']], 1)
    feats = torch.cat([p.permute(0, 2, 3, 1).reshape(b, -1, 3) for p in obs['rgb']], 1)
    return coords.contiguous(), feats.contiguous()


def random_state_dict(encoder, seed, scale_bias=0.05):
    """Seeded parameters for an encoder (ours or the reference's): the constructor's own init plus
    non-zero biases / perturbed LayerNorm affine so that no term of the forward is silently zero."""
    import zlib
    sd = {}
    for name, p in encoder.state_dict().items():
        if name.endswith(('pos_x', 'pos_y', 'pos_z')):
            continue
        # one generator per tensor, keyed by its NAME: the values do not depend on the order in
        # which an implementation registers its parameters
        gen = torch.Generator().manual_seed(int(seed) * 1000003 + zlib.crc32(name.encode()))
        shape = tuple(p.shape)
        r = torch.randn(shape, generator=gen)
        if name in ('pos_encoding', 'latents'):
            t = r
        elif name.endswith('.bias'):
            t = scale_bias * r
        elif 'norm' in name and name.endswith('.weight'):
            t = 1.0 + 0.1 * r
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = (2.0 / max(fan_in, 1)) ** 0.5 * r
        sd[name] = t.float()
    return sd


def make_depth_observation(seed, B, cameras=4, H=128, W=128, bounds=None):
    """Raw-depth observations: per camera a depth image in metres, RLBench-style intrinsics
    (negative focal lengths, vision_sensor.py:176-189) and a camera-to-world pose looking at the scene centre, such that the
    back-projected points fall in and around the scene bounds.  Returns dict(depth [B,cams,H,W] fp32, intrinsics
    [B,cams,3,3] f64, extrinsics [B,cams,4,4] f64, rgb [B,cams,3,H,W] fp32 in [-1,1], bounds [1,6])."""
    import math
    import numpy as np
    gen = torch.Generator().manual_seed(int(seed))
    scene = list(SCENE_BOUNDS if bounds is None else bounds)
    centre = np.array([(scene[i] + scene[i + 3]) / 2 for i in range(3)])
    depth = torch.empty(B, cameras, H, W)
    K = np.zeros((B, cameras, 3, 3))
    E = np.zeros((B, cameras, 4, 4))
    for b in range(B):
        for c in range(cameras):
            ang = float(torch.rand(1, generator=gen)) * 2 * math.pi
            elev = 0.3 + 0.6 * float(torch.rand(1, generator=gen))
            dist = 1.2 + 0.4 * float(torch.rand(1, generator=gen))
            pos = centre + dist * np.array([math.cos(ang) * math.cos(elev), math.sin(ang) * math.cos(elev), math.sin(elev)])
            fwd = (centre - pos) / np.linalg.norm(centre - pos)          # camera +z looks at the scene
            right = np.cross(fwd, np.array([0., 0., 1.]))
            right /= np.linalg.norm(right)
            up = np.cross(right, fwd)
            R = np.stack([-right, -up, fwd], 1)                           # negative focal lengths flip x / y
            E[b, c, :3, :3] = R
            E[b, c, :3, 3] = pos
            E[b, c, 3, 3] = 1.0
            f = -W / (2 * math.tan(math.radians(60.0) / 2))
            K[b, c] = np.array([[f, 0., W / 2], [0., f * H / W if H != W else f, H / 2], [0., 0., 1.]])
            base = dist - 0.5 + 1.0 * torch.rand(H, W, generator=gen)     # a slab of depths through the workspace
            plane = dist + 0.1 * torch.rand(1, generator=gen)              # plus a surface (contention)
            sel = torch.rand(H, W, generator=gen) < 0.6
            depth[b, c] = torch.where(sel, plane.expand(H, W) + 0.002 * torch.rand(H, W, generator=gen), base)
    rgb = torch.randint(0, 256, (B, cameras, 3, H, W), generator=gen).float() / 255.0 * 2.0 - 1.0
    return dict(depth=depth, intrinsics=K, extrinsics=E, rgb=rgb, bounds=torch.tensor([scene], dtype=torch.float32))
