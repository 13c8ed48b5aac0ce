#!/usr/bin/env python
"""bench.py -- policy forward passes/sec at 100^3 voxels (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 16] [--impl ours|reference]

A step = one pass of the hot path over one batch of synthetic observations per GPU:
voxelize (4 cameras x 128x128 RGB-D points -> 100^3 grid) + PerceiverActor Q-network forward.
N > 1: launched under torchrun, one process per GPU, the batch shards across ranks with no
data-path collective (weak scaling: --batch samples per GPU).

`value`  : whole-job passes/s with the observations already resident in HBM.
`e2e`    : the same through the public API (QFunction.forward + fused action selection) from
           pinned HOST buffers, host->device copies of the observations and device->host read of the
           selected actions inside the timed region.
`roofline`: the dominant kernel (conv3_umma_kernel: final 3x3x3 conv 128->64 at 100^3) timed live with CUDA
           events on the launching stream (library stage profiler), algorithmic FLOPs / time vs the measured
           tensor peak in MEASURED_PEAKS.json; `traffic` = DRAM bytes per launch from the committed
           ncu --set full capture (profiles/ncu_r01_conv3_umma_b16_summary.json, same batch).
`cpu_baseline`: the CPU oracle port (oracle/) of the reference's PyTorch path on the host cores, on
           a bounded sample (batch 1) of the same workload.
--impl reference: times that CPU implementation alone (the reference itself is Python and is not
           present on the GPU box; oracle/ is its restatement, pinned by tests/test_oracle.py).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_PASS = 1672729956352          # BASELINE.md section 4 (single-arm, V=100, k=s=5, L=2048, depth 6)
FINAL_CONV_FLOPS = 2 * 100 ** 3 * 64 * 128 * 27      # 442.4 GF / sample (direct convolution)
UPCONV_FLOPS = 2 * 100 ** 3 * 64 * 64 * 125          # 1024 GF / sample (direct convolution, as BASELINE.md counts it)
TRAIN_LAUNCHES_PER_STEP = 1205                     # ncu launch list of one training step (profiles/launches_r02_train.csv)
VOXELIZE_BYTES = 65536 * 24 + 100 ** 3 * 10 * 4      # 41 572 864 B / sample


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--batch', type=int, default=16, help='samples per GPU')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--math', default=os.environ.get('VXB_MATH', 'auto'), choices=['auto', 'fp32', 'f16x3', 'bf16x3', 'f16f8c'])
    ap.add_argument('--workload', default='single', choices=['single', 'dual', 'crop', 'train'],
                    help='single = BASELINE config 2 geometry at --batch (headline); dual = config 3 (acting + stabilizing '
                         'encoders, low_dim 7 + arm head, on the same observations; value counts agent-passes); '
                         'crop = config 4 (per-sample VLM-crop bounds [B,6]); train = config 5 (training step: forward + '
                         'losses + backward + NCCL gradient all-reduce + LAMB, --batch samples per GPU; value = samples/s)')
    ap.add_argument('--optimizer', default='lamb', choices=['lamb', 'adam'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-gpu-comparator', action='store_true', help='skip the informational torch-eager run of the Q-network on the same GPU')
    return ap.parse_args()


def ncu_traffic(batch, f8c=True):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu capture
    (taken at batch 16 on the kernel that is timed); None for any other batch."""
    name = 'ncu_r02_conv3_f8c_b16_summary.json' if f8c else 'ncu_r01_conv3_umma_b16_summary.json'
    path = os.path.join(ROOT, 'profiles', name)
    if batch != 16 or not os.path.exists(path):
        return None
    d = json.load(open(path))
    d = d[0] if isinstance(d, list) else d
    unit = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
    tot = 0.0
    for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        v, u = d[k].split()
        tot += float(v) * unit[u]
    return tot


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p['hbm_gbs'], tensor=p['bf16_tflops_sustained'], tensor_burst=p['bf16_tflops'],
                    source='measured (MEASURED_PEAKS.json)')
    return dict(hbm=6650.0, tensor=1400.0, tensor_burst=1590.0, source='fallback (B200_PROFILING.md)')


WORKLOADS = {
    'dual': 'batch={b}/GPU, 100^3 voxels, 4 cameras 128x128 RGB-D, CLIP lang tokens [77,512], bimanual acting + stabilizing dual-agent forward '
            '(two encoders, low_dim 7, arm head; 2 agent-passes per sample; bimanual steps/s = value / 2)',
    'crop': 'batch={b}/GPU, 100^3 voxels, 4 cameras 128x128 RGB-D, CLIP lang tokens [77,512], single-arm PerAct forward with per-sample VLM-crop bounds [B,6]',
}
WORKLOAD = 'batch={b}/GPU, 100^3 voxels, 4 cameras 128x128 RGB-D, CLIP lang tokens [77,512], single-arm PerAct forward (voxelize + Q-net, 2048 latents, depth 6)'


def cpu_model():
    try:
        for l in open('/proc/cpuinfo'):
            if l.startswith('model name'):
                return l.split(':', 1)[1].strip()
    except OSError:
        pass
    import platform
    return platform.processor() or 'unknown'


def gpu_torch_comparator(agent_enc, grid_cl, proprio, lang, B, dev, steps=2):
    """Informational: the reference's Q-network as plain PyTorch modules' functional restatement (oracle.qnet_oracle.qnet_forward,
    cuDNN / cuBLAS eager, fp32) on the SAME GPU and the same voxel grid, with TF32 off and on."""
    import torch
    from oracle import qnet_oracle
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import util
    sd = {k: v.detach().to(dev) for k, v in agent_enc.state_dict().items()}
    cfg = dict(voxel_patch_size=5, voxel_patch_stride=5, depth=6, iterations=1, cross_heads=1, latent_heads=8, activation='lrelu',
               num_collision_classes=2, arm_pred_loss=False, no_language=False)
    out = {}
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            b = B
            while True:
                try:
                    g = grid_cl[:b]
                    with torch.no_grad():
                        qnet_oracle.qnet_forward(sd, cfg, g, proprio[:b], lang[:b])          # warm-up (cuDNN autotune, allocator)
                        torch.cuda.synchronize()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        for _ in range(steps):
                            qnet_oracle.qnet_forward(sd, cfg, g, proprio[:b], lang[:b])
                        e1.record()
                        torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / steps
                    out['tf32_on' if tf32 else 'tf32_off'] = {'passes_per_s': b / (ms * 1e-3), 'ms_per_step': ms, 'batch': b}
                    break
                except torch.cuda.OutOfMemoryError:
                    torch.cuda.empty_cache()
                    if b == 1:
                        out['tf32_on' if tf32 else 'tf32_off'] = {'error': 'out of memory at batch 1'}
                        break
                    b //= 2
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
        torch.cuda.empty_cache()
    out['what'] = ('oracle.qnet_oracle.qnet_forward (torch eager fp32: cuDNN conv3d, cuBLAS matmul, materialised softmax) on this GPU, '
                   'Q-network only (voxel grid given), %d timed steps after 1 warm-up' % steps)
    return out


def cpu_pass_rate(steps, warmup, threads=None):
    """The reference's CPU PyTorch path (oracle port) on one sample: passes/s on the host cores."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
    from oracle import qnet_oracle, voxel_oracle
    import make_golden
    import util
    torch.set_num_threads(threads or os.cpu_count())
    c = dict(make_golden.QNET_CASES['qnet_v100_b1'])
    obs, enc, sd = util.make_case(c)
    cfg = util.oracle_cfg(c)

    def one():
        with torch.no_grad():
            return qnet_oracle.qfunction_forward(sd, cfg, voxel_oracle.voxelize, obs['rgb'], obs['pcd'],
                                                 obs['proprio'], obs['lang_token_embs'], obs['bounds'], c['V'])
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return steps / dt, torch.get_num_threads(), dt / steps


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = max(1, args.steps)
    rate, cores, spp = cpu_pass_rate(steps, max(1, min(args.warmup, 2)))
    line = {
        'impl': 'reference', 'metric': 'policy fwd passes/sec at 100^3 voxels', 'value': rate, 'unit': 'passes/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': max(1, min(args.warmup, 2)), 'ms_per_step': spp * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD.format(b=args.batch) + ' -- reference arm: 1 sample of that batch per step',
                   'note': 'reference CPU PyTorch path (oracle port) on the host cores; each step = 1 sample of the batch'},
        'cpu_baseline': {'value': rate, 'unit': 'passes/s', 'cores': cores, 'kind': 'port', 'cpu_model': cpu_model(),
                         'torch_threads': cores,
                         'sample': 'batch 1 of the workload per step (voxelize + Q-net forward), torch CPU fp32'},
        'e2e': {'value': rate, 'unit': 'passes/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


class ClockSampler:
    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + q, '--format=csv,noheader,nounits',
                                       '-lms', '100'], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(', ') for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.strip().lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from voxactb_b200 import QFunction, VoxelGrid, PerceiverVoxelLangEncoder, _lib, synth
    from voxactb_b200 import distributed as vdist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # NCCL prints its version banner to fd 1 while the communicator is created: keep stdout to the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    L = _lib.lib()
    _lib.check(L.vxb_check_device(), 'vxb_check_device')
    math_mode = {'fp32': _lib.MATH_FP32_SIMT, 'f16x3': _lib.MATH_F16X3, 'bf16x3': _lib.MATH_F16X3, 'f16f8c': _lib.MATH_F16F8C,
                 'auto': PerceiverVoxelLangEncoder.math_mode}[args.math]
    B, V = args.batch, 100
    torch.manual_seed(1234 + rank)
    dual = args.workload == 'dual'
    low_dim = 7 if dual else 4
    workload = WORKLOADS.get(args.workload, WORKLOAD).format(b=B)

    def make_agent(seed):
        enc = PerceiverVoxelLangEncoder(
            depth=6, iterations=1, voxel_size=V, initial_dim=10, low_dim_size=low_dim, layer=0, num_rotation_classes=72,
            num_grip_classes=2, num_collision_classes=2, input_axis=3, num_latents=2048, latent_dim=512, cross_heads=1,
            latent_heads=8, cross_dim_head=64, latent_dim_head=64, activation='lrelu', weight_tie_layers=False,
            pos_encoding_with_lang=True, input_dropout=0.1, attn_dropout=0.1, decoder_dropout=0.0,
            lang_fusion_type='seq', voxel_patch_size=5, voxel_patch_stride=5, final_dim=64,
            **({'arm_pred_loss': True} if dual else {})).eval()
        enc.load_state_dict(synth.random_state_dict(enc, seed), strict=False)
        enc.math_mode = math_mode
        vg = VoxelGrid(synth.SCENE_BOUNDS, V, dev, B, 3, 4 * 128 * 128)
        return enc, vg, QFunction(enc, vg, 0.15, 5, dev, False, dual).to(dev).eval()

    enc, vg, q = make_agent(2234)                      # single-arm / acting agent
    agents = [q]
    if dual:
        agents.append(make_agent(3234)[2])             # stabilizing agent: its own weights, same observations
    passes_per_sample = len(agents)

    obs = synth.make_observation(1234 + rank, B, 4, 128, 128, low_dim=low_dim, per_sample_crop=args.workload == 'crop')
    host = {k: [t.pin_memory() for t in obs[k]] for k in ('rgb', 'pcd')}
    for k in ('proprio', 'lang_goal_emb', 'lang_token_embs', 'bounds'):
        host[k] = obs[k].pin_memory()
    h2d_bytes = sum(t.numel() * 4 for k in ('rgb', 'pcd') for t in host[k]) + sum(
        host[k].numel() * 4 for k in ('proprio', 'lang_token_embs', 'bounds'))

    def upload():
        d = {k: [t.to(dev, non_blocking=True) for t in host[k]] for k in ('rgb', 'pcd')}
        for k in ('proprio', 'lang_token_embs', 'bounds'):
            d[k] = host[k].to(dev, non_blocking=True)
        return d

    resident = upload()
    resident['lang_goal_emb'] = host['lang_goal_emb'].to(dev)
    torch.cuda.synchronize()

    def step_device(d):
        rgb_pcd = [[r, p] for r, p in zip(d['rgb'], d['pcd'])]
        return [a(rgb_pcd, d['proprio'], d['pcd'], resident['lang_goal_emb'], d['lang_token_embs'], d['bounds'], None, None)
                for a in agents]

    out_hosts = [{'coords': torch.empty(B, 3, dtype=torch.int32).pin_memory(),
                  'rg': torch.empty(B, 4, dtype=torch.int32).pin_memory(),
                  'coll': torch.empty(B, dtype=torch.int32).pin_memory(),
                  'xyz': torch.empty(B, 3, dtype=torch.float32).pin_memory()} for _ in agents]
    d2h_bytes = sum(t.numel() * 4 for oh in out_hosts for t in oh.values())

    def step_e2e():
        d = upload()
        for a, out, out_host in zip(agents, step_device(d), out_hosts):
            trans, rot_grip, coll = out[0], out[1], out[2]
            coords, rg, ic, xyz = a.select_action(trans, rot_grip, coll, d['bounds'])
            out_host['coords'].copy_(coords, non_blocking=True)
            out_host['rg'].copy_(rg, non_blocking=True)
            out_host['coll'].copy_(ic, non_blocking=True)
            out_host['xyz'].copy_(xyz, non_blocking=True)
        torch.cuda.current_stream().synchronize()   # the caller needs the actions before the next step

    # end-to-end with the NEXT step's host->device copies on a copy stream (voxactb_b200.act.HostStager): the copies of step
    # i+1 overlap the Q-network of step i; every step still uploads its own inputs and reads its own actions back
    from voxactb_b200.act import HostStager
    stager = HostStager(dev, slots=2)
    host_in = {k: host[k] for k in ('rgb', 'pcd', 'proprio', 'lang_token_embs', 'bounds')}
    stager.stage(host_in)

    def step_e2e_pipelined():
        d = stager.acquire()
        outs = step_device(d)                        # this step's kernels are enqueued first ...
        stager.stage(host_in)                        # ... then the next step's uploads (copy stream): they overlap the compute
        for a, out, out_host in zip(agents, outs, out_hosts):
            trans, rot_grip, coll = out[0], out[1], out[2]
            coords, rg, ic, xyz = a.select_action(trans, rot_grip, coll, d['bounds'])
            out_host['coords'].copy_(coords, non_blocking=True)
            out_host['rg'].copy_(rg, non_blocking=True)
            out_host['coll'].copy_(ic, non_blocking=True)
            out_host['xyz'].copy_(xyz, non_blocking=True)
        stager.release()
        torch.cuda.current_stream().synchronize()

    # the same loop through the CUDA-graph actor (voxactb_b200.act.GraphedActor.act: the public closed-loop call): one graph
    # launch per agent instead of ~120 kernel launches behind the per-step synchronisation
    from voxactb_b200.act import GraphedActor
    gactors = [GraphedActor(a, 5) for a in agents]

    def step_e2e_graphed():
        d = stager.acquire()
        stager.stage(host_in)
        rgb_pcd = [[r, p] for r, p in zip(d['rgb'], d['pcd'])]
        for ga in gactors:
            ga.act(rgb_pcd, d['proprio'], d['pcd'], resident['lang_goal_emb'], d['lang_token_embs'], d['bounds'])
        stager.release()

    def timed(fn, steps, warmup, profile=False):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if profile:
            L.vxb_profile_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        stages = None
        if profile:
            buf = (ctypes.c_double * L.vxb_profile_stage_count())()
            n = L.vxb_profile_read(buf)
            L.vxb_profile_enable(0)
            stages = {L.vxb_profile_stage_name(i).decode(): buf[i] / max(n, 1) for i in range(len(buf))}
        ms = vdist.max_over_ranks(ms, dev)          # whole-job time of the step = the slowest rank's device time
        return ms, stages

    # voxelize-only timing (HBM-bound kernel of the path), same stream, CUDA events
    coords_flat, feats_flat = synth.flatten_cameras(obs)
    cf, ff, bb = coords_flat.to(dev), feats_flat.to(dev), resident['bounds']
    clocks = ClockSampler(local)
    ms_dev, stages = timed(lambda: step_device(resident), args.steps, args.warmup, profile=True)
    clk = clocks.stop()
    launches_per_step = passes_per_sample * (L.vxb_voxelize_launches() + enc.last_launch_count)
    vox_calls = 4 * max(args.steps, 10)              # ~50 ms region: a 10 ms one swings by 15 % with the clock state the step leaves behind
    ms_vox, _ = timed(lambda: vg.coords_to_bounding_voxel_grid(cf, ff, bb), vox_calls, 5)
    ms_e2e_serial, _ = timed(step_e2e, args.steps, args.warmup)
    ms_e2e, _ = timed(step_e2e_pipelined, args.steps, args.warmup)
    ms_e2e_graph, _ = timed(step_e2e_graphed, args.steps, args.warmup)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    per_step = ms_dev / args.steps
    value = world * B * passes_per_sample / (per_step * 1e-3)
    e2e_value = world * B * passes_per_sample / (ms_e2e / args.steps * 1e-3)
    final_ms = stages['final_conv']
    achieved_tf = FINAL_CONV_FLOPS * B / (final_ms * 1e-3) / 1e12
    vox_ms = ms_vox / vox_calls
    simt = math_mode == _lib.MATH_FP32_SIMT
    f8c = math_mode == _lib.MATH_F16F8C
    # MMA "units" (fp16-MMA-equivalents of tensor-pipe time) per logical product of the final conv / folded up-conv:
    # 3 = hi*hi + hi*lo + lo*hi in fp16; 2 = fp16 hi*hi + one E4M3 MMA (twice the rate, twice the K) for both corrections
    units = 2 if f8c else 3
    padded = (102 / 100) ** 2 * 1.02            # rows the kernel multiplies / useful rows (replicate halo of the flat plane tiles)
    whole_tf = FLOPS_PER_PASS * B * passes_per_sample / (per_step * 1e-3) / 1e12
    dtype = {_lib.MATH_FP32_SIMT: 'f32',
             _lib.MATH_F16X3: 'f16x3 (split-fp16 hi/lo planes, 3 tcgen05 kind::f16 MMAs per product, fp32 accumulate in TMEM)',
             _lib.MATH_F16F8C: 'f16x3 (split-fp16 hi/lo planes, fp32 accumulate in TMEM); final conv + folded up-conv: fp16 hi*hi '
                               '+ one kind::f8f6f4 E4M3 MMA carrying both 2^-11 correction terms'}[math_mode]
    kernel = {_lib.MATH_FP32_SIMT: 'final conv 3x3x3, 128->64 @100^3 (fp32 FFMA implicit GEMM)',
              _lib.MATH_F16X3: 'conv3_umma_kernel: final conv 3x3x3, 128->64 @100^3 (input-stationary tcgen05, split-fp16 x3)',
              _lib.MATH_F16F8C: 'conv3_f8c_kernel: final conv 3x3x3, 128->64 @100^3 (input-stationary tcgen05, fp16 + E4M3-corrected, '
                                'dz taps merged into N=192 MMAs)'}[math_mode]
    line = {
        'metric': 'policy fwd passes/sec at 100^3 voxels', 'value': value, 'unit': 'passes/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': dtype,
        'data': 'synthetic',
        'config': {'workload': workload, 'global_batch': world * B, 'parallelism': 'batch-sharded x%d, no collective' % world,
                   'math_mode': {_lib.MATH_FP32_SIMT: 'fp32_simt', _lib.MATH_F16X3: 'split16x3_tcgen05',
                                 _lib.MATH_F16F8C: 'split16x3_tcgen05 + f16_fp8c convs'}[math_mode],
                   'parity_gate': 'max|a-b| / max|b| per output tensor < 1e-3 vs the reference goldens (tests/util.py rel_err); '
                                  'voxel indices and arg-max actions bit-exact',
                   'l2': 'no explicit flush: each step streams >10 GB of activations, far above the 126 MB L2'},
        'e2e': {'value': e2e_value, 'unit': 'passes/s', 'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': d2h_bytes,
                'ms_per_step': ms_e2e / args.steps,
                'mode': 'pinned host buffers -> HostStager (the next step\'s H2D copies run on a copy stream while this step computes) -> '
                        'QFunction.forward -> select_action -> D2H of the chosen actions, synchronised every step',
                'cuda_graph_actor_value': world * B * passes_per_sample / (ms_e2e_graph / args.steps * 1e-3),
                'serial_value': world * B * passes_per_sample / (ms_e2e_serial / args.steps * 1e-3),
                'serial_ms_per_step': ms_e2e_serial / args.steps},
        'gpu_launches': (launches_per_step) * args.steps,
        'clocks': clk,
        'roofline': {'kernel': kernel,
                     'bound': 'tensor',
                     'achieved': achieved_tf, 'peak': pk['tensor'], 'unit': 'TFLOP/s', 'frac': achieved_tf / pk['tensor'],
                     'frac_of_burst_peak': achieved_tf / pk['tensor_burst'], 'peak_burst': pk['tensor_burst'],
                     'traffic': ncu_traffic(B, f8c) if not simt else None,
                     'algorithmic_flops_per_launch': FINAL_CONV_FLOPS * B,
                     # what the tensor pipe actually executes, in fp16-MMA-equivalents, on the padded (102/100)^2 x (V+2)/V rows
                     'mma_units_per_product': None if simt else units,
                     'executed_mma_tflops': None if simt else achieved_tf * units * padded,
                     'executed_frac_of_peak': None if simt else achieved_tf * units * padded / pk['tensor'],
                     'note': 'achieved = direct-convolution FLOPs (442.4 GF/sample) / CUDA-event time of the kernel inside the step; the '
                             'kernel spends %d fp16-MMA-equivalents per logical product for fp32-class accuracy, so its own ceiling is '
                             '1/%d of this peak' % (units, units),
                     'peak_source': pk['source'] + ' bf16 sustained (frac) and burst (frac_of_burst_peak)', 'ms_per_launch': final_ms,
                     'whole_forward_tflops': whole_tf,
                     'whole_forward_frac': whole_tf / pk['tensor'], 'whole_forward_frac_of_burst_peak': whole_tf / pk['tensor_burst'],
                     'voxelize': {'bound': 'hbm', 'achieved': VOXELIZE_BYTES * B / (vox_ms * 1e-3) / 1e9, 'peak': pk['hbm'],
                                  'unit': 'GB/s', 'frac': VOXELIZE_BYTES * B / (vox_ms * 1e-3) / 1e9 / pk['hbm'],
                                  'ms_per_call': vox_ms}},
        'stages_ms': stages,
    }
    if world == 1 and B <= 2 and args.workload != 'crop':
        # closed-loop acting latency (eval.py runs batch 1; the dual workload alternates the two agents as
        # rollout_generator does): observations arrive on the host, the 9-D action is needed on the host
        from voxactb_b200.act import GraphedActor
        actors = [GraphedActor(a, 5) for a in agents]

        def step_act():
            d = upload()
            rgb_pcd = [[r, p] for r, p in zip(d['rgb'], d['pcd'])]
            for ga in actors:
                ga.act(rgb_pcd, d['proprio'], d['pcd'], resident['lang_goal_emb'], d['lang_token_embs'], d['bounds'])
        ms_act, _ = timed(step_act, max(args.steps, 20), max(args.warmup, 3))
        ms_act /= max(args.steps, 20)
        line['act_latency'] = {'ms_per_step': ms_act, 'agent_calls_per_step': len(actors), 'batch': B,
                               'ungraphed_ms_per_step': ms_e2e_serial / args.steps,
                               'mode': 'H2D of the observation -> GraphedActor.act (CUDA-graph replay of voxelize + Q-net + select + act '
                                       'tail) -> 9-D action on the host, per agent'}
    if world == 1 and not args.no_cpu_baseline:
        rate, cores, spp = cpu_pass_rate(8, 1)
        line['cpu_baseline'] = {'value': rate, 'unit': 'passes/s', 'cores': cores, 'kind': 'port', 'cpu_model': cpu_model(),
                                'torch_threads': cores,
                                'sample': '8 timed single-sample passes (batch 1 of the workload) after 1 warm-up, torch CPU fp32 oracle port'}
    if world == 1 and not args.no_gpu_comparator and not dual:
        with torch.no_grad():
            grid_cl = vg.coords_to_bounding_voxel_grid(cf, ff, bb).permute(0, 4, 1, 2, 3).contiguous()
        line['gpu_torch_comparator'] = gpu_torch_comparator(enc, grid_cl, resident['proprio'], resident['lang_token_embs'], B, dev)
        del grid_cl
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


TRAIN_FLOPS_PER_SAMPLE = 3 * FLOPS_PER_PASS      # forward + dgrad + wgrad (~5.0e12 / sample)


def cpu_train_rate(steps):
    """The reference's CPU training step (oracle port: torch CPU forward + autograd + LAMB) on one sample."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
    from oracle import train_oracle
    import make_golden
    import util
    torch.set_num_threads(os.cpu_count())
    c = dict(make_golden.QNET_CASES['qnet_v100_b1'])
    obs, enc, sd = util.make_case(c)
    sd = {k: v for k, v in sd.items() if not k.endswith(('pos_x', 'pos_y', 'pos_z'))}
    lab = make_golden.train_labels(c)
    t0 = time.perf_counter()
    for _ in range(steps):
        train_oracle.training_step(sd, util.oracle_cfg(c), obs['rgb'], obs['pcd'], obs['proprio'], obs['lang_token_embs'],
                                   obs['bounds'], c['V'], lab)
    dt = time.perf_counter() - t0
    return steps / dt, torch.get_num_threads()


def run_train(args):
    """BASELINE config 5: training step = voxelize + forward (train mode, dropout 0.1) + CE losses + backward + NCCL gradient
    all-reduce (DDP semantics) + LAMB, --batch samples per GPU (weak scaling).  value = samples/s over all ranks."""
    import torch
    import torch.distributed as dist
    from voxactb_b200 import QFunction, VoxelGrid, PerceiverVoxelLangEncoder, _lib, synth, train

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    L = _lib.lib()
    _lib.check(L.vxb_check_device(), 'vxb_check_device')
    math_mode = {'fp32': _lib.MATH_FP32_SIMT, 'f16x3': _lib.MATH_F16X3, 'bf16x3': _lib.MATH_F16X3, 'f16f8c': _lib.MATH_F16F8C,
                 'auto': PerceiverVoxelLangEncoder.math_mode}[args.math]
    B, V = args.batch, 100
    torch.manual_seed(4321)                                # identical initial weights on every rank (DDP broadcast equivalent)
    enc = PerceiverVoxelLangEncoder(
        depth=6, iterations=1, voxel_size=V, initial_dim=10, low_dim_size=4, layer=0, num_rotation_classes=72,
        num_grip_classes=2, num_collision_classes=2, input_axis=3, num_latents=2048, latent_dim=512, cross_heads=1,
        latent_heads=8, cross_dim_head=64, latent_dim_head=64, activation='lrelu', weight_tie_layers=False,
        pos_encoding_with_lang=True, input_dropout=0.1, attn_dropout=0.1, decoder_dropout=0.0, lang_fusion_type='seq',
        voxel_patch_size=5, voxel_patch_stride=5, final_dim=64)
    enc.load_state_dict(synth.random_state_dict(enc, 2234), strict=False)
    enc.math_mode = math_mode
    vg = VoxelGrid(synth.SCENE_BOUNDS, V, dev, B, 3, 4 * 128 * 128)
    q = QFunction(enc, vg, 0.15, 5, dev, True, False).to(dev).train(True)
    tr = train.PerActTrainer(q, lr=5e-4, weight_decay=1e-6, optimizer=args.optimizer)
    obs = synth.make_observation(1234 + rank, B, 4, 128, 128, low_dim=4)
    g = torch.Generator().manual_seed(99 + rank)
    labels = {'trans': torch.randint(0, V, (B, 3), generator=g, dtype=torch.int32).to(dev),
              'rot_grip': torch.cat([torch.randint(0, 72, (B, 3), generator=g, dtype=torch.int32),
                                     torch.randint(0, 2, (B, 1), generator=g, dtype=torch.int32)], 1).to(dev),
              'collision': torch.randint(0, 2, (B, 1), generator=g, dtype=torch.int32).to(dev)}
    host = {k: [t.pin_memory() for t in obs[k]] for k in ('rgb', 'pcd')}
    for k in ('proprio', 'lang_goal_emb', 'lang_token_embs', 'bounds'):
        host[k] = obs[k].pin_memory()
    h2d_bytes = sum(t.numel() * 4 for k in ('rgb', 'pcd') for t in host[k]) + sum(
        host[k].numel() * 4 for k in ('proprio', 'lang_token_embs', 'bounds')) + sum(v.numel() * 4 for v in labels.values())

    def upload():
        d = {k: [t.to(dev, non_blocking=True) for t in host[k]] for k in ('rgb', 'pcd')}
        for k in ('proprio', 'lang_goal_emb', 'lang_token_embs', 'bounds'):
            d[k] = host[k].to(dev, non_blocking=True)
        return d

    resident = upload()
    torch.cuda.synchronize()
    loss_host = torch.empty(1).pin_memory()
    marks = []

    def step(d, record=False):
        rgb_pcd = [[r, p] for r, p in zip(d['rgb'], d['pcd'])]
        res = tr.update(rgb_pcd, d['proprio'], d['pcd'], d['lang_goal_emb'], d['lang_token_embs'], d['bounds'], labels)
        return res['total_loss']

    def step_e2e():
        loss = step(upload())
        loss_host.copy_(loss.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    clocks = ClockSampler(local)
    ms_dev = timed(lambda: step(resident), args.steps, args.warmup)
    clk = clocks.stop()
    ms_e2e = timed(step_e2e, args.steps, args.warmup)
    # stage split of one step (CUDA events on the launching stream)
    tr.profile = True
    step(resident)
    torch.cuda.synchronize()
    stages = dict(tr.last_stage_ms)
    tr.profile = False
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    per_step = ms_dev / args.steps
    value = world * B / (per_step * 1e-3)
    tf = TRAIN_FLOPS_PER_SAMPLE * B / (per_step * 1e-3) / 1e12
    bwd_ms = stages.get('backward', 0.0)
    line = {
        'metric': 'training samples/sec at 100^3 voxels (fwd + loss + bwd + NCCL grad all-reduce + %s)' % args.optimizer.upper(),
        'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': per_step,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32' if math_mode == _lib.MATH_FP32_SIMT else 'f16x3 (split-fp16 tcgen05: forward incl. fused attention with dropout, conv / linear dgrad + wgrad, attention-backward score products); the three sequence-length products of the attention backward fp32 FFMA',
        'data': 'synthetic',
        'config': {'workload': 'BASELINE config 5: batch=%d/GPU, 100^3 voxels, 4 cameras 128x128 RGB-D, training step (train-mode '
                               'dropout 0.1, CE losses, backward, gradient all-reduce over NCCL, %s)' % (B, args.optimizer.upper()),
                   'global_batch': world * B, 'parallelism': 'data-parallel x%d, one fp32 gradient all-reduce (133 MB) per step' % world,
                   'l2': 'no explicit flush: each step streams >50 GB of activations and gradients'},
        'e2e': {'value': world * B / (ms_e2e / args.steps * 1e-3), 'unit': 'samples/s', 'h2d_bytes_per_step': h2d_bytes,
                'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': TRAIN_LAUNCHES_PER_STEP * args.steps,
        'gpu_launches_source': 'launches per step counted from the ncu launch list of this step (profiles/launches_r02_train.csv: '
                               '1 205 kernels, the library\'s own except ~270 torch fill / copy kernels of the host glue)',
        'clocks': clk,
        'roofline': {'kernel': 'whole training step (forward + dgrad + wgrad ~ 3 x forward FLOPs)', 'bound': 'tensor',
                     'achieved': tf, 'peak': pk['tensor'], 'unit': 'TFLOP/s', 'frac': tf / pk['tensor'], 'traffic': None,
                     'algorithmic_flops_per_step': TRAIN_FLOPS_PER_SAMPLE * B, 'peak_source': pk['source'] + ' bf16 sustained',
                     'backward_ms': bwd_ms},
        'stages_ms': stages,
    }
    if world == 1 and not args.no_cpu_baseline:
        rate, cores = cpu_train_rate(1)
        line['cpu_baseline'] = {'value': rate, 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
                                'sample': '1 single-sample training step (forward + autograd + LAMB) of the torch CPU fp32 oracle port'}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    elif args.workload == 'train':
        run_train(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
