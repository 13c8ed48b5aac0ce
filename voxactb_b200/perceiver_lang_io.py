"""B200-native ``PerceiverVoxelLangEncoder`` -- host-side mirror of the reference interface.

Same constructor keywords/defaults, ``forward`` signature, return tuple and ``state_dict`` key
names/shapes as reference peract/agents/peract_bc/perceiver_lang_io.py:136-485, so a checkpoint
written by either implementation loads into the other (agent ``save_weights``/``load_weights``,
qattention_peract_bc_agent.py:826-880).  The module owns the parameters only; all arithmetic runs
in libvoxactb.so (hand-written sm_100a CUDA) through ``vxb_qnet_prepare`` / ``vxb_qnet_forward_f32``.
There is no PyTorch or CPU fallback: a CPU tensor or a missing library raises.
"""
import ctypes
import math

import numpy as np
import torch
from torch import nn

from . import _lib

LRELU_SLOPE = 0.02  # reference helpers/network_utils.py:12

_FIXED_SLOTS = [
    'pos_encoding', 'latents',
    'input_preprocess.conv3d.weight', 'input_preprocess.conv3d.bias',
    'patchify.conv3d.weight', 'patchify.conv3d.bias',
    'lang_preprocess.weight', 'lang_preprocess.bias',
    'proprio_preprocess.linear.weight', 'proprio_preprocess.linear.bias',
    None, None,  # 2-robot left-arm proprio
    'cross_attend_blocks.0.norm.weight', 'cross_attend_blocks.0.norm.bias',
    'cross_attend_blocks.0.norm_context.weight', 'cross_attend_blocks.0.norm_context.bias',
    'cross_attend_blocks.0.fn.to_q.weight', 'cross_attend_blocks.0.fn.to_kv.weight',
    'cross_attend_blocks.0.fn.to_out.weight', 'cross_attend_blocks.0.fn.to_out.bias',
    'cross_attend_blocks.1.norm.weight', 'cross_attend_blocks.1.norm.bias',
    'cross_attend_blocks.1.fn.net.0.weight', 'cross_attend_blocks.1.fn.net.0.bias',
    'cross_attend_blocks.1.fn.net.2.weight', 'cross_attend_blocks.1.fn.net.2.bias',
    'decoder_cross_attn.norm.weight', 'decoder_cross_attn.norm.bias',
    'decoder_cross_attn.norm_context.weight', 'decoder_cross_attn.norm_context.bias',
    'decoder_cross_attn.fn.to_q.weight', 'decoder_cross_attn.fn.to_kv.weight',
    'decoder_cross_attn.fn.to_out.weight', 'decoder_cross_attn.fn.to_out.bias',
    'up0.conv_up.0.conv3d.weight', 'up0.conv_up.0.conv3d.bias',
    'up0.conv_up.2.conv3d.weight', 'up0.conv_up.2.conv3d.bias',
    'final.conv3d.weight', 'final.conv3d.bias',
    'trans_decoder.conv3d.weight', 'trans_decoder.conv3d.bias',
    None, None,  # 2-robot left-arm trans decoder
    'dense0.linear.weight', 'dense0.linear.bias',
    'dense1.linear.weight', 'dense1.linear.bias',
    'rot_grip_collision_ff.linear.weight', 'rot_grip_collision_ff.linear.bias',
    'dense2.linear.weight', 'dense2.linear.bias',
    'arm_ff.linear.weight', 'arm_ff.linear.bias',
    None, None,
]
# 2-robot encoder: the slots that change meaning (include/voxactb.h, enum vxb_param_slot)
_TWO_ROBOT_SLOTS = {
    10: 'proprio_preprocess.linear.weight', 11: 'proprio_preprocess.linear.bias',   # same block for both arms
    42: 'trans_decoder_left_arm.conv3d.weight', 43: 'trans_decoder_left_arm.conv3d.bias',
    50: 'dense0_left_arm.linear.weight', 51: 'dense0_left_arm.linear.bias',
    52: 'rot_grip_collision_ff_left_arm.linear.weight', 53: 'rot_grip_collision_ff_left_arm.linear.bias',
    54: 'dense1_left_arm.linear.weight', 55: 'dense1_left_arm.linear.bias',
}
_LAYER_SLOTS = ['0.norm.weight', '0.norm.bias', '0.fn.to_q.weight', '0.fn.to_kv.weight',
                '0.fn.to_out.weight', '0.fn.to_out.bias', '1.norm.weight', '1.norm.bias',
                '1.fn.net.0.weight', '1.fn.net.0.bias', '1.fn.net.2.weight', '1.fn.net.2.bias']


class _Node(nn.Module):
    """Parameter container: gives parameters the reference's dotted state-dict names."""

    def child(self, name):
        if name not in self._modules:
            self.add_module(name, _Node())
        return self._modules[name]


def _register(root, dotted, tensor, buffer=False):
    parts = dotted.split('.')
    node = root
    for p in parts[:-1]:
        node = node.child(p) if isinstance(node, _Node) else _child_of(node, p)
    if buffer:
        node.register_buffer(parts[-1], tensor)
    else:
        node.register_parameter(parts[-1], nn.Parameter(tensor))


def _child_of(module, name):
    if name not in module._modules:
        module.add_module(name, _Node())
    return module._modules[name]


def _uniform(shape, bound):
    return torch.empty(shape).uniform_(-bound, bound)


def _init_weight(shape, fan_in, fan_out, activation):
    """Initialisers the reference blocks use (network_utils.py:140-156, 263-276): Kaiming-uniform
    for relu/lrelu, Xavier-uniform for linear outputs."""
    if activation == 'lrelu':
        gain = math.sqrt(2.0 / (1 + LRELU_SLOPE ** 2))
        return _uniform(shape, gain * math.sqrt(3.0 / fan_in))
    if activation == 'relu':
        return _uniform(shape, math.sqrt(2.0) * math.sqrt(3.0 / fan_in))
    if activation is None:
        return _uniform(shape, math.sqrt(6.0 / (fan_in + fan_out)))
    raise ValueError('%s not recognized.' % activation)


def _spatial_positions(n):
    """SpatialSoftmax3D buffers pos_x/pos_y/pos_z (network_utils.py:782-795); kept so the
    state_dict has the reference's keys -- the kernels regenerate them on the fly."""
    lin = np.linspace(-1., 1., n)
    px, py, pz = np.meshgrid(lin, lin, lin)
    return [torch.from_numpy(a.reshape(-1)).float() for a in (px, py, pz)]


class _QnetTrainFn(torch.autograd.Function):
    """Training-mode forward/backward of the encoder as ONE autograd node: `total_loss.backward()` in the reference's
    agent.update (qattention_peract_bc_agent.py:581) reaches vxb_qnet_backward_f32 through it, and the parameter
    gradients come back as ordinary `.grad` tensors (so torch optimizers, the reference's Lamb and DDP hooks all work)."""

    @staticmethod
    def forward(ctx, enc, grid, proprio, lang, *params):
        outs, gen = enc._forward_train(grid, proprio, lang)
        ctx.enc, ctx.gen = enc, gen
        ctx.inputs = (grid, proprio, lang)
        ctx.n_params = len(params)
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        grads = ctx.enc._backward_train(ctx.gen, ctx.inputs, gouts)
        return (None, None, None, None) + tuple(grads)


class PerceiverVoxelLangEncoder(nn.Module):
    """Drop-in for reference perceiver_lang_io.py:136 (same keywords and defaults)."""

    # arithmetic of the dense contractions: tcgen05 split-16-bit products with fp32-class accuracy.
    # MATH_F16F8C (default): three fp16 MMAs per product in the transformer, fp16 hi*hi + one E4M3 correction MMA in the two
    # large convolutions; MATH_F16X3: three fp16 MMAs everywhere; MATH_FP32_SIMT: fp32 FFMA everywhere (slow reference mode)
    math_mode = _lib.MATH_F16F8C
    TWO_ROBOTS = False     # PerceiverVoxelLang2RobotsEncoder: two proprio streams (C = 3 * im_channels), two head sets

    def __init__(self, depth, iterations, voxel_size, initial_dim, low_dim_size, layer=0,
                 num_rotation_classes=72, num_grip_classes=2, num_collision_classes=2, input_axis=3,
                 num_latents=512, im_channels=64, latent_dim=512, cross_heads=1, latent_heads=8,
                 cross_dim_head=64, latent_dim_head=64, activation='relu', weight_tie_layers=False,
                 pos_encoding_with_lang=True, input_dropout=0.1, attn_dropout=0.1,
                 decoder_dropout=0.0, lang_fusion_type='seq', voxel_patch_size=9,
                 voxel_patch_stride=8, no_skip_connection=False, no_perceiver=False,
                 no_language=False, final_dim=64, arm_pred_loss=False):
        super().__init__()
        self.depth = depth
        self.layer = layer
        self.init_dim = int(initial_dim)
        self.iterations = iterations
        self.input_axis = input_axis
        self.voxel_size = voxel_size
        self.low_dim_size = low_dim_size
        self.im_channels = im_channels
        self.pos_encoding_with_lang = pos_encoding_with_lang
        self.lang_fusion_type = lang_fusion_type
        self.voxel_patch_size = voxel_patch_size
        self.voxel_patch_stride = voxel_patch_stride
        self.num_rotation_classes = num_rotation_classes
        self.num_grip_classes = num_grip_classes
        self.num_collision_classes = num_collision_classes
        self.final_dim = final_dim
        self.input_dropout = input_dropout
        self.attn_dropout = attn_dropout
        self.decoder_dropout = decoder_dropout
        self.no_skip_connection = no_skip_connection
        self.no_perceiver = no_perceiver
        self.no_language = no_language
        self.arm_pred_loss = arm_pred_loss
        self.activation = activation
        self.num_latents = num_latents
        self.latent_dim = latent_dim
        self.cross_heads, self.cross_dim_head = cross_heads, cross_dim_head
        self.latent_heads, self.latent_dim_head = latent_heads, latent_dim_head

        unsupported = []
        if lang_fusion_type != 'seq':
            unsupported.append("lang_fusion_type=%r" % lang_fusion_type)
        if activation not in ('lrelu', 'relu'):
            unsupported.append('activation=%r' % activation)
        if num_rotation_classes <= 0:
            unsupported.append('num_rotation_classes=0 (non-final C2FARM layers)')
        if low_dim_size <= 0:
            unsupported.append('low_dim_size=0')
        if input_axis != 3:
            unsupported.append('input_axis=%r' % input_axis)
        if unsupported:
            raise NotImplementedError(
                'voxactb_b200.PerceiverVoxelLangEncoder builds the configuration launch_utils.create_agent '
                'uses (seq language fusion, positional encoding with language, skip connection); '
                'not built: ' + ', '.join(unsupported))

        spatial = voxel_size // voxel_patch_stride
        self.input_dim_before_seq = C = im_channels * (3 if self.TWO_ROBOTS else 2)
        if self.TWO_ROBOTS and arm_pred_loss:
            raise NotImplementedError('the 2-robot encoder has no arm-prediction head')
        k, D, L = voxel_patch_size, latent_dim, num_latents
        act = activation
        reg = lambda name, t: _register(self, name, t)
        if pos_encoding_with_lang:
            reg('pos_encoding', torch.randn(1, 77 + spatial ** 3, C))                       # reference :196-198
        else:
            # reference :192-194: the encoding covers the voxel tokens only (added before the language tokens are prepended)
            reg('pos_encoding', torch.randn(1, spatial, spatial, spatial, C))
        reg('input_preprocess.conv3d.weight', _init_weight((im_channels, self.init_dim, 1, 1, 1), self.init_dim, im_channels, act))
        reg('input_preprocess.conv3d.bias', torch.zeros(im_channels))
        reg('patchify.conv3d.weight', _init_weight((im_channels, im_channels, k, k, k), im_channels * k ** 3, im_channels * k ** 3, act))
        reg('patchify.conv3d.bias', torch.zeros(im_channels))
        b = 1.0 / math.sqrt(512)
        reg('lang_preprocess.weight', _uniform((C, 512), b))
        reg('lang_preprocess.bias', _uniform((C,), b))
        reg('proprio_preprocess.linear.weight', _init_weight((im_channels, low_dim_size), low_dim_size, im_channels, act))
        reg('proprio_preprocess.linear.bias', torch.zeros(im_channels))
        for name, n in (('ss0', voxel_size),):
            for ax, t in zip('xyz', _spatial_positions(n)):
                _register(self, '%s.pos_%s' % (name, ax), t, buffer=True)
        reg('latents', torch.randn(L, D))

        def attention(prefix, qdim, cdim, heads, dh):
            inner = heads * dh
            reg(prefix + '.fn.to_q.weight', _uniform((inner, qdim), 1 / math.sqrt(qdim)))
            reg(prefix + '.fn.to_kv.weight', _uniform((2 * inner, cdim), 1 / math.sqrt(cdim)))
            reg(prefix + '.fn.to_out.weight', _uniform((qdim, inner), 1 / math.sqrt(inner)))
            reg(prefix + '.fn.to_out.bias', _uniform((qdim,), 1 / math.sqrt(inner)))
            reg(prefix + '.norm.weight', torch.ones(qdim))
            reg(prefix + '.norm.bias', torch.zeros(qdim))

        def feedforward(prefix, dim):
            reg(prefix + '.fn.net.0.weight', _uniform((dim * 8, dim), 1 / math.sqrt(dim)))
            reg(prefix + '.fn.net.0.bias', _uniform((dim * 8,), 1 / math.sqrt(dim)))
            reg(prefix + '.fn.net.2.weight', _uniform((dim, dim * 4), 1 / math.sqrt(dim * 4)))
            reg(prefix + '.fn.net.2.bias', _uniform((dim,), 1 / math.sqrt(dim * 4)))
            reg(prefix + '.norm.weight', torch.ones(dim))
            reg(prefix + '.norm.bias', torch.zeros(dim))

        attention('cross_attend_blocks.0', D, C, cross_heads, cross_dim_head)
        reg('cross_attend_blocks.0.norm_context.weight', torch.ones(C))
        reg('cross_attend_blocks.0.norm_context.bias', torch.zeros(C))
        feedforward('cross_attend_blocks.1', D)
        self.weight_tie_layers = bool(weight_tie_layers)
        for i in range(depth):
            if weight_tie_layers and i > 0:
                # reference perceiver_lang_io.py:263-276 (cache_fn): every latent layer IS the first layer's modules -- the
                # state dict carries layers.<i>.* keys for all i, aliasing the same tensors
                self._modules['layers'].add_module(str(i), self._modules['layers']._modules['0'])
                continue
            attention('layers.%d.0' % i, D, D, latent_heads, latent_dim_head)
            feedforward('layers.%d.1' % i, D)
        attention('decoder_cross_attn', C, D, cross_heads, cross_dim_head)
        reg('decoder_cross_attn.norm_context.weight', torch.ones(D))
        reg('decoder_cross_attn.norm_context.bias', torch.zeros(D))
        reg('up0.conv_up.0.conv3d.weight', _init_weight((final_dim, C, k, k, k), C * k ** 3, final_dim * k ** 3, act))
        reg('up0.conv_up.0.conv3d.bias', torch.zeros(final_dim))
        reg('up0.conv_up.2.conv3d.weight', _init_weight((final_dim, final_dim, k, k, k), final_dim * k ** 3, final_dim * k ** 3, act))
        reg('up0.conv_up.2.conv3d.bias', torch.zeros(final_dim))
        for ax, t in zip('xyz', _spatial_positions(spatial)):
            _register(self, 'ss1.pos_%s' % ax, t, buffer=True)
        # reference :296-306: the skip-less / perceiver-less ablations feed the final convolution 64 channels instead of 128
        fin_in = im_channels if (no_skip_connection or no_perceiver) else im_channels * 2
        reg('final.conv3d.weight', _init_weight((im_channels, fin_in, 3, 3, 3), fin_in * 27, im_channels * 27, act))
        reg('final.conv3d.bias', torch.zeros(im_channels))
        reg('trans_decoder.conv3d.weight', _init_weight((1, final_dim, 3, 3, 3), final_dim * 27, 27, None))
        reg('trans_decoder.conv3d.bias', torch.zeros(1))
        for ax, t in zip('xyz', _spatial_positions(voxel_size)):
            _register(self, 'ss_final.pos_%s' % ax, t, buffer=True)
        flat = im_channels * 4 + C * 4 + im_channels * 4
        nout = num_rotation_classes * 3 + num_grip_classes + num_collision_classes
        reg('dense0.linear.weight', _init_weight((256, flat), flat, 256, act))
        reg('dense0.linear.bias', torch.zeros(256))
        reg('dense1.linear.weight', _init_weight((final_dim, 256), 256, final_dim, act))
        reg('dense1.linear.bias', torch.zeros(final_dim))
        reg('rot_grip_collision_ff.linear.weight', _init_weight((nout, final_dim), final_dim, nout, None))
        reg('rot_grip_collision_ff.linear.bias', torch.zeros(nout))
        if self.TWO_ROBOTS:
            # left-arm head set (reference perceiver_lang_io.py:679-692)
            reg('trans_decoder_left_arm.conv3d.weight', _init_weight((1, final_dim, 3, 3, 3), final_dim * 27, 27, None))
            reg('trans_decoder_left_arm.conv3d.bias', torch.zeros(1))
            for ax, t in zip('xyz', _spatial_positions(voxel_size)):
                _register(self, 'ss_final_left_arm.pos_%s' % ax, t, buffer=True)
            reg('dense0_left_arm.linear.weight', _init_weight((256, flat), flat, 256, act))
            reg('dense0_left_arm.linear.bias', torch.zeros(256))
            reg('dense1_left_arm.linear.weight', _init_weight((final_dim, 256), 256, final_dim, act))
            reg('dense1_left_arm.linear.bias', torch.zeros(final_dim))
            reg('rot_grip_collision_ff_left_arm.linear.weight', _init_weight((nout, final_dim), final_dim, nout, None))
            reg('rot_grip_collision_ff_left_arm.linear.bias', torch.zeros(nout))
        if arm_pred_loss:
            reg('dense2.linear.weight', _init_weight((final_dim, flat), flat, final_dim, act))
            reg('dense2.linear.bias', torch.zeros(final_dim))
            reg('arm_ff.linear.weight', _init_weight((2, final_dim), final_dim, 2, None))
            reg('arm_ff.linear.bias', torch.zeros(2))

        # runtime state (never pickled / deep-copied with live CUDA handles: rebuilt lazily)
        self._prepared = None
        self._prepared_key = None
        self._pos_table = None
        self._pos_table_key = None
        self._workspace = None
        self._train_ws = None
        self._train_gen = 0
        self._train_opts = None
        self.dropout_seed = None      # None: derived from torch.initial_seed(); advanced by one per training forward
        self.last_launch_count = 0

    # ------------------------------------------------------------------ plumbing
    def __getstate__(self):
        state = self.__dict__.copy()
        state['_prepared'] = None
        state['_prepared_key'] = None
        state['_workspace'] = None
        state['_train_ws'] = None
        state['_train_opts'] = None
        return state

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__getstate__().items():
            setattr(new, k, copy.deepcopy(v, memo))
        return new

    def _apply(self, fn, *a, **kw):
        self.invalidate_prepared()
        self._workspace = None
        self._train_ws = None
        return super()._apply(fn, *a, **kw)

    def invalidate_prepared(self):
        """Drop the prepared-weight cache (fp16 hi/lo planes, folded up-conv weights, to_q(LN(latents))).  The cache is
        keyed on each parameter's (data_ptr, _version), which catches in-place updates, load_state_dict and this
        package's fused optimizers; call this after writing through `p.data` (e.g. the reference's lamb.py) in EVAL
        mode -- in training mode the weights are re-prepared on every forward anyway."""
        self._prepared_key = None

    def _desc(self):
        d = _lib.QnetDesc()
        d.struct_bytes = ctypes.sizeof(_lib.QnetDesc)
        d.voxel_size = self.voxel_size
        d.patch_size = self.voxel_patch_size
        d.patch_stride = self.voxel_patch_stride
        d.initial_dim = self.init_dim
        d.im_channels = self.im_channels
        d.low_dim_size = self.low_dim_size
        d.two_robots = int(self.TWO_ROBOTS)
        d.lang_seq_len = 77
        d.lang_emb_dim = 512
        d.num_latents = self.num_latents
        d.latent_dim = self.latent_dim
        d.depth = self.depth
        d.iterations = self.iterations
        d.cross_heads, d.cross_dim_head = self.cross_heads, self.cross_dim_head
        d.latent_heads, d.latent_dim_head = self.latent_heads, self.latent_dim_head
        d.final_dim = self.final_dim
        d.num_rotation_classes = self.num_rotation_classes
        d.num_grip_classes = self.num_grip_classes
        d.num_collision_classes = self.num_collision_classes
        d.arm_pred_loss = int(bool(self.arm_pred_loss))
        d.no_language = int(bool(self.no_language))
        d.act_slope = LRELU_SLOPE if self.activation == 'lrelu' else 0.0
        d.math_mode = int(self.math_mode)
        d.final_input = 1 if self.no_skip_connection else (2 if self.no_perceiver else 0)      # VXB_FINAL_* (reference :456-462)
        return d

    def _param_table(self):
        named = dict(self.named_parameters(remove_duplicate=False))       # weight_tie_layers: layers.<i> alias layers.0
        names = list(_FIXED_SLOTS)
        if self.TWO_ROBOTS:
            names = [_TWO_ROBOT_SLOTS.get(i, n) for i, n in enumerate(names)]
        slots = [named.get(n) if n else None for n in names]
        for i in range(self.depth):
            slots += [named['layers.%d.%s' % (i, s)] for s in _LAYER_SLOTS]
        arr = (ctypes.c_void_p * len(slots))()
        keep = []
        for i, p in enumerate(slots):
            if p is None:
                arr[i] = None
                continue
            t = _lib.f32(p.detach())
            if i < len(names) and names[i] == 'pos_encoding' and not self.pos_encoding_with_lang:
                # the library adds one [77 + T, C] table to the whole sequence: zeros for the language rows.  Cached on the
                # module (a captured CUDA graph keeps reading this buffer) and rebuilt when the parameter changes.
                key = (t.data_ptr(), p._version, t.device)
                if getattr(self, '_pos_table_key', None) != key:
                    self._pos_table = torch.cat([t.new_zeros(77, t.shape[-1]), t.reshape(-1, t.shape[-1])], 0).contiguous()
                    self._pos_table_key = key
                t = self._pos_table
            keep.append(t)
            arr[i] = t.data_ptr()
        return arr, keep, slots

    def _ensure_prepared(self, desc, arr, slots, device):
        key = (device, int(self.math_mode),
               tuple((p.data_ptr(), p._version) for p in slots if p is not None))
        if self._prepared is not None and self._prepared_key == key and not self.training:
            return
        L = _lib.lib()
        nbytes = L.vxb_qnet_prepared_bytes(ctypes.byref(desc))
        if nbytes == 0:
            _lib.check(-2, 'vxb_qnet_prepared_bytes')
        if self._prepared is None or self._prepared.numel() < nbytes or self._prepared.device != device:
            self._prepared = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _lib.check(L.vxb_qnet_prepare(ctypes.byref(desc), arr, _lib.ptr(self._prepared), nbytes,
                                      _lib.stream()), 'vxb_qnet_prepare')
        self._prepared_key = key

    # ------------------------------------------------------------------ forward
    def forward(self, ins, proprio, lang_goal_emb, lang_token_embs, prev_layer_voxel_grid, bounds,
                prev_layer_bounds, mask=None):
        """ins [B,10,V,V,V] (normally the permuted channels-last voxel grid QFunction.forward passes,
        qattention_peract_bc_agent.py:100); returns (trans [B,1,V,V,V], rot_and_grip [B,3R+G],
        collision [B,Cc][, arm [B,2]]) like perceiver_lang_io.py:465-485."""
        out = self._run(ins, proprio, None, lang_token_embs, mask)
        trans, rot_grip, coll, arm = out[0], out[1], out[2], out[6]
        if self.arm_pred_loss:
            return trans, rot_grip, coll, arm
        return trans, rot_grip, coll

    def _run(self, ins, proprio, proprio2, lang_token_embs, mask):
        if mask is not None:
            raise NotImplementedError('attention mask is never passed by the agent (always None)')
        if not ins.is_cuda:
            raise RuntimeError('voxactb_b200.PerceiverVoxelLangEncoder runs on CUDA only (no CPU fallback)')
        B, C10, V = ins.shape[0], ins.shape[1], ins.shape[2]
        if C10 != self.init_dim or V != self.voxel_size:
            raise ValueError('expected ins [B,%d,%d,%d,%d], got %s' % (self.init_dim, self.voxel_size, self.voxel_size, self.voxel_size, tuple(ins.shape)))
        if self.training and torch.is_grad_enabled():
            # training step (row a18): one autograd node around vxb_qnet_forward_train_f32 / vxb_qnet_backward_f32
            if self.TWO_ROBOTS:
                raise NotImplementedError('the 2-robot encoder is inference-only in voxactb_b200 (training: single-arm encoders)')
            if self.no_skip_connection or self.no_perceiver:
                raise NotImplementedError('the no_skip_connection / no_perceiver ablations are inference-only in voxactb_b200')
            if not self.pos_encoding_with_lang:
                raise NotImplementedError('pos_encoding_with_lang=False is inference-only in voxactb_b200 (the backward writes the '
                                          'gradient of a [77 + T, C] table)')
            if self.weight_tie_layers and self.depth > 1:
                raise NotImplementedError('weight_tie_layers=True is inference-only in voxactb_b200 (the backward writes one '
                                          'gradient buffer per layer slot; tied layers would need their sum)')
            grid = _lib.f32(ins.detach().permute(0, 2, 3, 4, 1))
            plist = [p for p in self._param_table()[2] if p is not None]
            outs = _QnetTrainFn.apply(self, grid, _lib.f32(proprio.detach()), _lib.f32(lang_token_embs.detach()), *plist)
            arm = outs[3] if self.arm_pred_loss else None
            return outs[0], outs[1], outs[2], None, None, None, arm
        with torch.no_grad():
            grid = _lib.f32(ins.permute(0, 2, 3, 4, 1))      # no copy when ins is the permuted view
            proprio = _lib.f32(proprio)
            lang = _lib.f32(lang_token_embs)
            if proprio.shape != (B, self.low_dim_size):
                raise ValueError('proprio must be [%d,%d], got %s' % (B, self.low_dim_size, tuple(proprio.shape)))
            if self.TWO_ROBOTS:
                proprio2 = _lib.f32(proprio2)
                if proprio2.shape != (B, self.low_dim_size):
                    raise ValueError('proprio_left must be [%d,%d], got %s' % (B, self.low_dim_size, tuple(proprio2.shape)))
            if lang.shape != (B, 77, 512):
                raise ValueError('lang_token_embs must be [%d,77,512], got %s' % (B, tuple(lang.shape)))
            L = _lib.lib()
            desc = self._desc()
            arr, keep, slots = self._param_table()
            dev = ins.device
            self._ensure_prepared(desc, arr, slots, dev)
            ws_bytes = L.vxb_qnet_workspace_bytes(ctypes.byref(desc), B)
            if ws_bytes == 0:
                _lib.check(-2, 'vxb_qnet_workspace_bytes')
            if self._workspace is None or self._workspace.numel() < ws_bytes or self._workspace.device != dev:
                self._workspace = None
                self._workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            nrg = self.num_rotation_classes * 3 + self.num_grip_classes
            new = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)
            trans, rot_grip, coll = new(B, 1, V, V, V), new(B, nrg), new(B, self.num_collision_classes)
            trans2 = rot_grip2 = coll2 = None
            if self.TWO_ROBOTS:
                trans2, rot_grip2, coll2 = new(B, 1, V, V, V), new(B, nrg), new(B, self.num_collision_classes)
            arm = new(B, 2) if self.arm_pred_loss else None
            rc = L.vxb_qnet_forward_f32(ctypes.byref(desc), arr, _lib.ptr(self._prepared), _lib.ptr(grid),
                                        _lib.ptr(proprio), _lib.ptr(proprio2) if self.TWO_ROBOTS else None,
                                        _lib.ptr(lang), B, _lib.ptr(trans), _lib.ptr(trans2),
                                        _lib.ptr(rot_grip), _lib.ptr(coll), _lib.ptr(rot_grip2), _lib.ptr(coll2),
                                        _lib.ptr(arm), _lib.ptr(self._workspace), ws_bytes, _lib.stream())
            _lib.check(rc, 'vxb_qnet_forward_f32')
            self.last_launch_count = L.vxb_last_launch_count()
            del keep
        return trans, rot_grip, coll, trans2, rot_grip2, coll2, arm


def _train_methods():
    def _train_setup(self, grid, proprio, lang):
        B, V = grid.shape[0], grid.shape[1]
        if proprio.shape != (B, self.low_dim_size):
            raise ValueError('proprio must be [%d,%d], got %s' % (B, self.low_dim_size, tuple(proprio.shape)))
        if lang.shape != (B, 77, 512):
            raise ValueError('lang_token_embs must be [%d,77,512], got %s' % (B, tuple(lang.shape)))
        L = _lib.lib()
        desc = self._desc()
        arr, keep, slots = self._param_table()
        dev = grid.device
        ws_bytes = L.vxb_qnet_train_workspace_bytes(ctypes.byref(desc), B)
        if ws_bytes == 0:
            _lib.check(-2, 'vxb_qnet_train_workspace_bytes')
        if self._train_ws is None or self._train_ws.numel() < ws_bytes or self._train_ws.device != dev:
            self._train_ws = None
            self._train_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        return L, desc, arr, keep, slots, ws_bytes

    def _forward_train(self, grid, proprio, lang):
        """grid [B,V,V,V,10] channels-last.  Returns ((trans, rot_grip, collision[, arm]), generation)."""
        with torch.no_grad():
            B, V = grid.shape[0], grid.shape[1]
            dev = grid.device
            L, desc, arr, keep, slots, ws_bytes = _train_setup(self, grid, proprio, lang)
            self._ensure_prepared(desc, arr, slots, dev)          # training mode: always re-prepared
            if self.dropout_seed is None:
                self.dropout_seed = int(torch.initial_seed()) & 0xffffffffffff
            self._train_gen += 1
            o = _lib.TrainOpts()
            o.struct_bytes = ctypes.sizeof(_lib.TrainOpts)
            o.input_dropout, o.attn_dropout, o.decoder_dropout = self.input_dropout, self.attn_dropout, self.decoder_dropout
            o.seed = (self.dropout_seed + self._train_gen) & 0xffffffffffffffff
            self._train_opts = o
            nrg = self.num_rotation_classes * 3 + self.num_grip_classes
            new = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)
            trans, rot_grip, coll = new(B, 1, V, V, V), new(B, nrg), new(B, self.num_collision_classes)
            arm = new(B, 2) if self.arm_pred_loss else None
            rc = L.vxb_qnet_forward_train_f32(ctypes.byref(desc), arr, _lib.ptr(self._prepared), _lib.ptr(grid),
                                              _lib.ptr(proprio), _lib.ptr(lang), B, _lib.ptr(trans), _lib.ptr(rot_grip),
                                              _lib.ptr(coll), _lib.ptr(arm), ctypes.byref(o), _lib.ptr(self._train_ws),
                                              ws_bytes, _lib.stream())
            _lib.check(rc, 'vxb_qnet_forward_train_f32')
            del keep
        outs = (trans, rot_grip, coll) + ((arm,) if self.arm_pred_loss else ())
        return outs, self._train_gen

    def _backward_train(self, gen, inputs, gouts, debug=None, out=None):
        """Gradients of every parameter (slot order, Nones skipped) from d loss / d outputs.  out: optional list of
        pre-allocated contiguous gradient buffers (e.g. views of a flat all-reduce arena), one per parameter."""
        if gen != self._train_gen:
            raise RuntimeError('voxactb_b200: backward through a training forward whose saved activations were overwritten by '
                               'a later forward of the same encoder (one forward/backward pair at a time per encoder)')
        grid, proprio, lang = inputs
        with torch.no_grad():
            B, V = grid.shape[0], grid.shape[1]
            dev = grid.device
            L, desc, arr, keep, slots, ws_bytes = _train_setup(self, grid, proprio, lang)
            nrg = self.num_rotation_classes * 3 + self.num_grip_classes
            shapes = [(B, 1, V, V, V), (B, nrg), (B, self.num_collision_classes), (B, 2)]
            g = []
            for i in range(3 + int(bool(self.arm_pred_loss))):
                t = gouts[i] if i < len(gouts) else None
                g.append(torch.zeros(shapes[i], dtype=torch.float32, device=dev) if t is None else _lib.f32(t))
            if out is not None:
                it = iter(out)
                grads = [None if p is None else next(it) for p in slots]
                for p, t in zip(slots, grads):
                    if p is not None and (t.shape != p.shape or not t.is_contiguous() or t.dtype != torch.float32):
                        raise ValueError('gradient buffer does not match its parameter')
            else:
                grads = [None if p is None else torch.empty_like(p, memory_format=torch.contiguous_format) for p in slots]
            garr = (ctypes.c_void_p * len(slots))(*[None if t is None else t.data_ptr() for t in grads])
            darr = None
            if debug is not None:
                darr = (ctypes.c_void_p * 8)(*[None if t is None else t.data_ptr() for t in debug])
            rc = L.vxb_qnet_backward_f32(ctypes.byref(desc), arr, _lib.ptr(self._prepared), _lib.ptr(grid), _lib.ptr(proprio),
                                         _lib.ptr(lang), B, _lib.ptr(g[0]), _lib.ptr(g[1]), _lib.ptr(g[2]),
                                         _lib.ptr(g[3]) if self.arm_pred_loss else None, garr, ctypes.byref(self._train_opts),
                                         darr, _lib.ptr(self._train_ws), ws_bytes, _lib.stream())
            _lib.check(rc, 'vxb_qnet_backward_f32')
            del keep
        return [t for t in grads if t is not None]

    PerceiverVoxelLangEncoder._forward_train = _forward_train
    PerceiverVoxelLangEncoder._backward_train = _backward_train


_train_methods()


class PerceiverVoxelLang2RobotsEncoder(PerceiverVoxelLangEncoder):
    """Drop-in for reference perceiver_lang_io.py:488-860: one shared trunk with two proprio streams
    (C = 192) and a second (left-arm) set of translation / rotation-grip-collision heads."""

    TWO_ROBOTS = True

    def __init__(self, depth, iterations, voxel_size, initial_dim, low_dim_size, layer=0,
                 num_rotation_classes=72, num_grip_classes=2, num_collision_classes=2, input_axis=3,
                 num_latents=512, im_channels=64, latent_dim=512, cross_heads=1, latent_heads=8,
                 cross_dim_head=64, latent_dim_head=64, activation='relu', weight_tie_layers=False,
                 pos_encoding_with_lang=True, input_dropout=0.1, attn_dropout=0.1,
                 decoder_dropout=0.0, lang_fusion_type='seq', voxel_patch_size=9,
                 voxel_patch_stride=8, no_skip_connection=False, no_perceiver=False,
                 no_language=False, final_dim=64):
        super().__init__(depth, iterations, voxel_size, initial_dim, low_dim_size, layer,
                         num_rotation_classes, num_grip_classes, num_collision_classes, input_axis,
                         num_latents, im_channels, latent_dim, cross_heads, latent_heads,
                         cross_dim_head, latent_dim_head, activation, weight_tie_layers,
                         pos_encoding_with_lang, input_dropout, attn_dropout, decoder_dropout,
                         lang_fusion_type, voxel_patch_size, voxel_patch_stride, no_skip_connection,
                         no_perceiver, no_language, final_dim, arm_pred_loss=False)

    def forward(self, ins, proprio_right, proprio_left, lang_goal_emb, lang_token_embs,
                prev_layer_voxel_grid, bounds, prev_layer_bounds, mask=None):
        """Returns (trans_right, rot_and_grip_right, collision_right, trans_left, rot_and_grip_left,
        collision_left), reference perceiver_lang_io.py:860."""
        return self._run(ins, proprio_right, proprio_left, lang_token_embs, mask)[:6]
