"""ctypes binding of libvoxactb.so (the C ABI declared in include/voxactb.h).

No torch types cross the boundary: tensors are passed as raw device pointers plus sizes, and the
current torch CUDA stream as an opaque handle.  There is no CPU fallback: every compute entry
point raises if the library is missing or the tensors are not CUDA fp32.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libvoxactb.so')

MATH_FP32_SIMT = 0
MATH_F16X3 = 1
MATH_BF16X3 = MATH_F16X3   # round-1 name
MATH_F16F8C = 2

c_int, c_float, c_size_t, c_void_p, c_ll = (ctypes.c_int, ctypes.c_float, ctypes.c_size_t,
                                            ctypes.c_void_p, ctypes.c_longlong)


class QnetDesc(ctypes.Structure):
    """struct vxb_qnet_desc (include/voxactb.h)."""
    _fields_ = [(n, ctypes.c_int32) for n in (
        'struct_bytes', 'voxel_size', 'patch_size', 'patch_stride', 'initial_dim', 'im_channels',
        'low_dim_size', 'two_robots', 'lang_seq_len', 'lang_emb_dim', 'num_latents', 'latent_dim',
        'depth', 'iterations', 'cross_heads', 'cross_dim_head', 'latent_heads', 'latent_dim_head',
        'final_dim', 'num_rotation_classes', 'num_grip_classes', 'num_collision_classes',
        'arm_pred_loss', 'no_language')] + [('act_slope', ctypes.c_float), ('math_mode', ctypes.c_int32),
                                                 ('final_input', ctypes.c_int32)]


class TrainOpts(ctypes.Structure):
    """struct vxb_train_opts (include/voxactb.h)."""
    _fields_ = [('struct_bytes', ctypes.c_int32), ('input_dropout', ctypes.c_float), ('attn_dropout', ctypes.c_float),
                ('decoder_dropout', ctypes.c_float), ('seed', ctypes.c_uint64)]


# name -> (restype, argtypes); every symbol include/voxactb.h declares
SIGNATURES = {
    'vxb_version': (c_int, []),
    'vxb_last_error': (ctypes.c_char_p, []),
    'vxb_check_device': (c_int, []),
    'vxb_voxelize_workspace_bytes': (c_size_t, [c_int] * 4),
    'vxb_voxelize_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                 c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    'vxb_voxelize_depth_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                       c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'vxb_qnet_num_params': (c_int, [ctypes.POINTER(QnetDesc)]),
    'vxb_qnet_prepared_bytes': (c_size_t, [ctypes.POINTER(QnetDesc)]),
    'vxb_qnet_workspace_bytes': (c_size_t, [ctypes.POINTER(QnetDesc), c_int]),
    'vxb_qnet_prepare': (c_int, [ctypes.POINTER(QnetDesc), ctypes.POINTER(c_void_p), c_void_p,
                                 c_size_t, c_void_p]),
    'vxb_qnet_forward_f32': (c_int, [ctypes.POINTER(QnetDesc), ctypes.POINTER(c_void_p), c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_size_t, c_void_p]),
    'vxb_qnet_train_workspace_bytes': (c_size_t, [ctypes.POINTER(QnetDesc), c_int]),
    'vxb_qnet_forward_train_f32': (c_int, [ctypes.POINTER(QnetDesc), ctypes.POINTER(c_void_p), c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                           ctypes.POINTER(TrainOpts), c_void_p, c_size_t, c_void_p]),
    'vxb_qnet_backward_f32': (c_int, [ctypes.POINTER(QnetDesc), ctypes.POINTER(c_void_p), c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.POINTER(c_void_p),
                                      ctypes.POINTER(TrainOpts), ctypes.POINTER(c_void_p), c_void_p, c_size_t, c_void_p]),
    'vxb_last_launch_count': (c_int, []),
    'vxb_voxelize_launches': (c_int, []),
    'vxb_profile_stage_count': (c_int, []),
    'vxb_profile_stage_name': (ctypes.c_char_p, [c_int]),
    'vxb_profile_enable': (c_int, [c_int]),
    'vxb_profile_read': (c_int, [ctypes.POINTER(ctypes.c_double)]),
    'vxb_select_action_workspace_bytes': (c_size_t, [c_int, c_int]),
    'vxb_select_action_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                      c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_size_t, c_void_p]),
    'vxb_se3_perturb_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_ll, c_void_p]),
    'vxb_act_tail_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_void_p]),
    'vxb_umma_launch_count': (c_ll, []),
    'vxb_linear_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'vxb_linear_f32': (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p,
                               c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_void_p, c_size_t,
                               c_void_p]),
    'vxb_layernorm_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'vxb_spatial_softmax_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'vxb_spatial_softmax_f32': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                        c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    'vxb_conv3d_workspace_bytes': (c_size_t, [c_int] * 5),
    'vxb_conv3d_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                               c_int, c_int, c_float, c_int, c_void_p, c_size_t, c_void_p]),
    'vxb_upconv3d_workspace_bytes': (c_size_t, [c_int] * 6),
    'vxb_upconv3d_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                 c_int, c_int, c_float, c_int, c_void_p, c_size_t, c_void_p]),
    'vxb_attention_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int]),
    'vxb_attention_f32': (c_int, [c_void_p, c_int, c_ll, c_void_p, c_void_p, c_int, c_ll, c_void_p,
                                  c_int, c_ll, c_int, c_int, c_int, c_int, c_int, c_float, c_int,
                                  c_void_p, c_size_t, c_void_p]),
    'vxb_ce_loss_workspace_bytes': (c_size_t, [c_int, c_int]),
    'vxb_ce_loss_f32': (c_int, [c_void_p, c_ll, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_ll,
                                c_void_p, c_size_t, c_void_p]),
    'vxb_optimizer_workspace_bytes': (c_size_t, [c_int, ctypes.POINTER(c_ll)]),
    'vxb_lamb_step_f32': (c_int, [c_int, ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p),
                                  ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), ctypes.POINTER(c_ll),
                                  c_float, c_float, c_float, c_float, c_float, c_void_p, c_size_t, c_void_p]),
    'vxb_adam_step_f32': (c_int, [c_int, ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p),
                                  ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), ctypes.POINTER(c_ll),
                                  c_int, c_float, c_float, c_float, c_float, c_float, c_void_p, c_size_t,
                                  c_void_p]),
    'vxb_nccl_unique_id': (c_int, [ctypes.c_char_p]),
    'vxb_nccl_init': (c_int, [ctypes.c_char_p, c_int, c_int, ctypes.POINTER(c_void_p)]),
    'vxb_nccl_destroy': (c_int, [c_void_p]),
    'vxb_allreduce_grads': (c_int, [c_void_p, c_void_p, c_size_t, c_float, c_size_t, c_void_p]),
}

_lib = None


def lib():
    """Load libvoxactb.so (once).  Raises -- never falls back -- if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'voxactb_b200: %s is missing -- run `python -m voxactb_b200.build` '
                '(there is no CPU or PyTorch fallback for this path)' % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().vxb_last_error().decode('utf-8', 'replace')
        if rc == -2:
            raise ValueError('%s: %s' % (what, msg))
        raise RuntimeError('%s failed (%d): %s' % (what, rc, msg))


def ptr(t):
    """Device pointer of a contiguous CUDA fp32/int32 tensor (or NULL for None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('voxactb_b200 needs CUDA tensors (got %s); there is no CPU fallback' % t.device)
    if not t.is_contiguous():
        raise RuntimeError('voxactb_b200 needs contiguous tensors')
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def f32(t):
    """Contiguous fp32 view/copy; dtype conversion is explicit here, never inside the library."""
    if t.dtype != torch.float32:
        raise TypeError('voxactb_b200 computes in fp32; got %s' % t.dtype)
    return t.contiguous()
