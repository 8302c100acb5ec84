"""ctypes mirror of include/surf_b200.h (the C ABI of libsurf_b200.so)."""
from __future__ import annotations

import ctypes as C

SURF_ABI_VERSION = 2
SURF_MAX_SETS = 8
KIND = {'disk': 0, 'plane': 1, 'sphere': 2, 'triangle': 3}
KIND_NAME = {v: k for k, v in KIND.items()}

c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class SurfPrimSet(C.Structure):
    _fields_ = [('kind', C.c_int32), ('count', C.c_int32),
                ('pos', C.c_void_p), ('pos_stride', C.c_int32),
                ('normal', C.c_void_p), ('normal_stride', C.c_int32),
                ('radius', C.c_void_p), ('material_idx', C.c_void_p)]


class SurfScene(C.Structure):
    _fields_ = [('n_sets', C.c_int32), ('sets', SurfPrimSet * SURF_MAX_SETS),
                ('n_lights', C.c_int32), ('light_pos', C.c_void_p), ('light_pos_stride', C.c_int32),
                ('light_color_idx', C.c_void_p), ('light_attenuation', C.c_void_p), ('ambient', C.c_void_p),
                ('n_colors', C.c_int32), ('colors', C.c_void_p),
                ('n_materials', C.c_int32), ('albedo', C.c_void_p), ('coeffs', C.c_void_p),
                ('gamma', C.c_void_p)]


class SurfCamera(C.Structure):
    _fields_ = [('proj', C.c_int32), ('width', C.c_int32), ('height', C.c_int32),
                ('fovy', C.c_double), ('focal_length', C.c_double),
                ('eye', C.c_void_p), ('at', C.c_void_p), ('up', C.c_void_p),
                ('near_clip', C.c_float), ('far_clip', C.c_float)]


class SurfOptions(C.Structure):
    _fields_ = [('double_sided', C.c_int32), ('use_quartic', C.c_int32), ('shadow', C.c_int32),
                ('pixel_begin', C.c_int32), ('pixel_end', C.c_int32), ('forced_nearest', C.c_int32),
                ('pixels_per_thread', C.c_int32), ('chunk_prims', C.c_int32), ('math_mode', C.c_int32)]


class SurfOutputs(C.Structure):
    _fields_ = [('image', C.c_void_p), ('depth', C.c_void_p), ('normal', C.c_void_p), ('pos', C.c_void_p),
                ('nearest', C.c_void_p), ('ray_dir', C.c_void_p)]


class SurfOutGrads(C.Structure):
    _fields_ = [('image', C.c_void_p), ('depth', C.c_void_p), ('normal', C.c_void_p), ('pos', C.c_void_p)]


class SurfPrimSetGrads(C.Structure):
    _fields_ = [('pos', C.c_void_p), ('normal', C.c_void_p), ('radius', C.c_void_p)]


class SurfSceneGrads(C.Structure):
    _fields_ = [('sets', SurfPrimSetGrads * SURF_MAX_SETS),
                ('light_pos', C.c_void_p), ('light_attenuation', C.c_void_p), ('ambient', C.c_void_p),
                ('colors', C.c_void_p), ('albedo', C.c_void_p), ('coeffs', C.c_void_p), ('gamma', C.c_void_p)]


class SurfBatchLayout(C.Structure):
    """element strides between consecutive scenes of a strided batch (0 = shared by all scenes)"""
    _fields_ = [('set_pos', C.c_int64 * SURF_MAX_SETS), ('set_normal', C.c_int64 * SURF_MAX_SETS),
                ('set_radius', C.c_int64 * SURF_MAX_SETS), ('set_material_idx', C.c_int64 * SURF_MAX_SETS),
                ('light_pos', C.c_int64), ('light_color_idx', C.c_int64), ('light_attenuation', C.c_int64),
                ('ambient', C.c_int64), ('colors', C.c_int64), ('albedo', C.c_int64), ('coeffs', C.c_int64),
                ('gamma', C.c_int64), ('eye', C.c_int64), ('at', C.c_int64), ('up', C.c_int64)]


class SurfStepMSE(C.Structure):
    _fields_ = [('target_image', C.c_void_p), ('loss_scale', C.c_float), ('loss', C.c_void_p), ('grad_image', C.c_void_p)]


class SurfProjection(C.Structure):
    _fields_ = [('batch', C.c_int32), ('n_surfels', C.c_int32), ('pos_stride', C.c_int32), ('width', C.c_int32), ('height', C.c_int32),
                ('fovy', C.c_double), ('focal_length', C.c_double), ('eye', C.c_void_p), ('at', C.c_void_p), ('up', C.c_void_p),
                ('eye_stride', C.c_int64), ('at_stride', C.c_int64), ('up_stride', C.c_int64)]


class SurfScatter(C.Structure):
    _fields_ = [('batch', C.c_int32), ('n', C.c_int32), ('channels', C.c_int32), ('n_dst', C.c_int32), ('mode', C.c_int32),
                ('sigma', C.c_float), ('z_scale', C.c_float), ('use_depth', C.c_int32), ('use_center_dist', C.c_int32)]


class SurfBilinear(C.Structure):
    _fields_ = [('batch', C.c_int32), ('n', C.c_int32), ('channels', C.c_int32), ('width', C.c_int32), ('height', C.c_int32),
                ('sigma', C.c_float), ('z_scale', C.c_float), ('use_depth', C.c_int32), ('use_center_dist', C.c_int32),
                ('compute_depth', C.c_int32)]


SURF_ADAM_MAX_TENSORS = 16


class SurfAdamTensors(C.Structure):
    _fields_ = [('count', C.c_int32), ('param', C.c_void_p * SURF_ADAM_MAX_TENSORS), ('size', C.c_int64 * SURF_ADAM_MAX_TENSORS)]


class SurfSplats(C.Structure):
    _fields_ = [('count', C.c_int32), ('z', C.c_void_p), ('z_stride', C.c_int32), ('normal', C.c_void_p),
                ('normal_stride', C.c_int32), ('material_idx', C.c_void_p), ('light_vis', C.c_void_p), ('pos', C.c_void_p),
                ('samples', C.c_int32), ('estimate_normals', C.c_int32), ('ndc_stride', C.c_int32), ('norm_depth', C.c_void_p)]


class SurfSplatBatch(C.Structure):
    _fields_ = [('z', C.c_int64), ('normal', C.c_int64), ('material_idx', C.c_int64), ('light_vis', C.c_int64),
                ('light_pos', C.c_int64), ('eye', C.c_int64)]


class SurfSplatGrads(C.Structure):
    _fields_ = [('z', C.c_void_p), ('normal', C.c_void_p), ('pos', C.c_void_p)]


# every symbol include/surf_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    'surf_abi_version': (C.c_int, []),
    'surf_last_error': (C.c_char_p, []),
    'surf_workspace_bytes': (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    'surf_workspace_bytes_ex': (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    'surf_adam_step': (C.c_int, [C.POINTER(SurfAdamTensors), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                 C.c_float, C.c_float, C.c_void_p]),
    'surf_check_indices': (C.c_int, [C.c_void_p, C.c_void_p]),
    'surf_step_mse': (C.c_int, [C.POINTER(SurfScene), C.POINTER(SurfCamera), C.POINTER(SurfOptions), C.c_void_p, C.c_size_t,
                                C.POINTER(SurfOutputs), C.POINTER(SurfStepMSE), C.POINTER(SurfSceneGrads), C.c_void_p]),
    'surf_step_mse_strided': (C.c_int, [C.c_int32, C.POINTER(SurfScene), C.POINTER(SurfCamera), C.POINTER(SurfBatchLayout),
                                        C.POINTER(SurfOptions), C.c_void_p, C.c_size_t, C.POINTER(SurfOutputs),
                                        C.POINTER(SurfStepMSE), C.POINTER(SurfSceneGrads), C.c_void_p]),
    'surf_forward': (C.c_int, [C.POINTER(SurfScene), C.POINTER(SurfCamera), C.POINTER(SurfOptions),
                               C.c_void_p, C.c_size_t, C.POINTER(SurfOutputs), C.c_void_p]),
    'surf_backward': (C.c_int, [C.POINTER(SurfScene), C.POINTER(SurfCamera), C.POINTER(SurfOptions),
                                C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                C.POINTER(SurfOutGrads), C.POINTER(SurfSceneGrads), C.c_void_p]),
    'surf_forward_batch': (C.c_int, [C.c_int32, C.POINTER(SurfScene), C.POINTER(SurfCamera), C.POINTER(SurfOptions),
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(SurfOutputs), C.c_void_p]),
    'surf_backward_batch': (C.c_int, [C.c_int32, C.POINTER(SurfScene), C.POINTER(SurfCamera), C.POINTER(SurfOptions),
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p),
                                      C.POINTER(C.c_void_p), C.POINTER(SurfOutGrads), C.POINTER(SurfSceneGrads), C.c_void_p]),
    'surf_forward_strided': (C.c_int, [C.c_int32, C.POINTER(SurfScene), C.POINTER(SurfCamera), C.POINTER(SurfBatchLayout),
                                       C.POINTER(SurfOptions), C.c_void_p, C.c_size_t, C.POINTER(SurfOutputs), C.c_void_p]),
    'surf_backward_strided': (C.c_int, [C.c_int32, C.POINTER(SurfScene), C.POINTER(SurfCamera), C.POINTER(SurfBatchLayout),
                                        C.POINTER(SurfOptions), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                        C.POINTER(SurfOutGrads), C.POINTER(SurfSceneGrads), C.c_void_p]),
    'surf_project_surfels': (C.c_int, [C.POINTER(SurfProjection), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'surf_project_surfels_backward': (C.c_int, [C.POINTER(SurfProjection), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'surf_scatter_forward': (C.c_int, [C.POINTER(SurfScatter), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    'surf_scatter_backward': (C.c_int, [C.POINTER(SurfScatter), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'surf_bilinear_acc_floats': (C.c_size_t, [C.POINTER(SurfBilinear)]),
    'surf_bilinear_oit_forward': (C.c_int, [C.POINTER(SurfBilinear), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'surf_bilinear_oit_backward': (C.c_int, [C.POINTER(SurfBilinear), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p]),
    'surf_gaussian_blur': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p]),
    'surf_splats_workspace_bytes': (C.c_size_t, [C.c_int32, C.c_int32]),
    'surf_splats_forward_strided': (C.c_int, [C.c_int32, C.POINTER(SurfScene), C.POINTER(SurfCamera), C.POINTER(SurfOptions),
                                              C.POINTER(SurfSplats), C.POINTER(SurfSplatBatch), C.c_void_p, C.c_size_t,
                                              C.POINTER(SurfOutputs), C.c_void_p]),
    'surf_splats_backward_strided': (C.c_int, [C.c_int32, C.POINTER(SurfScene), C.POINTER(SurfCamera), C.POINTER(SurfOptions),
                                               C.POINTER(SurfSplats), C.POINTER(SurfSplatBatch), C.c_void_p, C.c_size_t,
                                               C.POINTER(SurfOutGrads), C.POINTER(SurfSceneGrads), C.POINTER(SurfSplatGrads),
                                               C.c_void_p]),
    'surf_splats_forward': (C.c_int, [C.POINTER(SurfScene), C.POINTER(SurfCamera), C.POINTER(SurfOptions),
                                      C.POINTER(SurfSplats), C.c_void_p, C.c_size_t, C.POINTER(SurfOutputs), C.c_void_p]),
    'surf_splats_backward': (C.c_int, [C.POINTER(SurfScene), C.POINTER(SurfCamera), C.POINTER(SurfOptions),
                                       C.POINTER(SurfSplats), C.c_void_p, C.c_size_t, C.POINTER(SurfOutGrads),
                                       C.POINTER(SurfSceneGrads), C.POINTER(SurfSplatGrads), C.c_void_p]),
    'surf_context_create': (C.c_void_p, [C.c_int32]),
    'surf_context_destroy': (None, [C.c_void_p]),
    'surf_render_host': (C.c_int, [C.c_void_p, C.POINTER(SurfScene), C.POINTER(SurfCamera),
                                   C.POINTER(SurfOptions), C.POINTER(SurfOutputs)]),
    'surf_render_backward_host': (C.c_int, [C.c_void_p, C.POINTER(SurfScene), C.POINTER(SurfCamera),
                                            C.POINTER(SurfOptions), C.POINTER(SurfOutputs),
                                            C.POINTER(SurfOutGrads), C.c_void_p, C.POINTER(C.c_float),
                                            C.POINTER(SurfSceneGrads)]),
    'surf_step_host_begin': (C.c_int, [C.c_void_p, C.POINTER(SurfScene), C.POINTER(SurfCamera), C.POINTER(SurfOptions),
                                       C.c_void_p, C.c_float]),
    'surf_context_device_grads': (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    'surf_context_stream': (C.c_void_p, [C.c_void_p]),
    'surf_step_host_end': (C.c_int, [C.c_void_p, C.POINTER(SurfSceneGrads), C.POINTER(C.c_float)]),
    'surf_context_last_transfer': (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    'surf_fma_peak': (C.c_double, [C.c_int32, C.c_int32, C.c_void_p]),
    'surf_last_launch_count': (C.c_int, []),
    'surf_set_kernel_timing': (None, [C.c_int32]),
    'surf_last_kernel_ms': (C.c_double, [C.c_int32]),
    'surf_mean_kernel_ms': (C.c_double, [C.c_int32, C.POINTER(C.c_int32)]),
}


def bind(lib):
    """Attach restype/argtypes for every declared symbol; raises AttributeError if one is missing."""
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib
