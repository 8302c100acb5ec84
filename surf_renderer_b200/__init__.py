"""surf_renderer_b200 - B200 (sm_100a) implementation of DiffRend's ray-cast render path.

    from surf_renderer_b200 import render      # instead of: from diffrend.torch.renderer import render

See DESIGN.md for the path and its boundary, INTEGRATION.md for the reference-side binding.
"""
from .along_ray import render_splats_NDC, render_splats_along_ray, render_splats_along_ray_batch, z_to_pcl_CC, z_to_pcl_CC_batched   # noqa: F401
from .graphs import GraphedStep   # noqa: F401
from .ingest import load_model, load_scene, make_torch_var, obj_to_triangle_spec, render_scene   # noqa: F401
from .marshal import select_scenes, set_default_intersect_mode   # noqa: F401
from .projection import (project_image_coordinates, project_surfels, projection_renderer,   # noqa: F401
                         projection_renderer_differentiable_fast, scatter_mean_dim0, scatter_weighted_blended_oit)
from .renderer import get_param_value, render, render_batch, render_flat   # noqa: F401
from .step import MSEStep, PackedAdam, render_mse_step   # noqa: F401

__all__ = ['GraphedStep', 'MSEStep', 'PackedAdam', 'project_image_coordinates', 'project_surfels', 'projection_renderer', 'projection_renderer_differentiable_fast', 'scatter_mean_dim0',
           'scatter_weighted_blended_oit', 'render_mse_step', 'render', 'render_batch', 'render_flat', 'render_splats_NDC', 'render_splats_along_ray', 'render_splats_along_ray_batch', 'get_param_value', 'load_scene', 'make_torch_var', 'load_model',
           'obj_to_triangle_spec', 'render_scene', 'select_scenes', 'set_default_intersect_mode', 'z_to_pcl_CC', 'z_to_pcl_CC_batched']
