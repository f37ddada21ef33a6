"""B200-native (sm_100a) AIR hot path: drop-in for the reference's ``air`` package.

    from air_b200 import transformer, batch_transformer, vae, AIRModel, \
        concrete_binary_pre_sigmoid_sample, concrete_binary_kl_mc_sample

Everything numeric runs in hand-written CUDA (csrc/) behind the C ABI in
include/air_b200.h; PyTorch only owns device memory, streams and torch.distributed.
"""
from . import _cabi
from ._cabi import AirError, build, launch_count
from .air.transformer import transformer, batch_transformer, writeback_canvas
from .air.vae import vae
from .air.air_model import AIRModel, reset_variable_scopes
from .air.params import ParamStore
from . import checkpoint, data, dp, ops, tfrecords
from .demo.model_wrapper import ModelWrapper, evaluation_summaries
from .demo.visualize import visualize_reconstructions, draw_colored_bounding_boxes
from .air.concrete import (concrete_binary_sample, concrete_binary_pre_sigmoid_sample,
                           concrete_binary_kl_mc_sample, concrete_step)

__all__ = ["AIRModel", "ModelWrapper", "evaluation_summaries", "visualize_reconstructions", "draw_colored_bounding_boxes", "checkpoint", "vae", "ParamStore", "ops", "reset_variable_scopes", "transformer", "batch_transformer", "writeback_canvas", "concrete_binary_sample",
           "concrete_binary_pre_sigmoid_sample", "concrete_binary_kl_mc_sample", "concrete_step",
           "AirError", "build", "launch_count"]
