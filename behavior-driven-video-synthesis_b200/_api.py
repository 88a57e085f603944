"""Public names of the package (re-exported by ``bdvs_b200``)."""
from . import _native
from .behavior import PoseSequenceSampler, ResidualBehaviorNet, ResidualRNNDecoder, UnsupervisedTransformer2
from ._native import build as build_native
from .losses import compute_kl_loss, compute_kl_with_prior, kl_loss, latent_kl
from .modules import (Act, DepthToSpace, Downsample, IDAct, NormConv2d, SpaceToDepth, Upsample, VunetRNB,
                      WeightBank, bump_weight_epoch, to_act)
from .optim import FusedAdam
from .perceptual import PerceptualVGG, vgg_loss
from .rendering import BehaviorVideoRenderer, PoseRenderer, StickmanRasterizer, get_line_colors, project_poses
from .training import GraphedTrainStep
from .validation import fp32_validation
from .vunets import DecDown, DecDownAlter, DecUp, EncDown, EncDownAlter, EncUp, VunetAlter, VunetOrg

__all__ = [
    "VunetAlter", "EncUp", "EncDownAlter", "DecUp", "DecDownAlter", "VunetOrg", "EncDown", "DecDown",
    "latent_kl", "compute_kl_loss", "PerceptualVGG", "vgg_loss",
    "UnsupervisedTransformer2", "ResidualBehaviorNet", "ResidualRNNDecoder", "PoseSequenceSampler",
    "NormConv2d", "VunetRNB", "Downsample", "Upsample", "DepthToSpace", "SpaceToDepth", "IDAct",
    "compute_kl_with_prior", "kl_loss", "Act", "to_act", "WeightBank", "build_native",
    "fp32_validation", "FusedAdam", "GraphedTrainStep", "bump_weight_epoch", "StickmanRasterizer", "PoseRenderer",
    "get_line_colors", "project_poses", "BehaviorVideoRenderer",
]
