"""Callers of the hot path: pair sampling, losses, the train step and its data-parallel wrapper
(host-side mirror of reference train.py / models/losses / utils/distributed.py)."""
from .losses import (assign_fake_images_to_clusters, flow_identity_loss, gangealing_cluster_loss, gangealing_loss,
                     sample_gan_supervised_pairs, total_variation_loss)
from .latent_learner import DirectionInterpolator
from .perceptual import PerceptualLoss, get_perceptual_loss
from .step import TrainConfig, Trainer, accumulate, requires_grad
from .classifier_step import ClassifierTrainer
