"""Host-side mirror of the reference's codes/loss.py (VGG-19 content + style loss).

Same class names, constructor arguments, forward signature, return tuples and state_dict layout
(feature_extractor_model.features.{idx}.{weight,bias}); the VGG convolutions and the loss reductions
run in the sm_100a kernels behind the C ABI; with a grad-requiring output image the call records the VGG
activations and its backward runs the hand-written adjoint kernels (train_engine.py).
"""
from __future__ import annotations

import os

import torch
from torch import nn

from . import engine
from .style_transformer import packed_weights, workspace_of


class VGG19_custom(nn.Module):
    """Mirror of loss.py:15-37: holds vgg19.features[:30]; forward returns [relu2_1, relu3_1, relu4_1, relu5_1]
    as NCHW fp32 tensors (converted from the kernels' bf16 NHWC taps)."""

    def __init__(self, features: nn.Module):
        super().__init__()
        self.features = features

    def forward(self, x: torch.Tensor):
        if not x.is_cuda:
            raise RuntimeError("mastermetastyletransfer_b200 runs on sm_100a only: inputs must be CUDA tensors (no CPU fallback)")
        if torch.is_grad_enabled() and x.requires_grad:
            raise NotImplementedError("gradients flow through custom_loss.forward (its autograd node runs the VGG adjoint kernels, "
                                      "autograd_fns.perceptual_loss_apply); a bare feature_extractor_model(x) call records no tape")
        with torch.no_grad():
            w = packed_weights(self, engine.VggWeights)
            ws = workspace_of(self, x.device)
            taps = engine.vgg_taps_forward(w, x.float().contiguous(), ws, "vggmod_")
            N = x.shape[0]
            return [t.view(N, h, wd, c).permute(0, 3, 1, 2).float() for (t, h, wd, c) in taps]


class VGG19_custom_with_batch_norm(nn.Module):
    """Mirror of loss.py:41-63: holds vgg19_bn.features[:43]; forward returns [relu2_1, relu3_1, relu4_1, relu5_1] NCHW fp32.
    Like the reference, BatchNorm follows the module's mode: batch statistics in train mode (the reference's scripts never put
    the loss in eval mode), running statistics after .eval().  (The train-mode forward does not move the running statistics:
    they play no role in anything the reference computes from this module.)"""

    def __init__(self, features: nn.Module):
        super().__init__()
        self.features = features

    def forward(self, x: torch.Tensor):
        if not x.is_cuda:
            raise RuntimeError("mastermetastyletransfer_b200 runs on sm_100a only: inputs must be CUDA tensors (no CPU fallback)")
        if torch.is_grad_enabled() and x.requires_grad:
            raise NotImplementedError("the VGG-19-BN loss variant has forward kernels only (SURVEY.md 8f-4)")
        with torch.no_grad():
            w = packed_weights(self, engine.VggBnWeights)
            taps = engine.vgg_bn_taps_forward(w, x.float().contiguous(), workspace_of(self, x.device), "vggbnmod_", self.training)
            N = x.shape[0]
            return [t.view(N, h, wd, c).permute(0, 3, 1, 2).float() for (t, h, wd, c) in taps]


class custom_loss(nn.Module):
    """Mirror of loss.py:71-336.  total = content + lambda * style (:243)."""

    def __init__(self, project_absolute_path, feature_extractor_model_relative_path=None, use_vgg19_with_batchnorm=False,
                 default_lambda_value=10, distance_content="euclidian", distance_style="euclidian"):
        super().__init__()
        assert distance_content in ["euclidian", "euclidian_squared"], "distance should be either 'euclidian' or 'euclidian_squared'"
        assert distance_style in ["euclidian", "euclidian_squared"], "distance should be either 'euclidian' or 'euclidian_squared'"
        self.use_vgg19_with_batchnorm = bool(use_vgg19_with_batchnorm)
        if feature_extractor_model_relative_path is None:
            feature_extractor_model_relative_path = os.path.join("weights", "vgg_19_last_layer_is_relu_5_1_output_bn.pt" if use_vgg19_with_batchnorm
                                                                 else "vgg_19_last_layer_is_relu_5_1_output.pt")
        self.lambda_value = default_lambda_value
        self.distance_content, self.distance_style = distance_content, distance_style
        # False (default): output_similarity_loss=True returns what the reference computes -- its get_similarity_loss hands the
        # CONTENT features to both arguments of every term (loss.py:333-334), i.e. exactly 0.  True: the paper's form, content vs
        # OUTPUT features, on the tensor cores (csrc/similarity.cu); forward only.
        self.similarity_content_vs_output = False
        # parameter-free, kept for attribute compatibility with the reference (loss.py:102-105)
        self.IN_0, self.IN_1 = nn.InstanceNorm2d(128), nn.InstanceNorm2d(256)
        self.IN_2, self.IN_3 = nn.InstanceNorm2d(512), nn.InstanceNorm2d(512)
        path = os.path.join(project_absolute_path, feature_extractor_model_relative_path)
        if os.path.exists(path):
            features = torch.load(path, weights_only=False)
        elif use_vgg19_with_batchnorm:  # offline: same architecture, random init (the reference would download IMAGENET1K_V1 here)
            from torchvision.models import vgg19_bn
            features = nn.Sequential(*list(vgg19_bn(weights=None).features)[0:43])
        else:
            from .synthetic import build_vgg19_to_relu5_1
            features = build_vgg19_to_relu5_1()
        self.feature_extractor_model = (VGG19_custom_with_batch_norm if use_vgg19_with_batchnorm else VGG19_custom)(features)
        for p in self.feature_extractor_model.parameters():
            p.requires_grad = False

    def forward(self, content_image, style_image, output_image, distance="euclidian", lambda_value=None,
                output_content_and_style_loss=False, output_similarity_loss=False):
        # the reference overwrites an explicitly passed lambda with the default (loss.py:189-190): reproduced
        if lambda_value is not None:
            lambda_value = self.lambda_value
        return self.get_overall_loss(content_image=content_image, style_image=style_image, output_image=output_image,
                                     loss_weight=lambda_value, output_content_and_style_loss=output_content_and_style_loss,
                                     output_similarity_loss=output_similarity_loss)

    def get_overall_loss(self, content_image, style_image, output_image, loss_weight=None, output_content_and_style_loss=False,
                         output_similarity_loss=False):
        assert content_image.shape == style_image.shape == output_image.shape, "All images should be in the exact same shape"
        assert content_image.requires_grad == False, "Content image should not require gradient"  # noqa: E712
        assert style_image.requires_grad == False, "Style image should not require gradient"  # noqa: E712
        if not output_image.is_cuda:
            raise RuntimeError("mastermetastyletransfer_b200 runs on sm_100a only: inputs must be CUDA tensors (no CPU fallback)")
        if loss_weight is None:
            loss_weight = self.lambda_value
        want_sim = bool(output_similarity_loss and self.similarity_content_vs_output)
        if self.use_vgg19_with_batchnorm:
            if (torch.is_grad_enabled() and output_image.requires_grad) or want_sim:
                raise NotImplementedError("the VGG-19-BN loss variant has forward kernels for the content / style loss only (SURVEY.md 8f-4)")
            with torch.no_grad():
                w = packed_weights(self.feature_extractor_model, engine.VggBnWeights)
                out3 = engine.perceptual_loss_forward_bn(w, content_image, style_image, output_image, float(loss_weight),
                                                         self.distance_content == "euclidian_squared", self.distance_style == "euclidian_squared",
                                                         workspace_of(self, output_image.device), self.feature_extractor_model.training)
            return self._pack(out3[0], out3[1], out3[2], output_content_and_style_loss, output_similarity_loss)
        if torch.is_grad_enabled() and output_image.requires_grad:
            if want_sim:
                raise NotImplementedError("the content-vs-output similarity loss has a forward kernel only (SURVEY.md 8f-3)")
            from .autograd_fns import perceptual_loss_apply
            out3 = perceptual_loss_apply(self, content_image, style_image, output_image, float(loss_weight))
            return self._pack(out3[0], out3[1], out3[2], output_content_and_style_loss, output_similarity_loss)
        with torch.no_grad():
            w = packed_weights(self.feature_extractor_model, engine.VggWeights)
            ws = workspace_of(self, output_image.device)
            out3 = engine.perceptual_loss_forward(w, content_image.float(), style_image.float(), output_image.float(), float(loss_weight),
                                                  self.distance_content == "euclidian_squared",
                                                  self.distance_style == "euclidian_squared", ws, similarity=want_sim)
        sim = None
        if want_sim:
            out3, sim = out3
        return self._pack(out3[0], out3[1], out3[2], output_content_and_style_loss, output_similarity_loss, sim)

    @staticmethod
    def _pack(total, content, style, want_parts: bool, want_similarity: bool, similarity=None):
        """Return tuple of get_overall_loss (loss.py:245-262).  The reference's get_similarity_loss compares the CONTENT image's
        relu3_1 / relu4_1 self-similarity maps with THEMSELVES (loss.py:333-334 passes VGG_features_content_layers twice), so the
        value it returns is mean(|tril(D) - tril(D)|) + ... = exactly 0 whatever the inputs; the drop-in returns that 0 (an fp32
        scalar on the device, no graph) without building the two B x N x N cosine maps.  The paper's content-vs-output form
         is computed when `similarity_content_vs_output` is set (`similarity` then holds it)."""
        if want_similarity:
            if similarity is None:
                similarity = torch.zeros((), dtype=torch.float32, device=total.device)
            return (total, content, style, similarity) if want_parts else (total, similarity)
        return (total, content, style) if want_parts else total
