"""Detection post-processing and input preparation on the device — the callers either side of the hot path.

``DFINEPostProcessor`` keeps the reference's name, constructor and ``forward(outputs, input_h, input_w)`` contract
(/root/reference/src/dl/export.py:20-100; the same selection as ``Torch_model._preds_postprocess``,
src/infer/torch_model.py:153-227): sigmoid -> top-K over the Q*C scores -> labels / queries -> cxcywh to xyxy in input
pixels.  On the CUDA table it is two launches (top-k selection + gather/convert kernel, csrc/select.cu + csrc/io.cu).

``prepare_inputs`` is the device half of ``Torch_model._prepare_inputs`` (src/infer/torch_model.py:262-292) and of the
train loader's multiscale resize (src/dl/dataset.py:675-683): uint8 HWC batches become float32 [0,1] tensors (resized,
BGR->RGB) in one kernel; the returned tensor is an NCHW *view* of the NHWC buffer the first conv kernel reads.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .kernels import K


class DFINEPostProcessor(nn.Module):
    def __init__(self, num_classes: int, num_top_queries: int = 300, use_focal_loss: bool = True):
        super().__init__()
        if not use_focal_loss:
            raise NotImplementedError("the softmax branch is not on any shipped config's path (use_focal_loss=True)")
        self.num_classes, self.num_top_queries, self.use_focal_loss = num_classes, num_top_queries, use_focal_loss

    @torch.no_grad()
    def forward(self, outputs: dict, input_h: int, input_w: int):
        logits, boxes = outputs["pred_logits"], outputs["pred_boxes"]
        k = min(self.num_top_queries, logits.shape[1] * logits.shape[2])
        labels, abs_boxes, scores, qidx = K.postprocess(logits, boxes, k, float(input_h), float(input_w), True)
        result = (labels, abs_boxes, scores)
        masks = outputs.get("pred_masks")
        if masks is not None:
            hm, wm = masks.shape[2], masks.shape[3]
            result = result + (masks.gather(1, qidx[..., None, None].expand(-1, -1, hm, wm)),)
        return result


@torch.no_grad()
def prepare_inputs(images_u8, size=None, bgr=True):
    """images_u8: uint8 [B,H,W,3] (HWC, as cv2 / the decoder delivers them) on the device.  Returns float32 [B,3,H',W'] in
    [0,1] — an NCHW view of an NHWC buffer — resized to ``size`` (bilinear, half-pixel centres) and BGR->RGB swapped."""
    x = K.preprocess_u8(images_u8, size, 1.0 / 255.0, bgr)        # [B,H',W',3] float32
    return x.permute(0, 3, 1, 2)


@torch.no_grad()
def multiscale_resize(images, size):
    """The train collate's multiscale augmentation (dataset.py:675-683) on a float32 [B,3,H,W] batch, on the device."""
    x = images.permute(0, 2, 3, 1)
    return K.resize_images(x, size).permute(0, 3, 1, 2)
