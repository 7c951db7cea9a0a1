"""Deterministic weights / inputs shared by the golden generator (runs the REAL reference in the build
container) and the tests (run this repo's graph): nothing but a seed has to travel."""
import torch


def seeded_fill(module, seed=0):
    """Overwrite every floating-point state entry in sorted key order from one generator.
    Conv/linear weights ~ N(0, 1/fan_in), norm weights ~ U(0.5,1.5), biases ~ N(0,0.1), running_var
    ~ U(0.5,1.5); zero-initialised heads become non-trivial so every loss term is exercised."""
    g = torch.Generator().manual_seed(seed)
    sd = module.state_dict()
    for k in sorted(sd.keys()):
        v = sd[k]
        if not v.dtype.is_floating_point or k.endswith("anchors") or k.endswith("num_points_scale") \
                or k.endswith(".up"):
            continue
        if k.endswith("running_var"):
            new = torch.rand(v.shape, generator=g) + 0.5
        elif k.endswith("running_mean"):
            new = torch.randn(v.shape, generator=g) * 0.1
        elif "lab.scale" in k:
            new = torch.rand(v.shape, generator=g) * 0.5 + 0.75
        elif "lab.bias" in k:
            new = torch.randn(v.shape, generator=g) * 0.05
        elif v.dim() >= 2:
            fan_in = v[0].numel()
            new = torch.randn(v.shape, generator=g) * (1.0 / max(fan_in, 1)) ** 0.5
            if "sampling_offsets.weight" in k or "attention_weights.weight" in k:
                new = new * 0.3
            if "enc_score_head.weight" in k or "dec_score_head" in k:
                new = new * 4.0          # spread the logits: no near-ties in the top-k query selection
        elif k.endswith("weight"):       # norm scales
            new = torch.rand(v.shape, generator=g) + 0.5
        elif "sampling_offsets.bias" in k:
            new = v.cpu() + torch.randn(v.shape, generator=g) * 0.1
        else:
            new = torch.randn(v.shape, generator=g) * 0.1
        v.copy_(new)
    module.load_state_dict(sd)
    return module


def synthetic_batch(B, H, W, seed=1234, T=(10, 7, 0, 3), num_classes=80):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, H, W, generator=g)
    targets = []
    for b in range(B):
        t = T[b % len(T)]
        cxcy = torch.rand(t, 2, generator=g) * 0.6 + 0.2
        wh = torch.rand(t, 2, generator=g) * 0.25 + 0.05
        targets.append({"labels": torch.randint(0, num_classes, (t,), generator=g),
                        "boxes": torch.cat([cxcy, wh], 1)})
    return x, targets


def rect_masks(boxes, H, W):
    """uint8 [T, H, W] filled GT rectangles of normalised cxcywh boxes (segment config targets)."""
    T = boxes.shape[0]
    m = torch.zeros(T, H, W, dtype=torch.uint8)
    for i in range(T):
        cx, cy, w, h = boxes[i].tolist()
        x0, x1 = int(round((cx - w / 2) * W)), int(round((cx + w / 2) * W))
        y0, y1 = int(round((cy - h / 2) * H)), int(round((cy + h / 2) * H))
        m[i, max(y0, 0):max(min(y1, H), 0), max(x0, 0):max(min(x1, W), 0)] = 1
    return m
