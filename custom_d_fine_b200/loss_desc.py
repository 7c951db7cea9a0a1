"""ctypes mirror of ``include/dfine_loss_desc.h`` (the launch description of one criterion evaluation) and the
builder that fills it from torch tensors.  ``dfine_loss_desc_size()`` of the library must equal ``ctypes.sizeof(LossDesc)``
(tests/test_abi_cpu.py)."""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int, c_long, c_void_p


class LossDesc(ctypes.Structure):
    _fields_ = [
        ("L", c_int), ("B", c_int), ("Qt", c_int), ("n_dn", c_int), ("Q", c_int), ("C", c_int), ("NB", c_int),
        ("logits", c_void_p), ("boxes", c_void_p), ("corners", c_void_p), ("ref0", c_void_p),
        ("pre_logits", c_void_p), ("pre_boxes", c_void_p), ("enc_logits", c_void_p), ("enc_boxes", c_void_p),
        ("table", c_void_p), ("ncols", c_long), ("n_layer", c_int), ("go_cap", c_int), ("n_dn_entries", c_int),
        ("labels", c_void_p), ("tboxes", c_void_p), ("counts", c_void_p), ("dn_groups", c_float),
        ("project", c_void_p), ("reg_scale", c_void_p), ("alpha", c_float), ("gamma", c_float), ("inv_t", c_float),
        ("maps", c_void_p), ("Qm", c_int), ("cnt", c_void_p),
    ]


def out_count(L):
    return 6 * (L + 2) + 8 * L


def split_out(out, L):
    """finalize's flat vector -> (vfl [2,L+2], l1 [2,L+2], giou [2,L+2], fgl [2,L], ddf [2,L]); head order of a group:
    decoder layers 0..L-1, pre, enc (group 0 = matching queries, group 1 = denoising queries)."""
    H = L + 2
    return (out[0:2 * H].view(2, H), out[2 * H:4 * H].view(2, H), out[4 * H:6 * H].view(2, H),
            out[6 * H:6 * H + 2 * L].view(2, L), out[6 * H + 2 * L:6 * H + 4 * L].view(2, L))


def build(t, meta):
    """t: dict of contiguous tensors (logits, boxes, corners, ref0, pre_logits, pre_boxes, enc_logits, enc_boxes, table,
    labels, tboxes, counts, project, reg_scale); meta: dict of ints / floats."""
    d = LossDesc()
    L, B, Qt, C = t["logits"].shape
    d.L, d.B, d.Qt, d.C = L, B, Qt, C
    d.n_dn, d.Q = int(meta["n_dn"]), int(Qt - meta["n_dn"])
    d.NB = t["corners"].shape[-1] // 4
    for k in ("logits", "boxes", "corners", "ref0", "pre_logits", "pre_boxes", "enc_logits", "enc_boxes", "table",
              "labels", "tboxes", "counts", "project", "reg_scale"):
        setattr(d, k, t[k].data_ptr())
    d.ncols = t["table"].shape[1]
    d.n_layer, d.go_cap, d.n_dn_entries = int(meta["n_layer"]), int(meta["go_cap"]), int(meta["n_dn_entries"])
    d.dn_groups = float(meta["dn_groups"])
    d.alpha, d.gamma, d.inv_t = float(meta["alpha"]), float(meta["gamma"]), 1.0 / float(meta.get("T", 5.0))
    d.maps, d.cnt, d.Qm = None, None, max(d.Q, d.n_dn)
    return d
