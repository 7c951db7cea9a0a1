"""Regenerates include/dfine_sm100.h from the DFINE_API definitions in csrc/*.cu.

The header is the reviewed, committed artefact (tests/test_abi.py checks that it matches the
sources and the built library); this script only saves retyping prototypes.
"""
import glob
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
GEOM = ("int B, int H, int W, int Cin, int OH, int OW, int Cout, int KH, int KW, int stride, int pad_t, "
        "int pad_l, long ldx, long ldy")

GROUPS = [
    ("lib.cu", "Library", "Error convention: every entry point returns 0 on success, a positive cudaError_t from the\n"
     " * launch, or a negative value for an argument / shape / alignment violation; dfine_last_error() returns a\n"
     " * thread-local description.  The library never allocates or frees device memory, never synchronises the\n"
     " * device and keeps no mutable global state besides immutable per-process caches (and the bring-up trace\n"
     " * pointer of dfine_tc_trace, null unless a profiling tool sets it); all entry points are\n"
     " * re-entrant (autograd calls backward kernels from another host thread).  `stream` is a cudaStream_t."),
    ("gemm_tc.cu", "tcgen05 implicit-GEMM convolution / linear (tensor cores, TMA, TMEM)",
     "Replaces cuDNN/cuBLAS behind nn.Conv2d (hgnetv2.py:53-64, hybrid_encoder.py:25-27,86-88) and nn.Linear\n"
     " * (dfine_decoder.py:33-46, nn.MultiheadAttention in/out projections) for 1x1 and 3x3 stride-1 shapes."),
    ("conv_simt.cu", "CUDA-core implicit-GEMM convolution / linear (stem, strided, ragged shapes)",
     "Same contract as above for the shapes the tensor-core path does not take (hgnetv2.py:115-166 stem)."),
    ("stem.cu", "Direct 3-channel image convolution of the stem (forward, weight gradient)",
     "hgnetv2.py:117-124 (stem1)."),
    ("spatial.cu", "Depthwise convolutions, stem max-pool, nearest upsample",
     "hgnetv2.py:83-112,154-162,295-304; hybrid_encoder.py:96-103,472."),
    ("norm_act.cu", "BatchNorm(+add)+activation(+LAB), LayerNorm, activations",
     "hgnetv2.py:25-32,65-80; hybrid_encoder.py:28-45,262-290; common.py:58-67; dfine_decoder.py:202-255."),
    ("attention.cu", "Multi-head attention core",
     "nn.MultiheadAttention as called at hybrid_encoder.py:277 and dfine_decoder.py:239."),
    ("msda.cu", "Multi-scale deformable attention",
     "MSDeformableAttention.forward dfine_decoder.py:137-178 + deformable_attention_core_func_v2 arch/utils.py:191-264."),
    ("fdr.cu", "FDR head (Integral / distance2bbox / LQE statistics)",
     "dfine_decoder.py:291-313, arch/utils.py:119-188."),
    ("loss.cu", "Criterion (VFL, L1 + GIoU, FGL + DDF over every loss head)",
     "DFINECriterion.forward dfine_criterion.py:609-777: loss_labels_vfl 92-122, loss_boxes 124-143, loss_local 145-237,\n"
     " * unimodal_distribution_focal_loss 837-858, bbox2distance / translate_gt arch/utils.py:267-354.  One launch\n"
     " * description (dfine_loss_desc.h) for all entry points; call order: prepare, {vfl,box,fgl_ddf}_fwd, finalize;\n"
     " * backward: {vfl,box,fgl_ddf}_bwd with the same description and workspace."),
    ("select.cu", "Query selection (row max + top-k) and the decoder gate",
     "DFINETransformer._select_topk dfine_decoder.py:875-910; Gate.forward dfine_decoder.py:258-271."),
    ("io.cu", "Input preparation (uint8 -> float, resize) and detection post-processing",
     "Torch_model._prepare_inputs / _preds_postprocess infer/torch_model.py:153-292; DFINEPostProcessor dl/export.py:20-100;\n"
     " * multiscale collate dl/dataset.py:675-683."),
    ("seg.cu", "Segmentation head: GroupNorm, bilinear-resize backward, mask matching cost",
     "MaskDecoder.forward dfine_decoder.py:353-370; mask cost matcher.py:19-71,175-237."),
    ("matcher.cu", "Hungarian matcher (cost blocks + LSAP)",
     "HungarianMatcher.forward matcher.py:110-257 (scipy.optimize.linear_sum_assignment at 243)."),
    ("optim.cu", "Optimizer / EMA",
     "train.py:62-73,512-535; dfine.py:87-124."),
]


def signatures(path):
    s = path.read_text().replace("GEOM_ARGS", GEOM)
    out = []
    for m in re.finditer(r"DFINE_API\s+([\w\s\*]+?)\s+(dfine_\w+)\s*\(([^)]*)\)\s*\{", s):
        ret, name, args = m.groups()
        out.append((ret.strip(), name, " ".join(args.split())))
    return out


def wrap(proto, width=112):
    if len(proto) <= width:
        return proto
    head, args = proto.split("(", 1)
    parts = [a.strip() for a in args.rstrip(");").split(",")]
    lines, cur = [], head + "("
    indent = " " * (len(head) + 1)
    for i, a in enumerate(parts):
        piece = a + (", " if i < len(parts) - 1 else ");")
        if len(cur) + len(piece) > width:
            lines.append(cur.rstrip())
            cur = indent + piece
        else:
            cur += piece
    lines.append(cur)
    return "\n".join(lines)


def main():
    csrc = ROOT / "custom_d_fine_b200" / "csrc"
    lines = ["/* C ABI of libdfine_sm100.so — the drop-in boundary of the B200-native D-FINE hot path.",
             " *",
             " * Plain C: raw device pointers, sizes and a stream; no torch types.  The Python host",
             " * (custom_d_fine_b200/cuda_ops.py) binds these with ctypes and parses THIS file for the argument",
             " * types, so the header is the single source of truth of the boundary.  Citations are paths in the",
             " * reference repository (ArgoHA/custom_d_fine) whose computation each group replaces.",
             " * All tensors are fp32 unless stated; activations are NHWC / token-major; `ld*` are row (pixel)",
             " * strides in elements.  GENERATED by tools/gen_header.py from csrc/*.cu — edit the sources.",
             " */",
             "#ifndef DFINE_SM100_H", "#define DFINE_SM100_H", "",
             '#include "dfine_loss_desc.h"   /* launch description of the criterion entry points (dfine_loss_*) */', "",
             "#ifdef __cplusplus", 'extern "C" {', "#endif", "",
             "/* activation codes */", "#define DFINE_ACT_NONE 0", "#define DFINE_ACT_RELU 1",
             "#define DFINE_ACT_SILU 2", "#define DFINE_ACT_GELU 3", ""]
    for fname, title, doc in GROUPS:
        p = csrc / fname
        if not p.exists():
            continue
        sigs = signatures(p)
        if not sigs:
            continue
        lines += [f"/* ---- {title} ({fname}) ----", f" * {doc}", " */"]
        for ret, name, args in sigs:
            lines.append(wrap(f"{ret} {name}({args});"))
        lines.append("")
    lines += ["#ifdef __cplusplus", "}", "#endif", "#endif /* DFINE_SM100_H */", ""]
    (ROOT / "include" / "dfine_sm100.h").write_text("\n".join(lines))


if __name__ == "__main__":
    main()
