/* Launch description of one D-FINE criterion evaluation (every loss head of a train step), shared by the entry points
 * dfine_loss_* of libdfine_sm100.so (see dfine_sm100.h).  Plain C: device pointers, sizes and host scalars.
 *
 * Replaces the arithmetic of DFINECriterion.forward (src/d_fine/dfine_criterion.py:609-777): loss_labels_vfl 92-122,
 * loss_boxes 124-143, loss_local (FGL + DDF) 145-237 with unimodal_distribution_focal_loss 837-858 and bbox2distance /
 * translate_gt (src/d_fine/arch/utils.py:267-354), over main / aux_i / pre / enc_0 / dn_i / dn_pre heads.
 *
 * Heads.  Group A = matching queries (rows [n_dn, n_dn + Q) of the stacked decoder tensors): decoder layers 0..L-1,
 * the `pre` head, the encoder head (own [B,Q,*] tensors).  Group DN = denoising queries (rows [0, n_dn)): decoder layers
 * 0..L-1 and `dn_pre`.  Index sets come from the plan table [4, ncols] (rows: image, query, global target, valid):
 * L+2 Hungarian sets of n_layer columns in the order main (= last layer), aux_0.., pre, enc; then the cross-layer GO
 * union (go_cap columns, padded with valid = 0); then the denoising set (n_dn_entries columns).
 */
#ifndef DFINE_LOSS_DESC_H
#define DFINE_LOSS_DESC_H

typedef struct dfine_loss_desc {
    int L, B, Qt, n_dn, Q, C, NB;     /* layers, images, rows per image (n_dn + Q), denoising rows, queries, classes, bins */
    const float* logits;              /* [L,B,Qt,C] */
    const float* boxes;               /* [L,B,Qt,4] cxcywh */
    const float* corners;             /* [L,B,Qt,4*NB] */
    const float* ref0;                /* [B,Qt,4] initial reference boxes (shared by all layers, detached) */
    const float* pre_logits;          /* [B,Qt,C] */
    const float* pre_boxes;           /* [B,Qt,4] */
    const float* enc_logits;          /* [B,Q,C] */
    const float* enc_boxes;           /* [B,Q,4] */
    const long* table;                /* [4, ncols] int64 */
    long ncols;
    int n_layer, go_cap, n_dn_entries;
    const long* labels;               /* [sumT] int64 */
    const float* tboxes;              /* [sumT,4] cxcywh */
    const float* counts;              /* [2] = (num_boxes_go, num_boxes): world-averaged, clamped at 1 (639-652) */
    float dn_groups;                  /* denoising groups (dn normaliser = num_boxes * groups, 733) */
    const float* project;             /* W(n) [NB] (arch/utils.py:145-188) */
    const float* reg_scale;           /* [1] */
    float alpha, gamma, inv_t;        /* VFL alpha / gamma, 1 / DDF temperature */
    int* maps;                        /* workspace: [L+4, B, Qm] target index per (map, image, query) or -1 */
    int Qm;                           /* max(Q, n_dn) */
    int* cnt;                         /* workspace: [2] valid entries of the GO / denoising sets */
} dfine_loss_desc;

#endif
