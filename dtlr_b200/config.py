"""The SLConfig surface of the hot path.

The reference reads every hyper-parameter as `args.<key>` after copying an SLConfig onto an argparse namespace
(finetuning.py:149-155).  `build_dino(args)` here reads the same key names; these helpers produce such a namespace
for the shipped configurations without needing the reference's config loader (util/slconfig.py needs addict+yapf).
A real SLConfig-derived namespace from the reference works unchanged.
"""
from types import SimpleNamespace

# hot-path keys of reference config/Latin_CTC.py:3-116 (values as shipped)
_LATIN_CTC = dict(
    modelname="dino", num_classes=166, backbone="resnet50", dilation=False, position_embedding="sine",
    pe_temperatureH=20, pe_temperatureW=20, return_interm_indices=[1, 2, 3], backbone_freeze_keywords=None,
    lr_backbone=1e-10, enc_layers=6, dec_layers=6, unic_layers=0, pre_norm=False, dim_feedforward=2048,
    hidden_dim=256, dropout=0.0, nheads=8, num_queries=900, query_dim=4, num_patterns=0,
    random_refpoints_xy=False, fix_refpoints_hw=-1, dec_layer_number=None, num_feature_levels=4, enc_n_points=4,
    dec_n_points=4, decoder_layer_noise=False, use_deformable_box_attn=False, box_attn_type="roi_align",
    add_channel_attention=False, add_pos_value=False, two_stage_type="standard", two_stage_pat_embed=0,
    two_stage_add_query_num=0, two_stage_bbox_embed_share=False, two_stage_class_embed_share=False,
    two_stage_learn_wh=False, two_stage_default_hw=0.05, two_stage_keep_all_tokens=False, num_select=300,
    transformer_activation="relu", batch_norm_type="FrozenBatchNorm2d", masks=False, aux_loss=True,
    decoder_sa_type="sa", decoder_module_seq=["sa", "ca", "ffn"], nms_iou_threshold=-1,
    dec_pred_bbox_embed_share=True, dec_pred_class_embed_share=True, use_dn=False, dn_number=100,
    dn_box_noise_scale=0.4, dn_label_noise_ratio=0.5, embed_init_tgt=True, dn_labelbook_size=167,
    match_unstable_error=True, use_detached_boxes_dec_out=False, focal_alpha=0.25, cls_loss_coef=1.0,
    bbox_loss_coef=5.0, giou_loss_coef=2.0, interm_loss_coef=1.0, no_interm_box_loss=False, CTC_loss_coef=1,
    set_cost_class=2.0, set_cost_bbox=5.0, set_cost_giou=2.0, matcher_type="HungarianMatcher", frozen_weights=None, device="cuda",
)


def latin_ctc_args(**overrides):
    """config/Latin_CTC.py (IAM English; BASELINE configs 2, 4, 5)."""
    d = dict(_LATIN_CTC)
    d.update(overrides)
    return SimpleNamespace(**d)


def latin_args(**overrides):
    """config/Latin.py (synthetic pre-training; differs from Latin_CTC on the hot path only by use_dn=True; lr_backbone 1e-5, Latin.py:7)."""
    d = dict(_LATIN_CTC, use_dn=True, lr_backbone=1e-5)
    d.update(overrides)
    return SimpleNamespace(**d)


def hwdb_args(**overrides):
    """config/HWDB_full.py (Chinese, 7356 classes; BASELINE config 3)."""
    d = dict(_LATIN_CTC, num_classes=7356, dn_labelbook_size=7356, dn_number=0, dn_label_noise_ratio=1.5, lr_backbone=1e-5)
    d.update(overrides)
    return SimpleNamespace(**d)
