"""Feed for the n-gram rescoring tool: the reference's ngram/prediction_helpers.py:5-46 `get_new_pred_logits`
builds the (B,Q,C+1) "CTC view" probabilities (queries in reading order, class probabilities times `multiply_pred_logits_by`,
synthesised blank) with a dozen torch ops and three full-size temporaries; here it is the optional third output of the fused
decode kernels (csrc/decode.cu), in the same layout, so a torchaudio / flashlight CTC beam-search decoder can consume it
unchanged (that decoder itself is a third-party host library and out of scope)."""
from . import ops


def get_new_pred_logits(output, multiply_pred_logits_by=1, eps=0.003):
    """same name / arguments as the reference helper.  output: model output dict (CUDA).  Returns fp32 (B,Q,C+1)."""
    _, new_pred = ops.ctc_decode(output["pred_logits"], output["pred_boxes"], eps, want_new_pred=True,
                                 prob_scale=float(multiply_pred_logits_by))
    return new_pred
