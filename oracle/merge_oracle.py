"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

numpy restatement of LlavaForConditionalGeneration._merge_input_ids_with_image_features
(reference: mantis/models/mllava/modeling_llava.py:293-360).  Pinned against the reference itself:
tests/golden/merge_kat.pt is produced by running the *reference* method (oracle/make_golden.py) and
tests/test_oracle.py checks this restatement against those vectors and the SURVEY section 8c KATs.
"""
import numpy as np


def merge_oracle(image_features, inputs_embeds, input_ids, attention_mask, labels,
                 image_token_index, pad_token_id, ignore_index=-100):
    """All arguments numpy arrays.  Returns (final_embedding, final_attention_mask, final_labels|None,
    position_ids, srcmap) ; raises ValueError exactly when the reference does."""
    num_images, P, D = image_features.shape
    B, T = input_ids.shape
    left_padding = not np.sum(input_ids[:, -1] == pad_token_id)                      # :296
    special = input_ids == image_token_index                                        # :298
    n_special = special.sum(-1)
    S = int(n_special.max() * (P - 1) + T)                                           # :301
    bi, ti = np.where(input_ids != image_token_index)                               # :302
    new_pos = np.cumsum(special * (P - 1) + 1, -1) - 1                               # :309
    nb_image_pad = S - 1 - new_pos[:, -1]                                            # :310
    if left_padding:
        new_pos = new_pos + nb_image_pad[:, None]                                    # :312
    text_to = new_pos[bi, ti]                                                        # :313
    final = np.zeros((B, S, D), dtype=inputs_embeds.dtype)                           # :316
    fmask = np.zeros((B, S), dtype=attention_mask.dtype)
    flabels = None
    if labels is not None:
        flabels = np.full((B, S), ignore_index, dtype=input_ids.dtype)               # :323
    final[bi, text_to] = inputs_embeds[bi, ti]                                       # :338
    fmask[bi, text_to] = attention_mask[bi, ti]
    if labels is not None:
        flabels[bi, text_to] = labels[bi, ti]
    img_to = np.all(final == 0, axis=-1)                                             # :344
    img_to &= (np.cumsum(img_to, -1) - 1) >= nb_image_pad[:, None]                   # :345
    if img_to.sum() != num_images * P:                                               # :347
        raise ValueError("The input provided to the model are wrong. image tokens / images mismatch")
    final[img_to] = image_features.reshape(-1, D)                                    # :353
    fmask = fmask | img_to.astype(fmask.dtype)                                       # :354
    pos = np.cumsum(fmask, -1) - 1
    pos[fmask == 0] = 1                                                              # :355
    # inverse map (what the CUDA index kernel emits): >=0 text token, <=-2 image row -(k)-2, -1 zero fill
    srcmap = np.full((B, S), -1, dtype=np.int32)
    srcmap[bi, text_to] = ti
    rank = np.cumsum(img_to.reshape(-1)) - 1
    flat = srcmap.reshape(-1)
    flat[img_to.reshape(-1)] = (-(rank[img_to.reshape(-1)]) - 2).astype(np.int32)
    return final, fmask, flabels, pos, srcmap
