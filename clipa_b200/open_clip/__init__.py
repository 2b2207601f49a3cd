"""Mirror of the `open_clip` surface that clipa_torch/training consumes (open_clip/__init__.py:1-13)."""
from .factory import (add_model_config, create_loss, create_model, create_model_and_transforms,
                      get_model_config, get_tokenizer, list_models, load_checkpoint, trace_model)
from .loss import ClipLoss, gather_features
from .model import (CLIP, CLIPTextCfg, CLIPVisionCfg, CustomTextCLIP, convert_to_custom_text_state_dict,
                    convert_weights_to_fp16, convert_weights_to_lp, get_cast_dtype, resize_pos_embed,
                    resize_text_pos_embed)
from .transformer import (LayerNorm, LayerNormFp32, PatchDropout, ResidualAttentionBlock, TextTransformer,
                          Transformer, VisionTransformer)
