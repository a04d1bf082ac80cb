"""Model zoo for the named configs: GPT-2 small/medium, BERT-base MLM, ViT-B/16, translation encoder-decoder."""
from .bert import BERT, BERTConfig, BERTForMaskedLM, bert_base, bert_base_mlm
from .gpt2 import GPT2, GPT2_CONFIGS, GPT2LMHead, GPT2Model, gpt2_large, gpt2_medium, gpt2_small, gpt2_xl
from .translation import PositionalEncoding, TranslationTransformer
from .vit import VisionTransformer, vit_b_16, vit_l_16
from .roberta import (ROBERTA_CONFIGS, RoBERTa, RoBERTaConfig, RoBERTaForMaskedLM, RoBERTaForSequenceClassification, roberta_base,
                      roberta_large)
from .clip import CLIP, CLIP_CONFIGS, CLIPModel, CLIPTextTransformer, CLIPVisionTransformer, clip_base, clip_large
from .modern_transformer import (ModelError, ModernTransformer, PreNormTransformer, PreNormTransformerConfig, RoPETransformer,
                                 prenorm_transformer_base, prenorm_transformer_large, prenorm_transformer_small)
