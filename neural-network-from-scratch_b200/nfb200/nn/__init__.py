"""Neural-network layers (mirror of neural_arch.nn for the Transformer hot path)."""
from ..core import Module, Parameter
from .layers import (GELU, Dropout, Embedding, LayerError, LayerNorm, ModuleList, ReLU, RMSNorm, Sequential, Sigmoid, Softmax,
                     Tanh)
from .linear import Linear, OptimizedLinear, StandardLinear
from .transformer import MultiHeadAttention, SelfAttention, TransformerBlock, TransformerDecoderBlock, TransformerEncoder
from .modern_attention import FlashAttentionConcept, GroupedQueryAttention, MultiQueryAttention
from .positional import (LearnedPE, LearnedPositionalEmbedding, RoPE, RotaryPositionalEmbedding, SinusoidalPE,
                         SinusoidalPositionalEncoding, create_rope, create_sinusoidal_pe)
from .modern_activations import GeGLU, SiLU, SwiGLU, Swish
from .differential_attention import DifferentialAttention, DifferentialTransformer, DifferentialTransformerBlock
