"""Small constructors shared by the oracle-side scripts (TEST INFRASTRUCTURE)."""
import torch


def tiny_block(state):
    """Decoder layer of oracle/gen_golden.tiny_llama, loaded from a recorded state dict."""
    from transformers import LlamaConfig
    from transformers.models.llama.modeling_llama import LlamaDecoderLayer

    cfg = LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, vocab_size=128, max_position_embeddings=64, rms_norm_eps=1e-5,
                      rope_theta=10000.0, tie_word_embeddings=False)
    cfg._attn_implementation = "sdpa"
    blk = LlamaDecoderLayer(cfg, 0).to(torch.bfloat16).eval()
    blk.load_state_dict(state)
    return blk
