"""(f)4 on the GPU (-m gpu): the CUDA-graphed static-cache decode step of `generate.generate_kv_cached` against the eager
cached loop and against full re-forwards (what the reference's loop computes, MT:1111-1117).  The reference tree does not
exist on the GPU box, so the spoken LM around the backbone is a small stand-in with the attributes the loop touches; the
equality of the loop itself with the reference's is the CPU test (tests/test_generate_kv_cached.py)."""
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu

from taste_spokenlm_b200.generate import _EagerDecoder, _GraphedDecoder, generate_kv_cached

torch.set_grad_enabled(False)


def _llama(hidden=128, layers=3, vocab=400, dtype=torch.float32):
    from transformers import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(0)
    cfg = LlamaConfig(vocab_size=vocab, hidden_size=hidden, intermediate_size=256, num_hidden_layers=layers,
                      num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=512)
    m = LlamaForCausalLM(cfg).eval()
    for p in m.parameters():
        p.mul_(3.0)
    return m.to("cuda", dtype)


def test_graphed_decoder_equals_full_forward_and_grows():
    m = _llama()
    emb = torch.randn(1, 60, 128, device="cuda")
    full = m.model(inputs_embeds=emb, use_cache=False, output_hidden_states=True)
    dec = _GraphedDecoder(m.model, torch.device("cuda", 0), torch.float32, max_len=24)      # forces two cache growths
    out = dec.prefill(emb[:, :10])
    assert torch.allclose(out.last_hidden_state, full.last_hidden_state[:, :10], atol=2e-4, rtol=1e-3)
    for t in range(10, 60):
        out = dec.step(emb[:, t:t + 1])
        assert torch.allclose(out.last_hidden_state[:, -1], full.last_hidden_state[:, t], atol=2e-4, rtol=1e-3), t
        for a, b in zip(out.hidden_states, full.hidden_states):
            assert torch.allclose(a[:, -1], b[:, t], atol=2e-4, rtol=1e-3), t
    assert dec.max_len == 96 and dec.length == 60 and dec.graph is not None


class _Sampler:
    """Greedy stand-in with TasteSampler's interface (sampler.py:74-188): word start on even token ids."""

    def reset(self, extra_words, has_prefix=True, stop_id=None):
        self.n, self.limit = 0, extra_words

    def update(self, text_logits, taste_logits, input_ids):
        text_id = int(text_logits[:, -1, :].argmax(-1))
        taste_ids = taste_logits[:, -1:, :, :].argmax(-1)
        self.n += 1
        if self.n > self.limit:
            return text_id, taste_ids, "terminate", "sample"
        return text_id, taste_ids, ("continue_at_word_start" if text_id % 2 == 0 or self.n == 1 else
                                    "continue_not_at_word_start"), "sample"


class _SpokenLM(nn.Module):
    def __init__(self, llama, audio_dim=96):
        super().__init__()
        hid = llama.config.hidden_size
        self.language_model, self._use_lora, self.sos_id = llama, False, 1
        self.taste_sampler = _Sampler()
        self.head = nn.Linear(hid, 4 * 512)
        self.audio = nn.Embedding(512, audio_dim)
        self.fuse = nn.Linear(audio_dim, hid)
        self.pad_audio_unit_embed = nn.Parameter(torch.zeros(audio_dim))

    def extract_for_bridge_out_llm(self, outputs, vq_module):
        h = outputs.last_hidden_state.float()
        return self.head(h).view(1, h.shape[1], 4, 512), None

    def encode_audio(self, taste_ids, vq_module):
        return self.audio(taste_ids.clamp_min(0)).sum(2)

    def fuse_for_bridge_in_llm(self, text_embeds, audio_embeds):
        return text_embeds + self.fuse(audio_embeds)


@pytest.mark.parametrize("mode", ["zero", "text"])
def test_generate_graph_equals_eager_cached_loop(mode):
    lm = _SpokenLM(_llama()).to("cuda").eval()
    vq = nn.Identity()
    kw = dict(extra_words=40)
    if mode == "text":
        kw["llm_token_ids"] = torch.randint(2, 400, (1, 11), device="cuda")
    eager = generate_kv_cached(lm, vq, mode, cuda_graph=False, **kw)
    graph = generate_kv_cached(lm, vq, mode, cuda_graph=True, **kw)
    for a, b in zip(eager, graph):
        assert (a is None) == (b is None)
        if a is not None:
            assert torch.equal(a, b)
    assert eager[1].shape[1] == 40
