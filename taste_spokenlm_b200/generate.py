"""KV-cached drop-in for `TasteSpokenLM.generate` (reference loop: MT:1027-1199).

The reference re-forwards the WHOLE growing `inputs_embeds` through the Llama backbone on every generated token
(MT:1111-1117 builds no cache and MT:1196-1199 appends one embedding per step), so a completion of n tokens after a
prompt of L costs sum_k (L + k) token-forwards.  Everything downstream of the backbone only ever looks at the LAST
position (`TasteSampler.text_sample` / `taste_sample` index `[:, -1:]`, sampler.py:88-110; the bridge extractors are
position-wise), so the loop below feeds the prompt once, keeps the backbone's `past_key_values`, and afterwards forwards
only the one new fused embedding: L + n token-forwards, same tokens.

On a GPU the one-position step is additionally captured in a **CUDA graph** over a static KV cache (`_GraphedDecoder`):
a 1 B-parameter Llama at batch 1 is launch-bound in eager PyTorch (~7 ms per forward whether it carries 1 or 100
positions), so the cache alone saves FLOPs but no time; replaying the captured step costs what its ~200 small
kernels and the 2.4 GB of weights cost the GPU.  `cuda_graph=False` (or a CPU backbone) keeps the eager cached loop.

It is host-side orchestration over the reference's own modules — the Llama backbone, `lm_head`, the bridge
(`extract_for_bridge_out_llm`, `fuse_for_bridge_in_llm`), `encode_audio`, `_prepare_single` and `TasteSampler` are called
exactly as the reference calls them — and it returns the same 4-tuple.  `tower.install(patch_generate=True)` binds it as
`taste_speech.modeling_taste.TasteSpokenLM.generate`; the RVQ methods the bridge reaches into (`get_indices_from_code`,
`get_output_from_indices`) are the CUDA ones of `ResidualVQB200` when the B200 tower is installed.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn.functional as F

IGNORE_ID = -1          # taste_speech/modules_taste/cosyvoice/utils.py


def _last_position(outputs, want_layers: bool):
    """What the bridge extractors and lm_head read from the backbone's output, restricted to the newest position."""
    hs = None
    if want_layers and getattr(outputs, "hidden_states", None) is not None:
        hs = tuple(h[:, -1:, :] for h in outputs.hidden_states)
    return SimpleNamespace(last_hidden_state=outputs.last_hidden_state[:, -1:, :], hidden_states=hs)


class _EagerDecoder:
    """KV-cached backbone steps with the backbone's own dynamic cache (any device)."""

    def __init__(self, backbone):
        self.backbone, self.past = backbone, None

    def prefill(self, embeds):
        return self.step(embeds)

    def step(self, embeds):
        out = self.backbone(inputs_embeds=embeds, past_key_values=self.past, use_cache=True, attention_mask=None,
                            output_hidden_states=True, return_dict=True)
        self.past = out.past_key_values
        return out


class _GraphedDecoder:
    """The same steps over a static KV cache, the one-position step captured once in a CUDA graph and replayed.

    Static buffers: the new position's embedding `x [1, 1, H]` (input), the backbone's `last_hidden_state` and per-layer
    `hidden_states` of that position (outputs, overwritten by the next replay; every consumer reads them before the next
    step), and the cache itself: transformers' `StaticCache` keeps the number of cached positions in a per-layer DEVICE
    tensor (`cumulative_length`, advanced in place by every update) from which the captured forward derives the write
    slot, the rotary position and the attention mask over the not-yet-written tail - so a replay is position-correct with
    no host-side argument.  When the cache fills up the history is re-fed into one twice as long (the reference has no
    length limit either)."""

    def __init__(self, backbone, device, dtype, max_len):
        from transformers import StaticCache
        self.backbone, self.device, self.dtype = backbone, device, dtype
        self.max_len = int(max_len)
        self.cache = StaticCache(config=backbone.config, max_cache_len=self.max_len)
        hid = backbone.config.hidden_size
        self.x = torch.zeros(1, 1, hid, dtype=dtype, device=device)
        self.graph, self.out = None, None
        self.length = 0
        self.history = []

    def _forward(self, embeds):
        return self.backbone(inputs_embeds=embeds, past_key_values=self.cache, use_cache=True, attention_mask=None,
                             output_hidden_states=True, return_dict=True)

    def _set_length(self, n):
        for layer in self.cache.layers:                  # in place: the captured graph reads these addresses
            layer.cumulative_length.fill_(int(n))

    def prefill(self, embeds):
        n = embeds.shape[1]
        if n > self.max_len:
            raise ValueError(f"prompt of {n} positions exceeds the static cache ({self.max_len})")
        self._set_length(0)                              # a reused decoder starts over; stale slots are masked out
        out = self._forward(embeds)
        self.length = n
        self.history = [embeds]
        return out

    def _capture(self):
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):                    # warm-up outside the capture (lazy initialisations, autotuning);
            for _ in range(2):                           # each run writes slot `length` (same values) and advances the
                self._forward(self.x)                    # counters, which are put back
                self._set_length(self.length)
        torch.cuda.current_stream(self.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._forward(self.x)
        self._set_length(self.length)                    # capture does not execute, but keep the invariant explicit

    def step(self, embeds):
        if self.length >= self.max_len:                  # grow: replay the history into a cache twice as long
            bigger = _GraphedDecoder(self.backbone, self.device, self.dtype, 2 * self.max_len)
            bigger.prefill(torch.concat(self.history, dim=1))
            self.__dict__.update(bigger.__dict__)
        self.x.copy_(embeds)
        if self.graph is None:
            self._capture()
        self.graph.replay()                              # writes slot `length`, advances the device-side counters
        self.length += 1
        self.history.append(embeds.clone())
        return self.out


def _prompt(lm, embed_tokens, vq_module, mode, llm_indices, llm_token_ids, llm_token_lengths, llm_word_ids, kwargs, dtype,
            device):
    """Prompt embeddings / ids / not-yet-consumed prefix audio embeddings for the four conditional modes (MT:1073-1106)."""
    pending = None
    if mode == "zero":
        embeds = embed_tokens.weight[lm.sos_id].reshape(1, 1, -1)
        ids = torch.tensor([[lm.sos_id]], device=device)
    elif mode == "text":
        embeds, ids = embed_tokens(llm_token_ids), llm_token_ids
    elif mode in ("audio", "instruct"):
        fused, _labels, audio = lm._prepare_single(embed_tokens, vq_module, single_indices=llm_indices[0],
                                                   single_token_ids=llm_token_ids[0], single_word_ids=llm_word_ids[0],
                                                   output_audio_embed=True)
        n_text = llm_token_lengths[0].item() + 1
        if mode == "audio":
            embeds = fused[:n_text, :].unsqueeze(0).to(dtype=dtype, device=device)
            pending = audio[n_text - 1:, :]
            ids = llm_token_ids
        else:
            body = fused[1:n_text, :].unsqueeze(0).to(dtype=dtype, device=device)
            pre = kwargs.get("instruct_prefix_ids").view(1, -1)
            suf = kwargs.get("instruct_suffix_ids").view(1, -1)
            embeds = torch.concat([embed_tokens(pre), body, embed_tokens(suf)], dim=1)
            ids = torch.concat([pre, llm_token_ids[:, 1:], suf], dim=1)
    else:
        raise ValueError(f"unknown conditional_mode {mode!r}")
    return embeds, ids, pending


@torch.no_grad()
def generate_kv_cached(self, vq_module, conditional_mode, llm_indices=None, llm_token_ids=None, llm_token_lengths=None,
                       llm_word_ids=None, extra_words=32, **kwargs):
    """Same arguments and result as `TasteSpokenLM.generate` (MT:1027-1199):
    `(generated_llm_indices [1, n, 4], generated_llm_token_ids [1, m], generated_llm_token_lengths [1, 1] int32,
    generated_llm_word_ids [1, m] int32)`, each `None` when nothing of its kind was generated."""
    vq_module.eval()
    assert llm_indices is None or llm_indices.size(0) == 1, \
        "batch size only allow 1 when `spoken_lm.conditional_generate`"
    base = self.language_model.base_model.model if self._use_lora else self.language_model
    embed_tokens, backbone, lm_head = base.model.embed_tokens, base.model, base.lm_head
    dtype = next(embed_tokens.parameters()).dtype
    device = base.device

    if conditional_mode == "text":
        has_prefix, stop_id = False, None
    elif conditional_mode == "instruct":
        has_prefix, stop_id = False, kwargs.get("stop_id")
    else:
        has_prefix, stop_id = llm_token_ids is not None, None
    self.taste_sampler.reset(extra_words=extra_words, has_prefix=has_prefix, stop_id=stop_id)

    step_embeds, input_ids, pending_audio = _prompt(self, embed_tokens, vq_module, conditional_mode, llm_indices,
                                                    llm_token_ids, llm_token_lengths, llm_word_ids, kwargs, dtype, device)
    want_layers = True          # the layer-mixing extractors (bridge.py: *WeightedLayer*, LinearAllConcat) read every layer

    taste_rows, token_ids, word_ids = [], [], []
    last_audio = None
    pad_audio = self.pad_audio_unit_embed.reshape(1, 1, -1)
    use_graph = bool(kwargs.get("cuda_graph", True)) and torch.device(device).type == "cuda"
    if use_graph:
        limit = int(getattr(backbone.config, "max_position_embeddings", 1 << 30))
        decoder = _GraphedDecoder(backbone, device, dtype, min(limit, step_embeds.shape[1] + 24 * int(extra_words) + 256))
    else:
        decoder = _EagerDecoder(backbone)
    first = True
    while True:
        out = decoder.prefill(step_embeds) if first else decoder.step(step_embeds)   # the KV cache is the only carried state
        first = False
        newest = _last_position(out, want_layers)
        text_logits = lm_head(newest.last_hidden_state)
        taste_logits, _ = self.extract_for_bridge_out_llm(newest, vq_module)
        text_id, taste_ids, action, taste_action = self.taste_sampler.update(text_logits, taste_logits, input_ids=input_ids)
        input_ids = F.pad(input_ids, (0, 1), "constant", text_id)

        if action not in ("wait_for_taste", "terminate"):
            token_ids.append(text_id)
        if action == "continue_at_word_start":
            word_ids.append(word_ids[-1] + 1 if word_ids else 0)
        elif action == "continue_not_at_word_start":
            word_ids.append(word_ids[-1])

        text_embed = embed_tokens.weight[text_id].reshape(1, 1, -1)
        if taste_action == "sample":
            taste_rows.append(taste_ids)
            if taste_ids[0, 0, 0].item() != IGNORE_ID:                  # a word start carries a fresh audio token
                last_audio = self.encode_audio(taste_ids, vq_module)
            audio_embed = last_audio
        elif taste_action.startswith("use_prefix"):
            if taste_action == "use_prefix":
                assert pending_audio is not None and pending_audio.size(0) > 0
                last_audio = pending_audio[0, :].reshape(1, 1, -1)
                pending_audio = pending_audio[1:, :] if pending_audio.size(0) > 1 else None
            audio_embed = last_audio
        else:
            audio_embed = pad_audio
        step_embeds = self.fuse_for_bridge_in_llm(text_embed, audio_embed).to(dtype=dtype, device=device)

        if action == "terminate":
            break

    gen_indices = torch.concat(taste_rows, dim=1) if taste_rows else None
    gen_ids = torch.tensor([token_ids], device=device, dtype=torch.int64) if token_ids else None
    gen_len = torch.full((1, 1), len(token_ids), device=device, dtype=torch.int32) if token_ids else None
    gen_wid = torch.tensor([word_ids], device=device, dtype=torch.int32) if word_ids else None
    return gen_indices, gen_ids, gen_len, gen_wid
