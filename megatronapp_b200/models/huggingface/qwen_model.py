"""HF Qwen language model as the LLaVA decoder (reference ``core/models/huggingface/qwen_model.py``): takes the combined
(text + image) embeddings as ``decoder_input [s, b, h]`` and returns logits ``[b, s, v]`` (or the loss when labels are
given), plus the ``embedding`` hook the container uses to embed text tokens."""
import torch

from .module import HuggingFaceModule, load_auto_model


class QwenHuggingFaceModel(HuggingFaceModule):
    def __init__(self, config):
        super().__init__(config)
        self.model = load_auto_model(config.huggingface_model_name_or_path, causal_lm=True)

    def embedding(self, input_ids, position_ids=None):
        return self.model.get_input_embeddings()(input_ids).transpose(0, 1)            # [s, b, h] like LanguageModelEmbedding

    def forward(self, input_ids=None, position_ids=None, attention_mask=None, decoder_input=None, labels=None, **kw):
        if decoder_input is not None:
            out = self.model(inputs_embeds=decoder_input.transpose(0, 1), position_ids=position_ids, attention_mask=attention_mask)
        else:
            out = self.model(input_ids=input_ids, position_ids=position_ids, attention_mask=attention_mask)
        logits = out.logits
        if labels is None:
            return logits
        return torch.nn.functional.cross_entropy(logits.float().flatten(0, 1), labels.flatten(), reduction="none").view_as(labels)
