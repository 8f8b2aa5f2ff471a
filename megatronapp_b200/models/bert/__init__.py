from .bert_model import BertLMHead, BertModel, Pooler, bert_extended_attention_mask, bert_layer_local_spec  # noqa: F401
