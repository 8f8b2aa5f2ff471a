"""LLaVA-style vision-language model: vision encoder -> projector -> decoder-only LM with image
embeddings spliced in at the image-token positions.

Parity: reference ``models/multimodal/llava_model.py`` (``LLaVAModel.__init__`` with separate
language / vision / projection configs and specs, ``add_encoder/add_decoder`` for encoder-decoder
pipeline splits, ``freeze(...)``, ``_preprocess_data`` combining text and tiled-image embeddings
and rebuilding labels / loss mask, ``forward(images, input_ids, position_ids, attention_mask,
labels, loss_mask, inference_params, num_image_tiles)``).
"""
from __future__ import annotations

from typing import List, Optional

import torch

from ...transformer.mlp import MLPSubmodules
from ...transformer.module import MegatronModule
from ..gpt.gpt_model import GPTModel
from ..vision.clip_vit_model import CLIPViTModel, get_num_image_embeddings
from ..vision.multimodal_projector import MultimodalProjector

IGNORE_INDEX = -100
DEFAULT_IMAGE_TOKEN_INDEX = -200
IMAGE_TOKEN = "<image>"


class LLaVAModel(MegatronModule):
    def __init__(self, language_transformer_config, language_transformer_layer_spec, language_vocab_size: int,
                 language_max_sequence_length: int, vision_transformer_config, vision_transformer_layer_spec,
                 drop_vision_class_token: bool, vision_projection_config, vision_projection_layer_spec: MLPSubmodules,
                 vision_projection_type: str = "mlp", allow_missing_vision_projection_checkpoint: bool = False,
                 parallel_output: bool = True, language_position_embedding_type: str = "learned_absolute",
                 language_rotary_percent: float = 1.0, pre_process: bool = True, post_process: bool = True,
                 add_encoder: bool = True, add_decoder: bool = True, img_h: int = 336, img_w: int = 336,
                 patch_dim: int = 14, language_rotary_base: int = 10000, language_rope_scaling: bool = False,
                 image_token_index: int = DEFAULT_IMAGE_TOKEN_INDEX, pixel_shuffle: bool = False,
                 tile_tags: Optional[list] = None, vision_model_type: str = "clip", max_img_h: int = 2048,
                 max_img_w: int = 2048, class_token_len: Optional[int] = None):
        super().__init__(language_transformer_config)
        self.pre_process, self.post_process = pre_process, post_process
        self.add_encoder, self.add_decoder = add_encoder, add_decoder
        self.image_token_index = image_token_index
        self.drop_vision_class_token = drop_vision_class_token
        self.language_max_sequence_length = language_max_sequence_length
        self.vision_model = self.vision_projection = self.language_model = None
        self.share_embeddings_and_output_weights = False
        if add_decoder:
            self.language_model = GPTModel(
                config=language_transformer_config, transformer_layer_spec=language_transformer_layer_spec,
                vocab_size=language_vocab_size, max_sequence_length=language_max_sequence_length,
                parallel_output=parallel_output, position_embedding_type=language_position_embedding_type,
                rotary_percent=language_rotary_percent, pre_process=pre_process, post_process=post_process,
                rotary_base=language_rotary_base, rope_scaling=language_rope_scaling,
                scatter_embedding_sequence_parallel=False)
            self.share_embeddings_and_output_weights = self.language_model.share_embeddings_and_output_weights
        if add_encoder:
            if str(vision_model_type).startswith("hf://"):
                # a transformers vision tower used as is (reference core/models/huggingface): CLIP / SigLIP
                from ..huggingface import build_hf_model

                self.vision_model = build_hf_model(vision_transformer_config, vision_model_type)
                if self.vision_model.hidden_size:
                    vision_transformer_config.hidden_size = self.vision_model.hidden_size
            elif vision_model_type == "radio":
                from ..vision.radio import RADIOViTModel

                self.vision_model = RADIOViTModel(vision_transformer_config, vision_transformer_layer_spec, img_h=img_h,
                                                  img_w=img_w, patch_dim=patch_dim, max_img_h=max_img_h, max_img_w=max_img_w,
                                                  class_token_len=class_token_len or 8)
            elif vision_model_type in ("clip", "siglip"):
                self.vision_model = CLIPViTModel(vision_transformer_config, vision_transformer_layer_spec, img_h=img_h,
                                                 img_w=img_w, patch_dim=patch_dim, model_subtype=vision_model_type,
                                                 add_class_token=vision_model_type == "clip",
                                                 class_token_len=class_token_len or 1)
            else:
                raise ValueError(f"unknown vision_model_type {vision_model_type!r}")
            self.vision_projection = MultimodalProjector(vision_projection_config, vision_projection_layer_spec,
                                                         vision_projection_type, vision_transformer_config.hidden_size)
        self.vision_model_type = vision_model_type
        if str(vision_model_type).startswith("hf://"):
            from ..huggingface import get_hf_model_type

            vision_model_type = get_hf_model_type(vision_model_type)                 # "clip" / "siglip": class-token accounting
        self.img_seq_len = get_num_image_embeddings(img_h, img_w, patch_dim, vision_model_type, drop_vision_class_token,
                                                    class_token_len or (8 if vision_model_type == "radio" else 1),
                                                    pixel_shuffle)

    def shared_embedding_or_output_weight(self):
        return self.language_model.shared_embedding_or_output_weight() if self.add_decoder else None

    def set_input_tensor(self, input_tensor):
        if isinstance(input_tensor, (list, tuple)):
            input_tensor = input_tensor[0]
        if self.add_encoder and self.add_decoder:
            self.vision_model.set_input_tensor(input_tensor)
        elif self.add_encoder:
            self.vision_model.set_input_tensor(input_tensor)
        elif self.pre_process:
            self.encoder_hidden_state = input_tensor
        else:
            self.language_model.set_input_tensor(input_tensor)

    def freeze(self, freeze_language_model: bool, freeze_vision_model: bool, freeze_vision_projection: bool):
        for on, m in ((freeze_language_model, self.language_model), (freeze_vision_model, self.vision_model),
                      (freeze_vision_projection, self.vision_projection)):
            if on and m is not None:
                for p in m.parameters():
                    p.requires_grad = False

    def _preprocess_data(self, image_embeddings, language_embeddings, input_ids, loss_mask, labels, num_image_tiles):
        """Splice ``[tiles, img_seq, h]`` image embeddings into ``[b, s, h]`` text embeddings at the image
        tokens.  Returns ``[s', b, h]`` embeddings and matching labels / loss mask (image positions ignored),
        truncated / padded to the language model's maximum sequence length."""
        b, s = input_ids.shape
        h = language_embeddings.shape[-1]
        img_seq = image_embeddings.shape[1]
        is_img = input_ids == self.image_token_index
        n_img = is_img.sum(dim=1)
        tiles_per_image = num_image_tiles.tolist() if num_image_tiles is not None else [1] * int(n_img.sum())
        out_e, out_l, out_m = [], [], []
        tile = img = 0
        for i in range(b):
            pieces_e, pieces_l, pieces_m = [], [], []
            prev = 0
            for pos in torch.nonzero(is_img[i], as_tuple=False).flatten().tolist():
                pieces_e.append(language_embeddings[i, prev:pos])
                if labels is not None:
                    pieces_l.append(labels[i, prev:pos])
                    pieces_m.append(loss_mask[i, prev:pos])
                nt = tiles_per_image[img]
                emb = image_embeddings[tile:tile + nt].reshape(nt * img_seq, h)
                pieces_e.append(emb)
                if labels is not None:
                    pieces_l.append(torch.full((nt * img_seq,), IGNORE_INDEX, dtype=labels.dtype, device=labels.device))
                    pieces_m.append(torch.zeros(nt * img_seq, dtype=loss_mask.dtype, device=loss_mask.device))
                    # the token *before* an image must not be trained to predict the first image slot
                    if pieces_l[-2].numel() > 0:
                        pieces_m[-2] = pieces_m[-2].clone()
                        pieces_m[-2][-1] = 0
                tile += nt
                img += 1
                prev = pos + 1
            pieces_e.append(language_embeddings[i, prev:])
            if labels is not None:
                pieces_l.append(labels[i, prev:])
                pieces_m.append(loss_mask[i, prev:])
            out_e.append(torch.cat(pieces_e, dim=0))
            if labels is not None:
                out_l.append(torch.cat(pieces_l))
                out_m.append(torch.cat(pieces_m))
        max_len = min(max(e.shape[0] for e in out_e), self.language_max_sequence_length)
        emb = language_embeddings.new_zeros(b, max_len, h)
        lab = msk = None
        if labels is not None:
            lab = torch.full((b, max_len), IGNORE_INDEX, dtype=labels.dtype, device=labels.device)
            msk = torch.zeros(b, max_len, dtype=loss_mask.dtype, device=loss_mask.device)
        for i, e in enumerate(out_e):
            n = min(e.shape[0], max_len)
            emb[i, :n] = e[:n]
            if labels is not None:
                lab[i, :n] = out_l[i][:n]
                msk[i, :n] = out_m[i][:n]
        return emb.transpose(0, 1).contiguous(), lab, msk

    def forward(self, images: torch.Tensor, input_ids: torch.Tensor, position_ids: torch.Tensor,
                attention_mask: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None,
                loss_mask: Optional[torch.Tensor] = None, inference_params=None,
                num_image_tiles: Optional[torch.Tensor] = None, image_token_index: Optional[int] = None,
                runtime_gather_output: Optional[bool] = None):
        if image_token_index is not None:
            self.image_token_index = image_token_index
        use_cache = inference_params is not None and "image_tokens_count" in getattr(inference_params, "key_value_memory_dict", {})
        image_embeddings = None
        if self.add_encoder and not use_cache and images is not None and images.shape[0] > 0:
            x = self.vision_model(images)                                   # [tiles, img_seq(+cls), hv]
            if self.drop_vision_class_token and getattr(self.vision_model, "add_class_token", False):
                x = x[:, self.vision_model.class_token_len:, :]
            image_embeddings = self.vision_projection(x)                   # [tiles, img_seq, h]
        elif not self.add_encoder and self.pre_process:
            image_embeddings = getattr(self, "encoder_hidden_state", None)
        if not self.add_decoder:
            return image_embeddings, loss_mask
        combined = None
        new_labels, new_mask = labels, loss_mask
        if self.pre_process:
            ids = input_ids.clone()
            ids[ids == self.image_token_index] = 0
            text = self.language_model.embedding(input_ids=ids, position_ids=position_ids).transpose(0, 1)   # [b, s, h]
            if image_embeddings is not None:
                combined, new_labels, new_mask = self._preprocess_data(image_embeddings.to(text.dtype), text, input_ids,
                                                                      loss_mask, labels, num_image_tiles)
            else:
                combined = text.transpose(0, 1).contiguous()
        elif self.post_process and labels is not None and loss_mask is not None:
            # later pipeline stage: the hidden states arrive already spliced; only labels / loss mask must be expanded
            # the same way (zero-width embeddings reuse the splice logic without touching any activation)
            n_tiles = int(num_image_tiles.sum()) if num_image_tiles is not None else int((input_ids == self.image_token_index).sum())
            if n_tiles > 0:
                z = torch.zeros(0, device=input_ids.device)
                _, new_labels, new_mask = self._preprocess_data(z.new_zeros(n_tiles, self.img_seq_len, 0),
                                                                z.new_zeros(*input_ids.shape, 0), input_ids, loss_mask,
                                                                labels, num_image_tiles)
        train_labels = None
        if new_labels is not None:
            train_labels = new_labels.clone()
            train_labels[train_labels == IGNORE_INDEX] = 0
        out = self.language_model(input_ids=None, position_ids=None, attention_mask=attention_mask,
                                  decoder_input=combined, labels=train_labels, inference_context=inference_params,
                                  runtime_gather_output=runtime_gather_output)
        if labels is None or loss_mask is None:
            return out
        return out, new_mask.contiguous()
