"""Pretrain a LLaVA-style vision-language model (reference ``pretrain_vlm.py``: CLIP ViT encoder +
MLP projector + GPT decoder; mock image/text data unless a multimodal dataset is configured)."""
import os

if not os.path.exists("/dev/nvidia0"):
    os.environ.setdefault("OMP_NUM_THREADS", "2")

import argparse
import copy
from functools import partial

import torch

from megatronapp_b200.models.gpt import get_gpt_layer_local_spec
from megatronapp_b200.models.multimodal import DEFAULT_IMAGE_TOKEN_INDEX, LLaVAModel
from megatronapp_b200.models.vision import get_vit_layer_spec
from megatronapp_b200.parallel import state as mpu
from megatronapp_b200.parallel.tp_layers import ColumnParallelLinear, RowParallelLinear
from megatronapp_b200.training.arguments import core_transformer_config_from_args
from megatronapp_b200.training.global_vars import get_args
from megatronapp_b200.training.training import pretrain
from megatronapp_b200.training.utils import print_rank_0
from megatronapp_b200.transformer.mlp import MLPSubmodules


def add_vlm_extra_args(parser):
    g = parser.add_argument_group("vision language model")
    g.add_argument("--img-h", type=int, default=336)
    g.add_argument("--img-w", type=int, default=336)
    g.add_argument("--patch-dim", type=int, default=14)

    def _vision_type(v):
        if v in ("clip", "siglip", "radio") or v.startswith("hf://"):
            return v
        raise argparse.ArgumentTypeError("clip | siglip | radio | hf://<transformers checkpoint>")

    g.add_argument("--vision-model-type", default="clip", type=_vision_type)
    g.add_argument("--max-img-h", type=int, default=2048)
    g.add_argument("--max-img-w", type=int, default=2048)
    g.add_argument("--vision-num-layers", type=int, default=None)
    g.add_argument("--vision-hidden-size", type=int, default=None)
    g.add_argument("--vision-num-attention-heads", type=int, default=None)
    g.add_argument("--freeze-LM", action="store_true")
    g.add_argument("--freeze-ViT", action="store_true")
    g.add_argument("--disable-vision-class-token", action="store_true")
    g.add_argument("--dataloader-seq-length", type=int, default=None)
    g.add_argument("--decoder-tp-comm-overlap", action="store_true",
                   help="accepted for CLI parity: the language model's TP collectives are fused into its GEMMs by default")
    g.add_argument("--use-packed-sequence", action="store_true")
    return parser


def model_provider(pre_process=True, post_process=True, add_encoder=True, add_decoder=True):
    args = get_args()
    print_rank_0("building a multimodal (LLaVA) model ...")
    if getattr(args, "use_packed_sequence", False):
        raise NotImplementedError("--use-packed-sequence (THD packing of image + text samples) is not implemented for the VLM entry "
                                  "point; packed sequences are available for text models (PackedSeqParams / ops.attention_varlen)")
    if args.pipeline_model_parallel_size > 1:
        # the vision tower and projector live with the text embedding on the first pipeline stage (reference
        # pretrain_vlm.py: add_encoder = is_pipeline_first_stage() when no separate encoder pipeline is configured)
        add_encoder = pre_process
    lcfg = core_transformer_config_from_args(args)
    vcfg = copy.deepcopy(lcfg)
    vcfg.num_layers = args.vision_num_layers or args.num_layers
    vcfg.hidden_size = args.vision_hidden_size or args.hidden_size
    vcfg.num_attention_heads = args.vision_num_attention_heads or args.num_attention_heads
    vcfg.num_query_groups = vcfg.num_attention_heads
    vcfg.kv_channels = vcfg.hidden_size // vcfg.num_attention_heads
    vcfg.ffn_hidden_size = 4 * vcfg.hidden_size
    vcfg.pipeline_model_parallel_size = 1
    vcfg.sequence_parallel = False
    vcfg.gated_linear_unit = False
    vcfg.normalization = "LayerNorm"
    vcfg.add_bias_linear = True
    pcfg = copy.deepcopy(lcfg)
    pcfg.sequence_parallel = False
    pcfg.gated_linear_unit = False
    pcfg.add_bias_linear = True
    pcfg.ffn_hidden_size = lcfg.hidden_size
    m = LLaVAModel(
        language_transformer_config=lcfg, language_transformer_layer_spec=get_gpt_layer_local_spec(),
        language_vocab_size=args.padded_vocab_size, language_max_sequence_length=args.max_position_embeddings,
        vision_transformer_config=vcfg, vision_transformer_layer_spec=get_vit_layer_spec(),
        drop_vision_class_token=args.disable_vision_class_token, vision_projection_config=pcfg,
        vision_projection_layer_spec=MLPSubmodules(linear_fc1=ColumnParallelLinear, linear_fc2=RowParallelLinear),
        vision_projection_type="mlp", parallel_output=True,
        language_position_embedding_type=args.position_embedding_type, pre_process=pre_process,
        post_process=post_process, add_encoder=add_encoder, add_decoder=add_decoder, img_h=args.img_h, img_w=args.img_w,
        patch_dim=args.patch_dim, language_rotary_base=args.rotary_base, vision_model_type=args.vision_model_type,
        max_img_h=args.max_img_h, max_img_w=args.max_img_w)
    m.freeze(args.freeze_LM, args.freeze_ViT, False)
    return m


class MockVLMDataset(torch.utils.data.Dataset):
    def __init__(self, n, text_len, vocab, img_h, img_w, seed):
        self.n, self.text_len, self.vocab, self.img_h, self.img_w, self.seed = n, text_len, vocab, img_h, img_w, seed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed + i)
        toks = torch.randint(1, self.vocab, (self.text_len + 1,), generator=g)
        tokens = toks[:-1].clone()
        tokens[1] = DEFAULT_IMAGE_TOKEN_INDEX
        return {"image": torch.rand(3, self.img_h, self.img_w, generator=g), "tokens": tokens, "labels": toks[1:].clone(),
                "loss_mask": torch.ones(self.text_len), "position_ids": torch.arange(self.text_len)}


def train_valid_test_datasets_provider(num_samples):
    args = get_args()
    from megatronapp_b200.models.vision import get_num_image_embeddings

    n_img = get_num_image_embeddings(args.img_h, args.img_w, args.patch_dim, "clip", args.disable_vision_class_token)
    text_len = args.dataloader_seq_length or max(8, args.seq_length - n_img + 1)
    mk = lambda n, s: MockVLMDataset(max(int(n), 1), text_len, args.padded_vocab_size, args.img_h, args.img_w, s)  # noqa: E731
    return mk(num_samples[0], 1), mk(num_samples[1], 2), mk(num_samples[2], 3)


def loss_func(loss_mask, output_tensor):
    losses = output_tensor.float()
    loss_mask = loss_mask.reshape(-1).float()
    total = loss_mask.sum()
    loss = torch.cat([torch.sum(losses.reshape(-1) * loss_mask).view(1), total.view(1)])
    rep = loss.clone().detach()
    g = mpu.get_data_parallel_group()
    if g is not None and not isinstance(g, mpu._SingleGroup) and torch.distributed.get_world_size(g) > 1:
        torch.distributed.all_reduce(rep, group=g)
    return loss[0], loss[1].clone().detach().to(torch.int), {"lm loss": (rep[0], rep[1])}


def forward_step(data_iterator, model):
    dev = next(model.parameters()).device
    d = next(data_iterator)
    out, mask = model(d["image"].to(dev), d["tokens"].to(dev), d["position_ids"].to(dev), None,
                      labels=d["labels"].to(dev), loss_mask=d["loss_mask"].to(dev))
    return out, partial(loss_func, mask)


if __name__ == "__main__":
    train_valid_test_datasets_provider.is_distributed = True
    pretrain(train_valid_test_datasets_provider, model_provider, "encoder_or_decoder", forward_step,
             extra_args_provider=add_vlm_extra_args, args_defaults={"tokenizer_type": "NullTokenizer"})
