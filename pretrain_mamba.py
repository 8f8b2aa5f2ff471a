"""Pretrain Mamba / hybrid models (entry point kept from the reference ``pretrain_mamba.py``)."""
import os

if not os.path.exists("/dev/nvidia0"):   # CPU/gloo mode: avoid OpenMP oversubscription
    os.environ.setdefault("OMP_NUM_THREADS", "2")

from megatronapp_b200.models.mamba import MambaModel
from megatronapp_b200.training.arguments import core_transformer_config_from_args
from megatronapp_b200.training.global_vars import get_args
from megatronapp_b200.training.training import pretrain
from megatronapp_b200.training.utils import print_rank_0
from pretrain_gpt import forward_step, train_valid_test_datasets_provider


def model_provider(pre_process=True, post_process=True):
    args = get_args()
    print_rank_0("building Mamba model ...")
    config = core_transformer_config_from_args(args)
    return MambaModel(config, vocab_size=args.padded_vocab_size, max_sequence_length=args.max_position_embeddings,
                      pre_process=pre_process, post_process=post_process,
                      hybrid_attention_ratio=args.hybrid_attention_ratio, hybrid_mlp_ratio=args.hybrid_mlp_ratio,
                      hybrid_override_pattern=args.hybrid_override_pattern, parallel_output=True,
                      share_embeddings_and_output_weights=not args.untie_embeddings_and_output_weights,
                      position_embedding_type=args.position_embedding_type if args.position_embedding_type in ("rope", "none") else "none",
                      rotary_percent=args.rotary_percent, rotary_base=args.rotary_base)


if __name__ == "__main__":
    pretrain(train_valid_test_datasets_provider, model_provider, "encoder_or_decoder", forward_step,
             args_defaults={"tokenizer_type": "NullTokenizer"})
