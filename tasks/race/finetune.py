"""RACE finetuning (reference ``tasks/race/finetune.py``): ``python tasks/main.py --task RACE ...``."""
from megatronapp_b200.models.bert.bert_model import bert_layer_local_spec
from megatronapp_b200.models.bert.classification import MultipleChoice
from megatronapp_b200.training.arguments import core_transformer_config_from_args
from megatronapp_b200.training.global_vars import get_args, get_tokenizer
from megatronapp_b200.training.utils import print_rank_0
from tasks.finetune_utils import finetune
from tasks.race.data import RaceDataset


def main():
    args = get_args()
    tok = get_tokenizer()
    cfg = core_transformer_config_from_args(args)
    model = MultipleChoice(cfg, bert_layer_local_spec(), args.padded_vocab_size, args.max_position_embeddings)
    train = RaceDataset("training", args.train_data, tok, args.seq_length)
    valid = RaceDataset("validation", args.valid_data, tok, args.seq_length)
    hist = finetune(model, train, valid, epochs=args.epochs, batch_size=args.micro_batch_size, lr=args.lr,
                    keep_last=getattr(args, "keep_last", False))
    for h in hist:
        print_rank_0(f" >> |epoch: {h['epoch']}| correct / total: {h['correct']} / {h['total']} = {100 * h['accuracy']:.4f} %")
    return hist
