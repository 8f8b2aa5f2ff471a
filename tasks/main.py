"""Downstream-task entry point (reference ``tasks/main.py``): ``--task {WIKITEXT103, LAMBADA}`` run the
zero-shot GPT evaluations; the model / parallel flags are the normal training arguments."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def get_tasks_args(parser):
    g = parser.add_argument_group("tasks")
    g.add_argument("--task", type=str, required=True,
                   help="WIKITEXT103 | LAMBADA | MNLI | QQP | RACE | RETRIEVER-EVAL | RET-FINETUNE-NQ | VISION-CLASSIFY (MSDP-*: tasks/msdp/main.py)")
    g.add_argument("--qa-data-test", type=str, default=None, help="NQ-open style TSV: question \\t [answers]")
    g.add_argument("--qa-data-dev", type=str, default=None, help="the development split, evaluated before the test split")
    g.add_argument("--faiss-use-gpu", action="store_true",
                   help="score on the GPU (the evidence embeddings already live on the model's device; kept for CLI parity)")
    g.add_argument("--keep-last", action="store_true", help="keep the last incomplete batch of a finetuning epoch")
    g.add_argument("--faiss-match", type=str, default="string", choices=["regex", "string"])
    g.add_argument("--faiss-topk-retrievals", type=int, default=100)
    g.add_argument("--train-data", nargs="*", default=None)
    g.add_argument("--epochs", type=int, default=3)
    g.add_argument("--valid-data", nargs="*", default=None, help="path(s) to the evaluation text")
    g.add_argument("--overlapping-eval", type=int, default=32, help="sliding window stride for perplexity")
    g.add_argument("--strict-lambada", action="store_true", help="whole-word LAMBADA (all word pieces must match)")
    g.add_argument("--eval-micro-batch-size", type=int, default=None)
    # supervised retriever fine-tuning (RET-FINETUNE-NQ)
    g.add_argument("--train-with-neg", action="store_true", help="add hard negatives to the in-batch negatives")
    g.add_argument("--train-hard-neg", type=int, default=0, help="hard negatives per question")
    g.add_argument("--val-av-rank-hard-neg", type=int, default=30)
    g.add_argument("--val-av-rank-other-neg", type=int, default=30)
    g.add_argument("--sample-rate", type=float, default=1.0)
    return parser


if __name__ == "__main__":
    from megatronapp_b200.training.global_vars import get_args
    from megatronapp_b200.training.initialize import initialize_megatron

    initialize_megatron(extra_args_provider=get_tasks_args)
    args = get_args()
    if args.task in ("LAMBADA", "WIKITEXT103"):
        from tasks.zeroshot_gpt.evaluate import main
    elif args.task in ("MNLI", "QQP"):
        from tasks.glue.finetune import main
    elif args.task == "RACE":
        from tasks.race.finetune import main
    elif args.task == "VISION-CLASSIFY":
        from tasks.vision.classification import main
    elif args.task == "RET-FINETUNE-NQ":
        from tasks.orqa.supervised import main
    elif args.task == "RETRIEVER-EVAL":
        def main():
            from megatronapp_b200.training import checkpointing
            from megatronapp_b200.training.global_vars import get_tokenizer
            from megatronapp_b200.training.training import get_model
            from pretrain_ict import model_provider
            from tasks.orqa.evaluate import ORQAEvaluator, read_evidence, read_nq_questions

            model = get_model(model_provider, wrap_with_ddp=False)
            if args.load:
                checkpointing.load_checkpoint(model, None, None)
            ev = ORQAEvaluator(model[0], get_tokenizer(), args.retriever_seq_length, args.indexer_batch_size)
            docs = read_evidence(args.evidence_data_path)
            ev.index_evidence(docs)
            for split, path in (("DEV", args.qa_data_dev), ("TEST", args.qa_data_test)):
                if path:
                    print(f"evaluating the {split} split: {path}", flush=True)
                    qs, ans = read_nq_questions(path)
                    ev.evaluate(qs, ans, docs, args.retriever_report_topk_accuracies or (1, 5, 20), args.faiss_match)
    else:
        raise NotImplementedError(f"Task {args.task} is not implemented.")
    main()
