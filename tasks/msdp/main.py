"""MSDP entry point (reference ``tasks/msdp/main.py``): ``--task MSDP-PROMPT`` generates, ``--task MSDP-EVAL-F1`` scores."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def get_tasks_args(parser):
    g = parser.add_argument_group("msdp")
    g.add_argument("--task", type=str, required=True, help="MSDP-PROMPT | MSDP-EVAL-F1")
    g.add_argument("--sample-input-file", type=str, default=None)
    g.add_argument("--sample-output-file", type=str, default=None)
    g.add_argument("--prompt-file", type=str, default=None)
    g.add_argument("--prompt-type", type=str, default=None, choices=["knowledge", "response"])
    g.add_argument("--num-prompt-examples", type=int, default=10)
    g.add_argument("--guess-file", type=str, default=None)
    g.add_argument("--answer-file", type=str, default=None)
    g.add_argument("--out-seq-length", type=int, default=100)
    g.add_argument("--api-prompt", action="store_true")
    g.add_argument("--megatron-api-url", type=str, default=None)
    return parser


def main():
    from megatronapp_b200.training.global_vars import get_args, get_tokenizer
    from megatronapp_b200.training.initialize import initialize_megatron

    initialize_megatron(extra_args_provider=get_tasks_args, args_defaults={"micro_batch_size": 1})
    args = get_args()
    if args.task == "MSDP-EVAL-F1":
        from tasks.msdp.evaluate import main as run

        return run()
    if args.task != "MSDP-PROMPT":
        raise NotImplementedError(f"Task {args.task} is not implemented.")
    from tasks.msdp import prompt

    if args.api_prompt:
        gen = prompt.api_generate_fn(args.megatron_api_url)
    else:
        from megatronapp_b200.inference import GPTInferenceWrapper, StaticInferenceEngine, TextGenerationController
        from megatronapp_b200.training import checkpointing
        from megatronapp_b200.training.training import get_model
        from pretrain_gpt import model_provider

        model = get_model(model_provider, wrap_with_ddp=False)
        if args.load:
            checkpointing.load_checkpoint(model, None, None)
        wrapper = GPTInferenceWrapper(model[0], 8, args.seq_length, args.padded_vocab_size)
        gen = prompt.engine_generate_fn(StaticInferenceEngine(TextGenerationController(wrapper, get_tokenizer()), 8))
    n = prompt.generate_samples(gen, args.sample_input_file, args.sample_output_file, args.prompt_file, args.prompt_type,
                                args.num_prompt_examples, args.out_seq_length)
    print(f"wrote {n} generations to {args.sample_output_file}")


if __name__ == "__main__":
    main()
