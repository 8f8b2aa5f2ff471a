"""Print the theoretical per-GPU memory of a configuration without touching a GPU
(reference ``tools/report_theoretical_memory.py``).  Takes the normal training flags."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from megatronapp_b200.training.arguments import parse_args, validate_args  # noqa: E402
from megatronapp_b200.training.theoretical_memory_usage import report_theoretical_memory  # noqa: E402

if __name__ == "__main__":
    os.environ.setdefault("WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))
    args = parse_args()
    args = validate_args(args)
    nmb = args.global_batch_size // (args.micro_batch_size * max(1, args.data_parallel_size))
    report_theoretical_memory(args, num_microbatches=nmb, verbose=True)
