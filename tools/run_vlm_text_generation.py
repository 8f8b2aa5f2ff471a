#!/usr/bin/env python
"""Generate text from image + prompt pairs with a LLaVA-style model (reference
``examples/multimodal/run_text_generation.py`` / ``tools/run_vlm_text_generation.py``).

    python tools/run_vlm_text_generation.py <model args> --load ck --input-file samples.jsonl --output-file out.jsonl

``samples.jsonl``: ``{"image": "<path to .pt/.npy [3,H,W] tensor or an image file>", "prompt": "... <image> ..."}``.
Images are resized to ``--img-h x --img-w`` (bilinear) and normalised with the CLIP statistics.

The reference tool's directory form is also accepted: ``--input-path DIR --output-path out.jsonl [--num-partitions N
--partition-id I] [--gt-path coco_annotations.json] [--drop-vision-class-token]`` captions every image of ``DIR`` (sorted;
partition ``I`` of ``N`` equal parts) with the fixed COCO prompt and writes ``{"question_id", "prompt", "caption",
"ground_truth"?}`` lines; the question id is the number at the end of the file name (COCO naming) or the file's index.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from megatronapp_b200.inference import SamplingParams, StaticInferenceEngine, TextGenerationController, VLMInferenceWrapper  # noqa: E402
from megatronapp_b200.training import checkpointing  # noqa: E402
from megatronapp_b200.training.global_vars import get_args, get_tokenizer  # noqa: E402
from megatronapp_b200.training.initialize import initialize_megatron  # noqa: E402
from megatronapp_b200.training.training import get_model  # noqa: E402
from pretrain_vlm import add_vlm_extra_args as add_vlm_args, model_provider  # noqa: E402

CLIP_MEAN, CLIP_STD = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)


def load_image(path: str, h: int, w: int) -> torch.Tensor:
    if path.endswith(".pt"):
        img = torch.load(path).float()
    elif path.endswith(".npy"):
        import numpy as np

        img = torch.from_numpy(np.load(path)).float()
    else:
        from PIL import Image  # optional dependency
        import numpy as np

        img = torch.from_numpy(np.asarray(Image.open(path).convert("RGB"))).permute(2, 0, 1).float() / 255.0
    img = torch.nn.functional.interpolate(img.unsqueeze(0), size=(h, w), mode="bilinear", align_corners=False)[0]
    mean, std = torch.tensor(CLIP_MEAN).view(3, 1, 1), torch.tensor(CLIP_STD).view(3, 1, 1)
    return (img - mean) / std


def _extra(parser):
    parser = add_vlm_args(parser)
    g = parser.add_argument_group("vlm generation")
    g.add_argument("--input-file", default=None)
    g.add_argument("--output-file", default=None)
    g.add_argument("--input-path", default=None, help="directory of images (reference CLI)")
    g.add_argument("--output-path", default=None)
    g.add_argument("--num-partitions", type=int, default=0)
    g.add_argument("--partition-id", type=int, default=0)
    g.add_argument("--gt-path", default=None, help="COCO-style annotations json: captions are copied into the output")
    g.add_argument("--prompt", default=None, help="prompt of the directory form (default: the reference's COCO captioning prompt)")
    g.add_argument("--drop-vision-class-token", action="store_true", dest="disable_vision_class_token")
    g.add_argument("--out-seq-length", type=int, default=64)
    g.add_argument("--temperature", type=float, default=1.0)
    g.add_argument("--top_k", type=int, default=1)
    g.add_argument("--top_p", type=float, default=0.0)
    return parser


def main():
    initialize_megatron(extra_args_provider=_extra, args_defaults={"tokenizer_type": "NullMultimodalTokenizer", "micro_batch_size": 1})
    args = get_args()
    model = get_model(model_provider, wrap_with_ddp=False)
    if args.load:
        checkpointing.load_checkpoint(model, None, None)
    m = model[0]
    dev = next(m.parameters()).device
    wrapper = VLMInferenceWrapper(m, 1, args.inference_max_seq_length, args.padded_vocab_size)
    engine = StaticInferenceEngine(TextGenerationController(wrapper, get_tokenizer()), 1)
    sp = SamplingParams(temperature=args.temperature, top_k=args.top_k, top_p=args.top_p,
                        num_tokens_to_generate=args.out_seq_length)
    if args.input_path:
        return _caption_directory(args, wrapper, engine, sp, dev)
    if not (args.input_file and args.output_file):
        raise SystemExit("give --input-file / --output-file, or --input-path / --output-path")
    with open(args.input_file) as f, open(args.output_file, "w") as out:
        for line in f:
            if not line.strip():
                continue
            s = json.loads(line)
            wrapper.set_images(load_image(s["image"], args.img_h, args.img_w).unsqueeze(0).to(dev))
            r = engine.generate([s["prompt"]], sp)[0]
            out.write(json.dumps({"prompt": s["prompt"], "image": s["image"], "text": r.generated_text}) + "\n")
            out.flush()


COCO_PROMPT = "<image>Give a short and clear explanation of the subsequent image.\n"


def _caption_directory(args, wrapper, engine, sp, dev):
    import glob
    import re
    from collections import defaultdict

    files = sorted(glob.glob(os.path.join(args.input_path, "*")))
    if args.num_partitions > 0:
        per = len(files) // args.num_partitions
        files = files[per * args.partition_id: per * (args.partition_id + 1)]
    gt = defaultdict(list)
    if args.gt_path:
        with open(args.gt_path) as f:
            for a in json.load(f)["annotations"]:
                gt[int(a["image_id"])].append(a["caption"])
    if not args.output_path:
        raise SystemExit("--input-path needs --output-path")
    with open(args.output_path, "w") as out:
        for idx, path in enumerate(files):
            m = re.search(r"(\d+)\.[A-Za-z0-9]+$", os.path.basename(path))
            qid = int(m.group(1)) if m else idx
            wrapper.set_images(load_image(path, args.img_h, args.img_w).unsqueeze(0).to(dev))
            prompt = args.prompt or COCO_PROMPT
            r = engine.generate([prompt], sp)[0]
            rec = {"question_id": qid, "prompt": prompt, "caption": r.generated_text}
            if args.gt_path:
                rec["ground_truth"] = gt.get(qid, [])
            if args.rank == 0:
                out.write(json.dumps(rec) + "\n")
                out.flush()


if __name__ == "__main__":
    main()
