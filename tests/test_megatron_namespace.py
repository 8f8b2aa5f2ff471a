"""The ``megatron`` compatibility namespace (``compat/`` on PYTHONPATH) resolves the reference's import paths."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "compat"))


def test_reference_import_paths_resolve():
    import megatron  # noqa: F401
    from megatron.core import mpu, parallel_state, tensor_parallel
    from megatron.core.models.gpt import GPTModel
    from megatron.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron.core.transformer import TransformerConfig
    from megatron.core.transformer.spec_utils import ModuleSpec
    from megatron.training import get_args, pretrain
    from megatron.training.arguments import core_transformer_config_from_args

    import megatronapp_b200.parallel.state as ps
    from megatronapp_b200.models.gpt.gpt_model import GPTModel as Native

    assert mpu is ps and parallel_state is ps
    assert GPTModel is Native
    assert hasattr(tensor_parallel, "ColumnParallelLinear") and hasattr(tensor_parallel, "vocab_parallel_cross_entropy")
    assert hasattr(tensor_parallel, "model_parallel_cuda_manual_seed")
    assert callable(get_args) and callable(pretrain) and callable(core_transformer_config_from_args)
    assert callable(get_gpt_layer_local_spec) and TransformerConfig.__name__ == "TransformerConfig" and ModuleSpec


def test_core_distributed_and_optimizer_paths():
    from megatron.core.distributed import DistributedDataParallel, finalize_model_grads
    from megatron.core.optimizer import OptimizerConfig, get_megatron_optimizer
    from megatron.core.pipeline_parallel import get_forward_backward_func
    from megatron.core.tensor_parallel.mappings import gather_from_sequence_parallel_region

    assert DistributedDataParallel and finalize_model_grads and OptimizerConfig and get_megatron_optimizer
    assert callable(get_forward_backward_func) and callable(gather_from_sequence_parallel_region)


def test_namespace_covers_datasets_inference_checkpointing_and_legacy():
    import subprocess
    import sys

    code = """
from megatron.core import dist_checkpointing
from megatron.core.datasets.gpt_dataset import GPTDatasetConfig, MockGPTDataset
from megatron.core.datasets.blended_megatron_dataset_builder import BlendedMegatronDatasetBuilder
from megatron.core.inference.engines import StaticInferenceEngine, DynamicInferenceEngine
from megatron.core.inference.sampling_params import SamplingParams
from megatron.core.packed_seq_params import PackedSeqParams
from megatron.core.models.vision.radio import RADIOViTModel
from megatron.core.models.multimodal.llava_model import LLaVAModel
from megatron.legacy.model import GPTModel, BertModel
from megatron.training.tokenizer import build_tokenizer
import shm_tensor_new_rdma_pre_alloc as shm
assert callable(dist_checkpointing.save) and callable(dist_checkpointing.load) and callable(shm.put_forward_tensor)
print("ok")
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(root, "compat"), root]))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0 and p.stdout.strip().endswith("ok"), p.stderr[-2000:]


def test_function_level_api_of_the_reference_resolves():
    """Names that user scripts import from the reference (``megatron.core.utils``, enums, clip_grads, fusions, TE layer names,
    text-generation helpers, training utilities) resolve through ``compat/`` and behave."""
    import subprocess
    import sys

    code = r'''
import sys
sys.path[:0] = [sys.argv[1] + "/compat", sys.argv[1]]
import argparse
import torch
import megatron
from megatron.core.utils import get_model_config, divide, make_viewless_tensor, StragglerDetector, init_method_normal, get_attr_wrapped_model
from megatron.core.enums import ModelType
from megatron.core.transformer.enums import AttnMaskType, AttnType, AttnBackend
from megatron.core.transformer.utils import attention_mask_func, make_sharded_tensors_for_checkpoint, get_linear_layer
from megatron.core.optimizer.clip_grads import get_grad_norm_fp32, clip_grad_by_total_norm_fp32, count_zeros_fp32
from megatron.core.tensor_parallel.utils import split_tensor_along_last_dim, VocabUtility
from megatron.core.tensor_parallel.layers import linear_with_grad_accumulation_and_async_allreduce, copy_tensor_model_parallel_attributes
from megatron.core.tensor_parallel.random import initialize_rng_tracker, get_all_rng_states
from megatron.core.fusions.fused_bias_dropout import get_bias_dropout_add
from megatron.core.fusions.fused_bias_gelu import bias_gelu_impl
from megatron.core.fusions.fused_bias_swiglu import bias_swiglu_impl
from megatron.core.fusions.fused_softmax import FusedScaleMaskSoftmax
from megatron.core.extensions.transformer_engine import TENorm, TEDotProductAttention, TELayerNormColumnParallelLinear, TEColumnParallelLinear
from megatron.core.inference_params import InferenceParams
from megatron.core.num_microbatches_calculator import reconfigure_num_microbatches_calculator, get_num_microbatches
from megatron.core.pipeline_parallel.p2p_communication import send_forward_recv_forward, send_forward_backward_recv_forward_backward
from megatron.core.transformer.moe.moe_utils import get_updated_expert_bias, sinkhorn
from megatron.core.transformer.module import fp32_to_float16, float16_to_fp32
from megatron.inference.text_generation.sampling import sample
from megatron.inference.text_generation.tokenization import tokenize_prompts, detokenize_generations
from megatron.inference.text_generation.communication import broadcast_int_list
from megatron.training.utils import get_ltor_masks_and_position_ids, is_rank0, get_blend_and_blend_per_split
from megatron.training.checkpointing import checkpoint_exists, find_checkpoint_rank_0, CheckpointType
from megatron.training.arguments import add_megatron_arguments, moe_freq_type
from megatron.training.training import print_datetime, update_train_iters
from megatron.training.global_vars import get_one_logger, rebuild_tokenizer
from megatron.legacy.model.utils import openai_gelu
assert str(ModelType.encoder_and_decoder) == "encoder_and_decoder" and AttnMaskType.causal.value == 2 and divide(8, 4) == 2
p = torch.nn.Parameter(torch.ones(4)); p.grad = torch.tensor([3.0, 0.0, 4.0, 0.0])
assert abs(get_grad_norm_fp32([p.grad]) - 5.0) < 1e-6 and count_zeros_fp32([p]) == 2.0
clip_grad_by_total_norm_fp32([p], 1.0, 5.0)
assert torch.allclose(p.grad, torch.tensor([0.6, 0.0, 0.8, 0.0]), atol=1e-5)
a = add_megatron_arguments(argparse.ArgumentParser(conflict_handler="resolve")).parse_args(["--num-layers", "3"])
assert a.num_layers == 3 and moe_freq_type("([1]*2+[0])") == [1, 1, 0]
assert sample(torch.tensor([[1.0, 3.0, 2.0]]), top_k=1).item() == 1
x = torch.randn(4, 8)
assert torch.allclose(float16_to_fp32(fp32_to_float16((x,), lambda t: t.bfloat16()))[0], x, atol=0.05)
assert torch.equal(get_updated_expert_bias(torch.tensor([4.0, 1.0, 1.0]), torch.zeros(3), 0.5), torch.tensor([-0.5, 0.5, 0.5]))
print("API-OK")
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-c", code, root], capture_output=True, text=True, timeout=300, cwd="/tmp")
    assert p.returncode == 0 and "API-OK" in p.stdout, (p.stdout + p.stderr)[-3000:]


def test_compat_coverage_report():
    """tools/compat_coverage.py against the reference tree (skipped when it is not mounted): most module paths resolve."""
    import subprocess
    import sys

    import pytest

    if not os.path.isdir("/root/reference/megatron"):
        pytest.skip("reference tree not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "compat_coverage.py"), "/root/reference"], capture_output=True,
                       text=True, timeout=600, cwd="/tmp")
    assert p.returncode == 0, p.stderr[-2000:]
    import json

    summary = json.loads(p.stdout.splitlines()[0])
    assert summary["modules_resolved"] >= 320 and summary["public_names_available"] >= 760, summary


def test_dist_checkpointing_dict_helpers():
    """core_api.dist_ckpt_utils (reference dict_utils / utils / mapping): extraction by type, prefix rewriting, factories."""
    import torch

    from megatronapp_b200.ckpt.sharded import ShardedTensor
    from megatronapp_b200.core_api import dist_ckpt_utils as u

    t = lambda k: ShardedTensor.from_rank_offsets(k, torch.zeros(2, 2))  # noqa: E731
    sd = {"a": t("layer.w"), "b": {"c": t("layer.b"), "n": 3}, "l": [t("x.0"), 5]}
    yes, no = u.extract_sharded_tensors(sd)
    assert sorted(yes) == ["a", "b", "l"] and no == {"b": {"n": 3}, "l": [5]} and len(yes["l"]) == 1
    u.add_prefix_for_sharding(sd, "model.")
    u.apply_prefix_mapping(sd, {"model.layer.": "m.L."})
    assert (sd["a"].key, sd["b"]["c"].key, sd["l"][0].key) == ("m.L.w", "m.L.b", "model.x.0")
    assert u.diff({"a": 1, "b": [1, 2]}, {"a": 2, "c": 0, "b": [1, 3]})[1] == [("c",)]
    assert u.merge({"a": {"x": 1}}, {"a": {"y": 2}}) == {"a": {"x": 1, "y": 2}}
    f = u.ShardedTensorFactory("w", torch.arange(4.0), lambda k, d, r, fr: {"lo": t(k + ".lo"), "hi": t(k + ".hi")},
                               lambda sub: torch.cat([sub["lo"], sub["hi"]]))
    s2 = {"w": f}
    u.apply_factories(s2)
    assert sorted(s2["w"]) == ["hi", "lo"] and s2["w"]["lo"].key == "w.lo"
    merged = u.apply_factory_merges({"w": {"lo": torch.zeros(1), "hi": torch.ones(1)}}, {"w": f})
    assert merged["w"].tolist() == [0.0, 1.0]
    assert u.is_main_replica((0, 0, 0)) and not u.is_main_replica((0, 1, 0))
    import pytest

    with pytest.raises(ValueError):
        list(u.zip_strict([1], [1, 2]))
