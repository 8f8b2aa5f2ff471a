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
