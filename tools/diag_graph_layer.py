
import torch
from megatronapp_b200.config import TransformerConfig
from megatronapp_b200.models.gpt import get_gpt_layer_local_spec
from megatronapp_b200.parallel import state as ps
from megatronapp_b200.parallel.random import model_parallel_cuda_manual_seed
from megatronapp_b200.transformer.cuda_graphs import CudaGraphManager
from megatronapp_b200.transformer.spec_utils import build_module

ps.initialize_model_parallel(1, 1)
model_parallel_cuda_manual_seed(1)
cfg = TransformerConfig(num_layers=1, hidden_size=512, num_attention_heads=4, bf16=True, params_dtype=torch.bfloat16,
                        hidden_dropout=0.0, attention_dropout=0.0)
layer = build_module(get_gpt_layer_local_spec(), config=cfg, layer_number=1).cuda()
mgr = CudaGraphManager(layer, warmup_steps=1)
outs = []
for i in range(4):
    torch.manual_seed(i)
    x = torch.randn(256, 2, 512, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    y = mgr(x)
    y.float().pow(2).sum().backward()
    xe = x.detach().clone().requires_grad_(True)
    ye = layer(xe)[0]
    ye.float().pow(2).sum().backward()
    outs.append((float((y - ye).abs().max()), float((x.grad - xe.grad).abs().max())))
    del x, y, xe, ye      # nothing from the eager pass may stay alive across the (lazy) capture of the next call
assert mgr.replays >= 2 and mgr.refused is None, (mgr.replays, mgr.refused)
assert max(a for a, _ in outs) < 2e-2 and max(b for _, b in outs) < 2e-1, outs
print("GRAPH_OK", mgr.replays)
