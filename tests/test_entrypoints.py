"""Every ``pretrain_*.py`` entry point runs a few iterations end to end (CPU, gloo, world size 1) through the real CLI:
argument parsing -> initialize -> model/optimizer setup -> data iterators -> train loop -> eval."""
import os
import re
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

COMMON = ("--num-layers 2 --hidden-size 32 --num-attention-heads 4 --seq-length 32 --max-position-embeddings 32 "
          "--micro-batch-size 2 --global-batch-size 4 --train-iters 3 --lr 1e-3 --vocab-size 200 --mock-data "
          "--log-interval 1 --eval-iters 1 --eval-interval 100 --distributed-backend gloo "
          "--no-gradient-accumulation-fusion").split()

CASES = {
    "gpt": ["pretrain_gpt.py"],
    "gpt_legacy": ["pretrain_gpt.py", "--use-legacy-models"],
    "gpt_rope_swiglu_gqa": ["pretrain_gpt.py", "--position-embedding-type", "rope", "--swiglu", "--group-query-attention",
                            "--num-query-groups", "2", "--normalization", "RMSNorm", "--disable-bias-linear",
                            "--untie-embeddings-and-output-weights"],
    "gpt_moe": ["pretrain_gpt.py", "--num-experts", "4", "--moe-router-topk", "2", "--moe-aux-loss-coeff", "0.01"],
    "gpt_mla_mtp": ["pretrain_gpt.py", "--multi-latent-attention", "--q-lora-rank", "16", "--kv-lora-rank", "16", "--qk-head-dim", "8",
                    "--qk-pos-emb-head-dim", "8", "--v-head-dim", "8", "--position-embedding-type", "rope", "--mtp-num-layers", "1"],
    "gpt_mxfp8_recompute": ["pretrain_gpt.py", "--fp8-format", "hybrid", "--fp8-recipe", "mxfp8", "--checkpoint-activations"],
    "gpt_heterogeneous": ["pretrain_gpt.py", "--heterogeneous-layers-config-encoded-json",
                          '{"block_configs": [{"attention": {"n_heads_in_group": 2}, "ffn": {"ffn_mult": 1.5}}, '
                          '{"attention": {"no_op": true}, "ffn": {"intermediate_size": 48}}]}'],
    "gpt_bf16_moments_offload": ["pretrain_gpt.py", "--use-precision-aware-optimizer", "--exp-avg-dtype", "bf16", "--exp-avg-sq-dtype",
                                 "bf16", "--optimizer-cpu-offload", "--optimizer-offload-fraction", "0.5"],
    "gpt_user_spec": ["pretrain_gpt.py", "--spec", "megatronapp_b200.models.gpt.gpt_layer_specs", "get_gpt_layer_local_spec"],
    "bert": ["pretrain_bert.py"],
    "bert_legacy": ["pretrain_bert.py", "--use-legacy-models"],
    "t5": ["pretrain_t5.py", "--encoder-seq-length", "32", "--decoder-seq-length", "16"],
    "t5_relative": ["pretrain_t5.py", "--encoder-seq-length", "32", "--decoder-seq-length", "16",
                    "--position-embedding-type", "relative"],
    "mamba_hybrid": ["pretrain_mamba.py", "--hybrid-override-pattern", "M*"],
    "vlm": ["pretrain_vlm.py", "--img-h", "14", "--img-w", "14", "--patch-dim", "7"],
    "vlm_radio": ["pretrain_vlm.py", "--img-h", "14", "--img-w", "14", "--patch-dim", "7", "--vision-model-type", "radio",
                  "--max-img-h", "28", "--max-img-w", "28"],
    "retro": ["pretrain_retro.py", "--retro-chunk-length", "8", "--retro-retrieved-length", "12"],
    "vision_classify": ["pretrain_vision_classify.py", "--img-h", "16", "--img-w", "16", "--patch-dim", "4", "--num-classes", "5"],
    "vision_dino": ["pretrain_vision_dino.py", "--img-h", "16", "--img-w", "16", "--patch-dim", "4", "--num-classes", "5",
                    "--dino-local-img-size", "8", "--dino-local-crops-number", "2", "--dino-out-dim", "64",
                    "--dino-head-hidden-size", "32", "--dino-bottleneck-size", "16", "--iter-per-epoch", "2"],
    "vision_dino_swin": ["pretrain_vision_dino.py", "--img-h", "224", "--img-w", "224", "--num-classes", "5",
                         "--vision-backbone-type", "swin", "--dino-local-img-size", "224", "--dino-local-crops-number", "0",
                         "--dino-out-dim", "64", "--dino-head-hidden-size", "32", "--dino-bottleneck-size", "16",
                         "--micro-batch-size", "1", "--global-batch-size", "1"],
    "vision_inpaint": ["pretrain_vision_inpaint.py", "--img-h", "16", "--img-w", "16", "--patch-dim", "4"],
    "vision_inpaint_mit": ["pretrain_vision_inpaint.py", "--img-h", "32", "--img-w", "32", "--patch-dim", "4",
                           "--vision-backbone-type", "mit", "--mask-type", "row"],
    "ict": ["pretrain_ict.py", "--retriever-report-topk-accuracies", "1", "2", "--retriever-score-scaling",
            "--biencoder-projection-dim", "16"],
}


def _port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("case", sorted(CASES))
def test_pretrain_entrypoint_runs(case):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_port()), RANK="0", WORLD_SIZE="1", LOCAL_RANK="0",
               CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2")
    script, *extra = CASES[case]
    p = subprocess.run([sys.executable, os.path.join(ROOT, script), *COMMON, *extra], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=600)
    out = p.stdout + p.stderr
    assert p.returncode == 0, out[-3000:]
    m = re.search(r"iteration\s+3/\s+3 \|.*?(?:lm loss|loss): ([0-9.E+\-]+)", out)
    assert m, out[-2000:]
    loss = float(m.group(1))
    assert loss == loss and 0.0 < loss < 50.0
    assert "validation loss at the end of training" in out


MULTI = {
    "mamba_tp2_sp": ["pretrain_mamba.py", "--hybrid-override-pattern", "M*", "--tensor-model-parallel-size", "2",
                     "--sequence-parallel", "--mamba-num-groups", "2", "--mamba-head-dim", "8", "--mamba-state-dim", "8"],
    "bert_tp2": ["pretrain_bert.py", "--tensor-model-parallel-size", "2"],
    "t5_tp2": ["pretrain_t5.py", "--encoder-seq-length", "32", "--decoder-seq-length", "16", "--tensor-model-parallel-size", "2"],
    "t5_pp2_split": ["pretrain_t5.py", "--encoder-seq-length", "32", "--decoder-seq-length", "16",
                     "--pipeline-model-parallel-size", "2"],
    "gpt_moe_ep2": ["pretrain_gpt.py", "--num-experts", "4", "--moe-router-topk", "2", "--expert-model-parallel-size", "2"],
    "gpt_tp2_sp": ["pretrain_gpt.py", "--tensor-model-parallel-size", "2", "--sequence-parallel"],
    "gpt_dpp_vpp_eval": ["pretrain_gpt.py", "--num-layers", "4", "--pipeline-model-parallel-size", "2",
                         "--num-layers-per-virtual-pipeline-stage", "1", "--use-dpp", "--eval-interval", "2"],
    "gpt_fbd": ["pretrain_gpt.py", "--forward-backward-disaggregating"],
    "gpt_torch_fsdp2": ["pretrain_gpt.py", "--use-torch-fsdp2"],
    "gpt_custom_fsdp_bf16": ["pretrain_gpt.py", "--use-custom-fsdp", "--bf16"],
    "gpt_trace_pp2": ["pretrain_gpt.py", "--pipeline-model-parallel-size", "2", "--trace", "--trace-interval", "1",
                      "--continuous-trace-iterations", "2"],
    "vlm_pp2": ["pretrain_vlm.py", "--img-h", "14", "--img-w", "14", "--patch-dim", "7", "--pipeline-model-parallel-size", "2"],
}


@pytest.mark.parametrize("case", sorted(MULTI))
def test_pretrain_entrypoint_two_ranks(case, tmp_path):
    """Two gloo ranks through torchrun: only TP rank 0 owns data iterators, so batches must be broadcast and every rank
    must agree on whether to evaluate; the T5 case exercises the encoder/decoder pipeline split."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    script, *extra = MULTI[case]
    if "--trace" in extra:
        extra += ["--trace-dir", str(tmp_path / "trace")]
    common = [("8" if a == "4" and COMMON[i - 1] == "--global-batch-size" else a) for i, a in enumerate(COMMON)]
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(_port()), os.path.join(ROOT, script), *common, *extra],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    out = p.stdout + p.stderr
    assert p.returncode == 0, out[-3000:]
    assert re.search(r"iteration\s+3/\s+3 \|", out) and "validation loss at the end of training" in out, out[-2000:]


def test_text_generation_server_two_tensor_parallel_ranks():
    """Rank 0 serves HTTP, rank 1 follows its announcements; both run the TP=2 forward passes of every request."""
    import json
    import signal
    import time
    import urllib.request

    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    port = _port()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_port()), os.path.join(ROOT, "tools", "run_text_generation_server.py"),
           "--tensor-model-parallel-size", "2", "--num-layers", "2", "--hidden-size", "32", "--num-attention-heads", "4",
           "--seq-length", "32", "--max-position-embeddings", "64", "--micro-batch-size", "1", "--vocab-size", "200",
           "--distributed-backend", "gloo", "--port", str(port), "--inference-max-seq-length", "64"]
    proc = subprocess.Popen(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                            start_new_session=True)
    try:
        def ask(prompts, n):
            body = json.dumps({"prompts": prompts, "tokens_to_generate": n, "top_k": 1}).encode()
            req = urllib.request.Request(f"http://127.0.0.1:{port}/api", data=body, method="PUT",
                                         headers={"Content-Type": "application/json"})
            return json.loads(urllib.request.urlopen(req, timeout=60).read())

        out = None
        for _ in range(90):
            if proc.poll() is not None:
                break
            try:
                out = ask(["5 6 7", "9 10"], 5)
                break
            except OSError:
                time.sleep(1.0)
        assert out is not None, proc.stdout.read()[-2000:] if proc.poll() is not None else "server never came up"
        assert [len(t) for t in out["tokens"]] == [8, 7]
        again = ask(["5 6 7"], 3)                                     # the worker rank keeps following
        assert again["tokens"][0] == out["tokens"][0][:6]
    finally:
        os.killpg(proc.pid, signal.SIGTERM)                          # exactly the process group started above
        proc.wait(timeout=30)


@pytest.mark.parametrize("fmt,model", [("torch_dist", "dense"), ("torch", "moe_ep2"), ("torch", "fsdp"), ("torch_dist", "fsdp")])
def test_checkpoint_resume_is_exact(tmp_path, fmt, model):
    """4 iterations straight == 2 iterations, save (distributed optimizer, dp=2), exit, resume for 2 more: same losses and
    gradient norms to the printed precision (weights, optimizer shards, LR schedule, data position, RNG all restored)."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    base = ("--num-layers 2 --hidden-size 32 --num-attention-heads 4 --seq-length 32 --max-position-embeddings 32 "
            "--micro-batch-size 2 --global-batch-size 8 --lr 1e-2 --vocab-size 200 --mock-data --log-interval 1 --eval-iters 1 "
            "--eval-interval 100 --distributed-backend gloo --no-gradient-accumulation-fusion --lr-decay-iters 100 "
            "--train-iters 4 --attention-dropout 0 --hidden-dropout 0").split()
    # dense: ZeRO-1 optimizer shards over dp=2;  moe_ep2: every rank owns two of the four experts (and must write them)
    # fsdp: ZeRO-3 -- parameters exist in full only while gathered, so saving / loading is a collective
    base += {"dense": ["--use-distributed-optimizer"], "fsdp": ["--use-custom-fsdp"],
             "moe_ep2": ["--num-experts", "4", "--moe-router-topk", "2", "--expert-model-parallel-size", "2"]}[model]

    def run(*extra):
        p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                            "--master-addr", "127.0.0.1", "--master-port", str(_port()), os.path.join(ROOT, "pretrain_gpt.py"),
                            *base, *extra], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
        out = p.stdout + p.stderr
        assert p.returncode == 0, out[-3000:]
        return re.findall(r"\] iteration\s+(\d+)/.*?lm loss: ([0-9.E+\-]+).*?grad norm: ([0-9.]+)", out)

    straight = run()
    assert [r[0] for r in straight] == ["1", "2", "3", "4"]
    ck = str(tmp_path / "ck")
    first = run("--exit-interval", "2", "--save", ck, "--ckpt-format", fmt)
    assert first == straight[:2]
    resumed = run("--load", ck, "--ckpt-format", fmt)
    assert resumed == straight[2:], (resumed, straight)


def test_optimizer_cpu_offload_matches_device_optimizer(tmp_path):
    """--optimizer-cpu-offload keeps 60 % of the master weights / Adam moments in host memory and steps them there:
    same losses and gradient norms as the all-device optimizer, before and after a checkpoint round trip."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", CUDA_VISIBLE_DEVICES="",
               OMP_NUM_THREADS="2")
    base = COMMON + ["--train-iters", "4", "--attention-dropout", "0", "--hidden-dropout", "0",
                                                     "--lr", "1e-2", "--lr-decay-iters", "100"]

    def run(*extra):
        p = subprocess.run([sys.executable, os.path.join(ROOT, "pretrain_gpt.py"), *base, *extra], cwd=ROOT,
                           env=dict(env, MASTER_PORT=str(_port())), capture_output=True, text=True, timeout=600)
        out = p.stdout + p.stderr
        assert p.returncode == 0, out[-3000:]
        return re.findall(r"\] iteration\s+(\d+)/.*?lm loss: ([0-9.E+\-]+).*?grad norm: ([0-9.]+)", out)

    ref = run()
    off = ["--optimizer-cpu-offload", "--optimizer-offload-fraction", "0.6"]
    assert run(*off) == ref and len(ref) == 4
    ck = str(tmp_path / "ck")
    assert run(*off, "--exit-interval", "2", "--save", ck) == ref[:2]
    assert run(*off, "--load", ck) == ref[2:]


def test_megascope_ws_server_two_tensor_parallel_ranks():
    """MegaScope capture over TP=2: the worker rank mirrors the capture hooks (their TP gathers are collectives), rank 0
    streams the compressed tensors of every requested kind back to the WebSocket client."""
    import json
    import signal
    import time

    from websockets.sync.client import connect

    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    port = _port()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_port()), os.path.join(ROOT, "tools", "run_text_generation_server.py"),
           "--tensor-model-parallel-size", "2", "--enable-ws-server", "--ws-port", str(port), "--num-layers", "2",
           "--hidden-size", "32", "--num-attention-heads", "4", "--seq-length", "32", "--max-position-embeddings", "64",
           "--micro-batch-size", "1", "--vocab-size", "200", "--distributed-backend", "gloo", "--inference-max-seq-length", "64"]
    proc = subprocess.Popen(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                            start_new_session=True)
    try:
        ws = None
        for _ in range(90):
            if proc.poll() is not None:
                break
            try:
                ws = connect(f"ws://127.0.0.1:{port}")
                break
            except OSError:
                time.sleep(1.0)
        assert ws is not None, "ws server never came up"
        ws.send(json.dumps({"type": "generate", "prompts": ["4 5 6"], "tokens_to_generate": 3, "top_k": 1,
                            "visualization_flags": {"MLP1_mat_mul": "true", "QKV_mat_mul": "true",
                                                    "RawAttentionScore_mat_mul": "true", "Result": "true"},
                            "compressor_config": {"MLP": {"pixels": 4, "method": "mean"}, "QKV": {"pixels": 6, "method": "mean"}}}))
        kinds, qkv = set(), None
        while True:
            m = json.loads(ws.recv(timeout=120))
            kinds.add((m["type"], m.get("update_type")))
            if m["type"] == "update" and m["update_type"] == 1 and qkv is None:
                qkv = m["result"]
            if m["type"] in ("finish", "error"):
                break
        assert m["type"] == "finish" and len(m["tokens"][0]) == 6
        assert {("update", 1), ("update", 2), ("update", 4), ("update", 6)} <= kinds
        flat = qkv
        while isinstance(flat[0], list):
            flat = flat[-1]
        assert len(flat) > 0 and all(isinstance(v, float) for v in flat)   # compressed QKV row assembled from both TP shards
        ws.close()
    finally:
        os.killpg(proc.pid, signal.SIGTERM)
        proc.wait(timeout=30)


def test_example_scripts_only_use_flags_that_exist():
    """Every ``--flag`` in the shipped launch scripts is accepted by the parser of the program it launches."""
    import glob

    sys.path.insert(0, ROOT)
    from megatronapp_b200.training.arguments import build_parser

    known = set(build_parser()._option_string_actions)
    declared = set()                                   # flags declared by the individual entry points / tools
    for f in glob.glob(os.path.join(ROOT, "*.py")) + glob.glob(os.path.join(ROOT, "tools", "**", "*.py"), recursive=True) \
            + glob.glob(os.path.join(ROOT, "tasks", "**", "*.py"), recursive=True) \
            + glob.glob(os.path.join(ROOT, "examples", "**", "*.py"), recursive=True):
        declared |= set(re.findall(r'add_argument\(\s*"(--[a-zA-Z0-9_\-]+)"', open(f).read()))
    launcher = {"--nnodes", "--nproc-per-node", "--master-addr", "--master-port", "--node-rank", "--standalone", "--local-addr"}
    scripts = glob.glob(os.path.join(ROOT, "examples", "**", "*.sh"), recursive=True) + glob.glob(os.path.join(ROOT, "scripts", "*.sh"))
    assert len(scripts) >= 15
    bad = {}
    for f in scripts:
        for flag in set(re.findall(r"(?<![\w-])(--[a-zA-Z][a-zA-Z0-9_\-]*)", open(f).read())):
            if flag not in known and flag not in declared and flag not in launcher:
                bad.setdefault(os.path.relpath(f, ROOT), []).append(flag)
    assert not bad, bad


def test_checkpoint_format_conversion_and_progress_log(tmp_path):
    """``--ckpt-convert-format`` re-saves a per-rank checkpoint as a sharded one (and stops); ``--log-progress`` writes the
    job start and every checkpoint with its throughput to ``<save>/progress.txt``."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", CUDA_VISIBLE_DEVICES="",
               OMP_NUM_THREADS="2")
    a, b = str(tmp_path / "a"), str(tmp_path / "b")

    def run(*extra):
        p = subprocess.run([sys.executable, os.path.join(ROOT, "pretrain_gpt.py"), *COMMON, *extra], cwd=ROOT,
                           env=dict(env, MASTER_PORT=str(_port())), capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, (p.stdout + p.stderr)[-3000:]
        return p.stdout + p.stderr

    run("--save", a, "--save-interval", "3", "--ckpt-format", "torch", "--log-progress")
    lines = open(os.path.join(a, "progress.txt")).read().strip().split("\n")
    assert "Starting job" in lines[0] and "Saved checkpoint\tIteration: 3" in lines[-1] and "TFLOP/s/GPU" in lines[-1]
    out = run("--load", a, "--ckpt-format", "torch", "--ckpt-convert-format", "torch_dist", "--ckpt-convert-save", b)
    assert "converted checkpoint" in out and "iteration        1/" not in out          # converts and stops, no training
    assert os.path.exists(os.path.join(b, "iter_0000003", "metadata.json"))
    out = run("--load", b, "--ckpt-format", "torch_dist", "--train-iters", "4", "--override-opt_param-scheduler")
    assert "successfully loaded checkpoint" in out and re.search(r"iteration\s+4/\s+4 \|", out)


def test_non_persistent_checkpoints_cover_the_gap_between_persistent_ones(tmp_path):
    """Persistent checkpoint every 4 iterations, a non-persistent ("global") one every iteration: after a crash at
    iteration 5 the job resumes from the non-persistent iteration-5 state, not from iteration 4."""
    import shutil

    env = dict(os.environ, MASTER_ADDR="127.0.0.1", RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", CUDA_VISIBLE_DEVICES="",
               OMP_NUM_THREADS="2")
    base = COMMON + ["--train-iters", "6", "--attention-dropout", "0", "--hidden-dropout", "0", "--lr", "1e-2",
                     "--lr-decay-iters", "100"]
    ck = str(tmp_path / "ck")

    def run(*extra):
        p = subprocess.run([sys.executable, os.path.join(ROOT, "pretrain_gpt.py"), *base, *extra], cwd=ROOT,
                           env=dict(env, MASTER_PORT=str(_port())), capture_output=True, text=True, timeout=600)
        out = p.stdout + p.stderr
        assert p.returncode == 0, out[-3000:]
        return out, re.findall(r"\] iteration\s+(\d+)/.*?lm loss: ([0-9.E+\-]+)", out)

    _, straight = run()
    run("--save", ck, "--save-interval", "4", "--non-persistent-save-interval", "1", "--non-persistent-ckpt-type", "global",
        "--exit-interval", "5")
    assert sorted(os.listdir(os.path.join(ck, "non_persistent"))) == ["iter_0000005", "latest_checkpointed_iteration.txt"]
    shutil.rmtree(os.path.join(ck, "iter_0000005"))                  # "crash": the exit checkpoint never made it to disk
    open(os.path.join(ck, "latest_checkpointed_iteration.txt"), "w").write("4")
    out, resumed = run("--load", ck, "--non-persistent-ckpt-type", "global")
    assert len(straight) == 6 and "non_persistent at iteration 5" in out and resumed == straight[5:]


def test_local_non_persistent_checkpoint_resumes_exactly(tmp_path):
    """``--non-persistent-ckpt-type local``: per-rank state on node-local storage every iteration; after a crash at
    iteration 5 (last persistent checkpoint: 4) the run resumes from the local iteration-5 state and reproduces the
    uninterrupted losses."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", CUDA_VISIBLE_DEVICES="",
               OMP_NUM_THREADS="2")
    base = COMMON + ["--train-iters", "6", "--attention-dropout", "0", "--hidden-dropout", "0", "--lr", "1e-2",
                     "--lr-decay-iters", "100"]
    ck, loc = str(tmp_path / "ck"), str(tmp_path / "local_ssd")

    def run(*extra):
        p = subprocess.run([sys.executable, os.path.join(ROOT, "pretrain_gpt.py"), *base, *extra], cwd=ROOT,
                           env=dict(env, MASTER_PORT=str(_port())), capture_output=True, text=True, timeout=600)
        out = p.stdout + p.stderr
        assert p.returncode == 0, out[-3000:]
        return out, re.findall(r"\] iteration\s+(\d+)/.*?lm loss: ([0-9.E+\-]+)", out)

    _, straight = run()
    np_args = ["--non-persistent-save-interval", "1", "--non-persistent-ckpt-type", "local", "--non-persistent-local-ckpt-dir", loc]
    run("--save", ck, "--save-interval", "4", *np_args, "--exit-interval", "5")
    assert os.listdir(loc) == ["iter_0000005"]
    import shutil

    shutil.rmtree(os.path.join(ck, "iter_0000005"), ignore_errors=True)   # "crash": the exit checkpoint never made it to disk
    open(os.path.join(ck, "latest_checkpointed_iteration.txt"), "w").write("4")
    out, resumed = run("--load", ck, *np_args)
    assert len(straight) == 6 and "(local) at iteration 5" in out and resumed == straight[5:]


def test_t5_trains_on_a_preprocessed_sentence_corpus(tmp_path):
    """jsonl -> ``tools/preprocess_data.py --split-sentences`` -> ``pretrain_t5.py --data-path`` (span corruption with the
    tokenizer's extra ids as sentinels) -- the whole real-data path through the CLIs."""
    import json
    import random

    random.seed(0)
    words = ["the", "cat", "dog", "sat", "ran"] + list("abcdefghij")
    vocab = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + words + ["."]
    (tmp_path / "vocab.txt").write_text("\n".join(vocab) + "\n")
    with open(tmp_path / "c.jsonl", "w") as f:
        for _ in range(60):
            f.write(json.dumps({"text": ". ".join(" ".join(random.choice(words) for _ in range(random.randint(4, 9)))
                                                   for _ in range(random.randint(3, 6)))}) + "\n")
    tok = ["--tokenizer-type", "BertWordPieceLowerCase", "--vocab-file", str(tmp_path / "vocab.txt"), "--vocab-extra-ids", "10"]
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "preprocess_data.py"), "--input", str(tmp_path / "c.jsonl"),
                        "--split-sentences", "--output-prefix", str(tmp_path / "c"), "--workers", "1", *tok],
                       cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_port()), RANK="0", WORLD_SIZE="1", LOCAL_RANK="0",
               CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2")
    common = [a for a in COMMON if a not in ("--mock-data",)]
    i = common.index("--vocab-size")
    common = common[:i] + common[i + 2:]
    p = subprocess.run([sys.executable, os.path.join(ROOT, "pretrain_t5.py"), *common, *tok, "--encoder-seq-length", "32",
                        "--decoder-seq-length", "16", "--data-path", str(tmp_path / "c_text_sentence"), "--split", "90,5,5"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    out = p.stdout + p.stderr
    assert p.returncode == 0, out[-3000:]
    assert re.search(r"iteration\s+3/\s+3 \|.*lm loss: [0-9.E+\-]+", out) and "validation loss at the end of training" in out


@pytest.mark.parametrize("family", ["gpt", "bert"])
def test_shipped_sample_corpus_preprocesses_and_trains(tmp_path, family):
    """The in-tree sample corpora (``datasets_gpt/``: jsonl + byte-level BPE vocab / merges; ``datasets_bert/``: pre-split
    sentences + WordPiece vocab -- generated by ``tools/make_sample_corpus.py``) go through ``tools/preprocess_data.py`` and
    then train, like the reference README's quick-start on its own ``datasets_*`` folders."""
    import json

    if family == "gpt":
        tok = ["--tokenizer-type", "GPT2BPETokenizer", "--vocab-file", os.path.join(ROOT, "datasets_gpt", "vocab.json"),
               "--merge-file", os.path.join(ROOT, "datasets_gpt", "merges.txt")]
        src, extra, prefix, script = os.path.join(ROOT, "datasets_gpt", "dataset.json"), ["--append-eod"], "c_text_document", "pretrain_gpt.py"
    else:
        tok = ["--tokenizer-type", "BertWordPieceLowerCase", "--vocab-file", os.path.join(ROOT, "datasets_bert", "vocab.txt")]
        src, extra, prefix, script = os.path.join(ROOT, "datasets_bert", "dataset_ss.json"), ["--split-sentences"], "c_text_sentence", "pretrain_bert.py"
    small = tmp_path / "c.jsonl"
    with open(src) as f, open(small, "w") as g:                     # the first 150 documents keep the CPU test short
        for _, line in zip(range(150), f):
            d = json.loads(line)
            g.write(json.dumps({"text": " ".join(d["text"]) if isinstance(d["text"], list) else d["text"]}) + "\n")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "preprocess_data.py"), "--input", str(small),
                        "--output-prefix", str(tmp_path / "c"), "--workers", "1", *extra, *tok],
                       cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_port()), RANK="0", WORLD_SIZE="1", LOCAL_RANK="0",
               CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2")
    common = [a for a in COMMON if a != "--mock-data"]
    i = common.index("--vocab-size")
    common = common[:i] + common[i + 2:]
    p = subprocess.run([sys.executable, os.path.join(ROOT, script), *common, *tok, "--data-path", str(tmp_path / prefix),
                        "--split", "90,5,5"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    out = p.stdout + p.stderr
    assert p.returncode == 0, out[-3000:]
    assert re.search(r"iteration\s+3/\s+3 \|.*lm loss: [0-9.E+\-]+", out) and "validation loss at the end of training" in out


@pytest.mark.parametrize("script", ["gpt_static_inference.py", "gpt_dynamic_inference.py"])
def test_inference_examples_generate(script, tmp_path):
    """examples/inference/gpt: static batch and continuous batching (Poisson arrivals) through the CLI, random weights."""
    import json

    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_port()), RANK="0", WORLD_SIZE="1", LOCAL_RANK="0",
               CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2")
    out_path = str(tmp_path / "gen.json")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "inference", "gpt", script), "--num-layers", "2", "--hidden-size", "32",
                        "--num-attention-heads", "4", "--seq-length", "64", "--max-position-embeddings", "64", "--micro-batch-size", "1",
                        "--tokenizer-type", "NullTokenizer", "--vocab-size", "127", "--distributed-backend", "gloo", "--prompts", "5 9 23 7",
                        "11 3", "--num-synthetic-prompts", "3", "--synthetic-prompt-length", "6", "--num-tokens-to-generate", "5",
                        "--incoming-requests-per-sec", "50", "--inference-dynamic-batching-buffer-size-gb", "0.001", "--output-path", out_path],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, (p.stdout + p.stderr)[-3000:]
    rows = json.load(open(out_path))
    assert len(rows) == 5 and all(len(r["tokens"]) == 5 for r in rows)


def test_trtllm_export_examples(tmp_path):
    """single-device export to an inference layout (tp 2) and the on-device distributed export on 2 gloo ranks"""
    import json

    single, distd = str(tmp_path / "single"), str(tmp_path / "dist")
    ex = os.path.join(ROOT, "examples", "export", "trtllm_export")
    p = subprocess.run([sys.executable, os.path.join(ex, "single_device_export", "gpt_single_device_cpu_export.py"), "--out", single,
                        "--export-tp", "2"], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, (p.stdout + p.stderr)[-2000:]
    conf = json.load(open(os.path.join(single, "config.json")))
    assert conf["mapping"]["tp_size"] == 2 and sorted(f for f in os.listdir(single) if f.startswith("rank")) == ["rank0.safetensors", "rank1.safetensors"]
    assert json.load(open(os.path.join(single, "build_config.json")))["plugin_config"]["paged_kv_cache"] is True
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
                        str(_port()), os.path.join(ex, "distributed_export", "gpt_distributed_gpu_export.py"), "--tensor-model-parallel-size", "2",
                        "--out", distd], cwd=str(tmp_path), env=dict(os.environ, CUDA_VISIBLE_DEVICES=""), capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, (p.stdout + p.stderr)[-2000:]
    import torch

    from safetensors.torch import load_file

    a, b = load_file(os.path.join(single, "rank1.safetensors")), torch.load(os.path.join(distd, "rank1.pt"))
    assert sorted(a) == sorted(b) and all(a[k].shape == b[k].shape for k in a)


@pytest.mark.parametrize("script,extra", [("pretrain_vision_classify.py", ["--num-classes", "3"]),
                                          ("pretrain_vision_inpaint.py", ["--mask-type", "row"]),
                                          ("pretrain_vision_dino.py", ["--dino-local-img-size", "8", "--dino-local-crops-number", "2", "--dino-out-dim", "32",
                                                                       "--dino-head-hidden-size", "32", "--dino-bottleneck-size", "16", "--iter-per-epoch", "2"])])
def test_vision_entrypoints_train_from_an_image_folder(tmp_path, script, extra):
    """``--data-path <dir>`` with ``train/<class>/*.png`` and ``val/...``: torchvision transforms (AutoAugment, multi-crop,
    patch masks) from ``legacy/data/vit_dataset.py`` feed the three vision pre-training entry points."""
    import numpy as np
    from PIL import Image

    rng = np.random.RandomState(0)
    for split in ("train", "val"):
        for ci, c in enumerate(("cat", "dog", "eel")):
            os.makedirs(tmp_path / split / c)
            for i in range(6):
                a = (rng.rand(24, 24, 3) * 60).astype(np.uint8)
                a[..., ci] += 150
                Image.fromarray(a).save(tmp_path / split / c / f"{i}.png")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_port()), RANK="0", WORLD_SIZE="1", LOCAL_RANK="0",
               CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2")
    common = [a for a in COMMON if a != "--mock-data"]
    p = subprocess.run([sys.executable, os.path.join(ROOT, script), *common, "--img-h", "16", "--img-w", "16", "--patch-dim", "4",
                        "--data-path", str(tmp_path), *extra], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    out = p.stdout + p.stderr
    assert p.returncode == 0, out[-3000:]
    assert re.search(r"iteration\s+3/\s+3 \|.*loss: [0-9.E+\-]+", out) and "validation loss at the end of training" in out


def test_knowledge_distillation_from_a_teacher_checkpoint(tmp_path):
    """--export-kd-teacher-load / --export-kd-cfg (reference post_training/model_provider.py + loss_func.py): a TP2 + SP student
    of another width learns from a frozen teacher checkpoint through the logit KL and one intermediate pair (with a trained
    projection); the teacher's architecture is read from model_config.yaml; --export-kd-finalize then resumes the plain student."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    common = [("8" if a == "4" and COMMON[i - 1] == "--global-batch-size" else a) for i, a in enumerate(COMMON)]
    common = [a for i, a in enumerate(common) if a not in ("--num-layers", "--hidden-size", "--num-attention-heads")
              and common[i - 1] not in ("--num-layers", "--hidden-size", "--num-attention-heads")]
    tp = ["--tensor-model-parallel-size", "2", "--sequence-parallel"]

    def run(*extra):
        p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                            "127.0.0.1", "--master-port", str(_port()), os.path.join(ROOT, "pretrain_gpt.py"), *common, *tp, *extra],
                           cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
        out = p.stdout + p.stderr
        assert p.returncode == 0, out[-3000:]
        return out

    teacher, student, final = (str(tmp_path / d) for d in ("teacher", "student", "final"))
    run("--num-layers", "2", "--hidden-size", "32", "--num-attention-heads", "4", "--save", teacher, "--save-interval", "3")
    with open(os.path.join(teacher, "model_config.yaml"), "w") as f:
        f.write("num_layers: 2\nhidden_size: 32\nnum_attention_heads: 4\nencoder_seq_length: 32\nbias: true\n")
    cfg = str(tmp_path / "kd.yaml")
    with open(cfg, "w") as f:
        f.write('logit_layers: ["output_layer", "output_layer"]\n'
                'intermediate_layer_pairs:\n  - ["decoder.final_layernorm", "decoder.layers.1.input_layernorm"]\n'
                "skip_lm_loss: false\nkd_loss_scale: 2.0\n")
    small = ["--num-layers", "1", "--hidden-size", "16", "--num-attention-heads", "2"]
    out = run(*small, "--export-kd-teacher-load", teacher, "--export-kd-cfg", cfg, "--save", student, "--save-interval", "3",
              "--ckpt-format", "torch")
    assert "building the distillation teacher (2 layers, hidden size 32)" in out
    kd = [float(x) for x in re.findall(r"kd loss: ([0-9.E+\-]+)", out)]
    assert len(kd) == 3 and all(0 < v < 1 for v in kd) and kd[-1] < kd[0], kd
    assert "kd loss" not in out.split("validation loss at the end of training")[1]        # evaluation reports the LM loss only
    sd = torch.load(os.path.join(student, "iter_0000003", "mp_rank_00", "model_optim_rng.pt"), weights_only=False)["model"]
    assert sd["kd_projections.0.fit.weight"].shape == (32, 16)
    out = run(*small, "--export-kd-teacher-load", teacher, "--export-kd-finalize", "--load", student, "--save", final,
              "--save-interval", "4", "--train-iters", "4", "--override-opt_param-scheduler", "--ckpt-format", "torch")
    assert re.search(r"iteration\s+4/\s+4 \|", out) and "kd loss" not in out
    sd = torch.load(os.path.join(final, "iter_0000004", "mp_rank_00", "model_optim_rng.pt"), weights_only=False)["model"]
    assert not [k for k in sd if k.startswith("kd_")]


def test_moe_upcycling_from_a_dense_checkpoint(tmp_path):
    """--moe-use-upcycling (reference training.py:1284-1311): a 4-expert EP2 model seeded from a trained dense checkpoint evaluates
    to the dense model's loss, the upcycled weights are saved as iteration 1, and repeating the flag on that directory is refused."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    common = [("8" if a == "4" and COMMON[i - 1] == "--global-batch-size" else a) for i, a in enumerate(COMMON)]
    common += ["--attention-dropout", "0", "--hidden-dropout", "0"]
    moe = ["--num-experts", "4", "--expert-model-parallel-size", "2", "--moe-router-topk", "2", "--moe-token-dispatcher-type", "alltoall"]

    def run(*extra, ok=True):
        p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                            "127.0.0.1", "--master-port", str(_port()), os.path.join(ROOT, "pretrain_gpt.py"), *common, *extra],
                           cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
        out = p.stdout + p.stderr
        assert (p.returncode == 0) == ok, out[-3000:]
        return out

    def val(out):
        return re.findall(r"validation loss at the end of training for val data \| lm loss value: ([0-9.E+\-]+)", out)

    dense, up = str(tmp_path / "dense"), str(tmp_path / "moe")
    ref = val(run("--save", dense, "--save-interval", "3"))
    out = run(*moe, "--moe-use-upcycling", "--load", dense, "--save", up, "--skip-train")
    assert "upcycled checkpoint saved" in out and val(out) == ref and len(ref) == 1
    assert os.path.isdir(os.path.join(up, "iter_0000001"))
    out = run(*moe, "--load", up, "--skip-train", "--finetune")                  # the saved MoE checkpoint reproduces it too
    assert val(out) == ref
    assert "only for the first" in run(*moe, "--moe-use-upcycling", "--load", dense, "--save", up, ok=False)


def test_dist_ckpt_strictness_modes(tmp_path):
    """--dist-ckpt-strictness: loading a tied-embedding checkpoint into an untied model misses output_layer.weight -- raise_all
    refuses, log_all reports and trains on, ignore_all trains on silently."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", CUDA_VISIBLE_DEVICES="",
               OMP_NUM_THREADS="2")

    def run(*extra):
        p = subprocess.run([sys.executable, os.path.join(ROOT, "pretrain_gpt.py"), *COMMON, *extra], cwd=ROOT,
                           env=dict(env, MASTER_PORT=str(_port())), capture_output=True, text=True, timeout=600)
        return p.returncode, p.stdout + p.stderr

    ck = str(tmp_path / "ck")
    assert run("--save", ck, "--save-interval", "3")[0] == 0
    untied = ["--untie-embeddings-and-output-weights", "--load", ck, "--finetune", "--no-load-optim", "--dist-ckpt-strictness"]
    rc, out = run(*untied, "raise_all")
    assert rc != 0 and "missing keys: ['output_layer.weight']" in out
    rc, out = run(*untied, "log_all")
    assert rc == 0 and "1 requested keys are not in the checkpoint" in out and re.search(r"iteration\s+3/\s+3 \|", out)
    rc, out = run(*untied, "ignore_all")
    assert rc == 0 and "requested keys are not" not in out and re.search(r"iteration\s+3/\s+3 \|", out)


def test_per_split_data_paths(tmp_path):
    """--train-data-path / --valid-data-path (each split from its own blend, used whole) and the file forms --data-args-path /
    --per-split-data-args-path (reference get_blend_and_blend_per_split)."""
    import json

    tok = ["--tokenizer-type", "GPT2BPETokenizer", "--vocab-file", os.path.join(ROOT, "datasets_gpt", "vocab.json"),
           "--merge-file", os.path.join(ROOT, "datasets_gpt", "merges.txt")]
    with open(os.path.join(ROOT, "datasets_gpt", "dataset.json")) as f:
        lines = [next(f) for _ in range(160)]
    for name, part in (("a", lines[:100]), ("b", lines[100:130]), ("c", lines[130:])):
        with open(tmp_path / f"{name}.jsonl", "w") as g:
            g.writelines(part)
        p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "preprocess_data.py"), "--input", str(tmp_path / f"{name}.jsonl"),
                            "--output-prefix", str(tmp_path / name), "--workers", "1", "--append-eod", *tok],
                           cwd=ROOT, capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr[-2000:]
    a, b, c = (str(tmp_path / f"{n}_text_document") for n in "abc")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2")
    common = [x for x in COMMON if x != "--mock-data"]
    i = common.index("--vocab-size")
    common = common[:i] + common[i + 2:]

    def run(*extra, ok=True):
        p = subprocess.run([sys.executable, os.path.join(ROOT, "pretrain_gpt.py"), *common, *tok, *extra], cwd=ROOT,
                           env=dict(env, MASTER_PORT=str(_port())), capture_output=True, text=True, timeout=600)
        out = p.stdout + p.stderr
        assert (p.returncode == 0) == ok, out[-3000:]
        return out

    def losses(out):
        return re.findall(r"lm loss: ([0-9.E+\-]+)", out), re.findall(r"for (val|test) data \| lm loss value: ([0-9.E+\-]+)", out)

    ref = losses(run("--train-data-path", "0.7", a, "0.3", c, "--valid-data-path", b))
    assert len(ref[0]) == 3 and [k for k, _ in ref[1]] == ["val"]                       # no test split given: no test evaluation
    cfg = tmp_path / "splits.json"
    json.dump({"train": ["0.7", a, "0.3", c], "valid": [b]}, open(cfg, "w"))
    assert losses(run("--per-split-data-args-path", str(cfg))) == ref
    blend = tmp_path / "blend.txt"
    blend.write_text(f"0.7 {a} 0.3 {c}\n")
    one = losses(run("--data-args-path", str(blend), "--split", "90,5,5"))
    assert len(one[0]) == 3 and one != ref
    assert "not both" in run("--data-path", a, "--valid-data-path", b, ok=False)


def test_tensorboard_and_wandb_logging(tmp_path):
    """--tensorboard-dir with --log-timers-to-tensorboard / --log-world-size-to-tensorboard, and --wandb-project (offline mode)."""
    pytest.importorskip("tensorboard")
    wandb_ok = True
    try:
        import wandb  # noqa: F401
    except ImportError:
        wandb_ok = False
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_port()), RANK="0", WORLD_SIZE="1", LOCAL_RANK="0",
               CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2", WANDB_MODE="offline", WANDB_SILENT="true")
    extra = ["--tensorboard-dir", str(tmp_path / "tb"), "--tensorboard-log-interval", "1", "--log-timers-to-tensorboard",
             "--log-world-size-to-tensorboard", "--timing-log-level", "1"]
    if wandb_ok:
        extra += ["--wandb-project", "proj", "--wandb-exp-name", "run", "--wandb-save-dir", str(tmp_path / "wb")]
    p = subprocess.run([sys.executable, os.path.join(ROOT, "pretrain_gpt.py"), *COMMON, *extra], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, (p.stdout + p.stderr)[-3000:]
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator

    acc = EventAccumulator(str(tmp_path / "tb"))
    acc.Reload()
    tags = set(acc.Tags()["scalars"])
    assert {"learning-rate", "lm loss", "world-size", "iteration-time", "optimizer-time"} <= tags, sorted(tags)
    assert [e.step for e in acc.Scalars("lm loss")] == [1, 2, 3]
    if wandb_ok:
        runs = [d for d in os.listdir(tmp_path / "wb" / "wandb") if d.startswith("offline-run")]
        assert len(runs) == 1 and any(f.endswith(".wandb") for f in os.listdir(tmp_path / "wb" / "wandb" / runs[0]))


def test_ict_starts_from_a_pretrained_bert(tmp_path):
    """--bert-load (reference biencoder_model.py:init_state_dict_from_bert): on iteration 0 both towers take embedding + encoder from
    a BERT checkpoint (any checkpoint format), and the context tower copies the query tower's projection."""
    import glob

    env = dict(os.environ, MASTER_ADDR="127.0.0.1", RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2")

    def run(script, *extra):
        p = subprocess.run([sys.executable, os.path.join(ROOT, script), *COMMON, *extra], cwd=ROOT,
                           env=dict(env, MASTER_PORT=str(_port())), capture_output=True, text=True, timeout=600)
        out = p.stdout + p.stderr
        assert p.returncode == 0, out[-3000:]
        return out

    bert, ict = str(tmp_path / "bert"), str(tmp_path / "ict")
    run("pretrain_bert.py", "--save", bert, "--save-interval", "3")
    out = run("pretrain_ict.py", "--train-iters", "1", "--lr", "0", "--biencoder-projection-dim", "16", "--bert-load", bert, "--save", ict,
              "--save-interval", "1", "--ckpt-format", "torch")
    assert "initializing the bi-encoder from the pretrained BERT model" in out
    from megatronapp_b200.ckpt import serialization

    ref = serialization.load_plain_tensors(glob.glob(os.path.join(bert, "iter_*"))[0])
    sd = torch.load(os.path.join(ict, "iter_0000001", "mp_rank_00", "model_optim_rng.pt"), weights_only=False)["model"]
    q, c = sd["query_model"], sd["context_model"]
    for k in ("embedding.word_embeddings.weight", "embedding.position_embeddings.weight"):
        assert torch.equal(q[k].float(), ref[k].float()) and torch.equal(c[k].float(), ref[k].float())
    layer0 = [k for k in q if k.startswith("encoder.layers.0.") and k.endswith("linear_fc1.weight")]
    assert layer0 and torch.equal(q[layer0[0]], c[layer0[0]])
    assert torch.equal(q["projection_enc.weight"], c["projection_enc.weight"])


@pytest.mark.parametrize("fmt", ["torch", "torch_dist"])
def test_retro_model_initialises_from_a_gpt_checkpoint(tmp_path, fmt):
    """--retro-add-retriever makes checkpoint loading non-strict (reference checkpointing.py:1455): the decoder of a Retro model
    starts from a GPT checkpoint, the retriever keeps its initialisation; without the flag the missing keys are an error."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2")

    def run(script, *extra):
        p = subprocess.run([sys.executable, os.path.join(ROOT, script), *COMMON, "--ckpt-format", fmt, *extra], cwd=ROOT,
                           env=dict(env, MASTER_PORT=str(_port())), capture_output=True, text=True, timeout=600)
        return p.returncode, p.stdout + p.stderr

    gpt = str(tmp_path / "gpt")
    assert run("pretrain_gpt.py", "--save", gpt, "--save-interval", "3")[0] == 0
    retro = ["--retro-chunk-length", "8", "--retro-retrieved-length", "12", "--load", gpt, "--finetune", "--no-load-optim"]
    rc, out = run("pretrain_retro.py", *retro, "--retro-add-retriever")
    assert rc == 0 and "successfully loaded checkpoint" in out and re.search(r"iteration\s+3/\s+3 \|", out), out[-2000:]
    rc, out = run("pretrain_retro.py", *retro)
    assert rc != 0 and ("Missing key" in out or "missing" in out.lower()), out[-2000:]
