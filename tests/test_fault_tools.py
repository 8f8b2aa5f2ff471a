"""Rerun state machine (fault classification + injection) and straggler detector."""
import math

from megatronapp_b200.training.rerun_state_machine import (EXIT_CODE_FAILED_ON_RESULT_VALIDATION,
                                                          EXIT_CODE_RESUME_TO_DISAMBIGUATE, RerunDataIterator,
                                                          RerunErrorInjector, RerunMode, RerunStateMachine)
from megatronapp_b200.utils.straggler import StragglerDetector


def _run_step(rsm, it, compute):
    runs = 0
    while rsm.should_run_forward_backward(it):
        batch = next(it)
        runs += 1
        rsm.validate_result(compute(batch, runs), lambda v: math.isnan(v) or v > 100, "loss check")
    return runs


def test_disabled_mode_runs_once():
    rsm = RerunStateMachine(RerunMode.DISABLED)
    it = RerunDataIterator(iter(range(10)))
    assert _run_step(rsm, it, lambda b, r: 1.0) == 1 and next(it) == 1
    try:                                   # reruns off: a rejected result still stops the job (reference behaviour)
        _run_step(RerunStateMachine(RerunMode.DISABLED), RerunDataIterator(iter(range(10))), lambda b, r: float("nan"))
        raise AssertionError("a NaN must be fatal")
    except RuntimeError as e:
        assert "loss check" in str(e)


def test_transient_error_is_rerun_on_same_data_and_classified():
    rsm = RerunStateMachine(RerunMode.VALIDATE_RESULTS)
    it = RerunDataIterator(iter(range(10)))
    seen = []
    assert _run_step(rsm, it, lambda b, r: (seen.append(b), float("nan") if r == 1 else 1.0)[1]) == 2
    assert seen == [0, 0]                                   # replayed the identical batch
    assert rsm.diagnostics[-1][1] == "transient_error"
    assert rsm.should_checkpoint_and_exit() == (False, True, EXIT_CODE_FAILED_ON_RESULT_VALIDATION)


def test_persistent_and_correct_results():
    rsm = RerunStateMachine(RerunMode.VALIDATE_RESULTS)
    it = RerunDataIterator(iter(range(10)))
    _run_step(rsm, it, lambda b, r: 1e6)                    # reproducible but rejected value
    assert rsm.diagnostics[-1][1] == "persistent_error"
    assert rsm.should_checkpoint_and_exit() == (True, True, EXIT_CODE_RESUME_TO_DISAMBIGUATE)
    rsm2 = RerunStateMachine(RerunMode.VALIDATE_RESULTS)
    it2 = RerunDataIterator(iter(range(10)))
    assert _run_step(rsm2, it2, lambda b, r: 2.0) == 1 and not rsm2.diagnostics   # healthy step: one run


def test_error_injector_exercises_recovery():
    inj = RerunErrorInjector(error_injection_rate=1, error_injection_type="transient_error")
    rsm = RerunStateMachine(RerunMode.VALIDATE_RESULTS, inj)
    it = RerunDataIterator(iter(range(10)))
    assert _run_step(rsm, it, lambda b, r: 1.0) == 2
    assert rsm.diagnostics[-1][1] == "transient_error" and inj.injected == 1


def test_determinism_stats_mode():
    rsm = RerunStateMachine(RerunMode.REPORT_DETERMINISM_STATS)
    it = RerunDataIterator(iter(range(10)))
    assert _run_step(rsm, it, lambda b, r: 1.0 + 1e-3 * (r - 1)) == 2
    assert abs(rsm.stats["max_rel_diff"] - 1e-3) < 1e-9


def test_straggler_detector_single_rank():
    d = StragglerDetector()
    d.configure(1, 0, enabled=True)
    for _ in range(3):
        with d(bdata=True):
            pass
        with d:
            sum(range(1000))
    rep = d.report(total_flops=1e9, log_interval=3)
    assert rep.max_rank == 0 and rep.max_elapsed > 0


def test_yaml_arguments_and_config_logger(tmp_path):
    from types import SimpleNamespace

    from megatronapp_b200.config import TransformerConfig
    from megatronapp_b200.training.yaml_arguments import args_from_yaml, load_yaml
    from megatronapp_b200.utils.config_logger import log_config_to_disk

    y = tmp_path / "cfg.yaml"
    y.write_text("language_model:\n  num_layers: 4\n  hidden_size: 128\nmodel_parallel:\n  tensor_model_parallel_size: 2\nlr: 0.001\n")
    args = args_from_yaml(str(y), SimpleNamespace(num_layers=1, hidden_size=8, tensor_model_parallel_size=1, lr=0.1))
    assert (args.num_layers, args.hidden_size, args.tensor_model_parallel_size, args.lr) == (4, 128, 2, 0.001)
    assert load_yaml(str(y)).language_model.hidden_size == 128
    import pytest

    y.write_text("bogus_option: 3\n")
    with pytest.raises(ValueError):
        args_from_yaml(str(y), SimpleNamespace(num_layers=1))
    cfg = TransformerConfig(num_layers=1, hidden_size=8, num_attention_heads=2, config_logger_dir=str(tmp_path / "log"))
    fn = log_config_to_disk(cfg, {"config": cfg, "n": 3}, prefix="TransformerBlock", rank_str="0")
    import json

    assert json.load(open(fn))["config"]["hidden_size"] == 8


def test_file_autoresume_protocol(tmp_path):
    from megatronapp_b200.training import autoresume as ar

    sentinel = str(tmp_path / "stop")
    obj = ar.FileAutoResume(sentinel)
    ar.set_adlr_autoresume(obj)
    saved = []
    try:
        assert not ar.check_adlr_autoresume_termination(3, None, None, None, lambda *a: saved.append(a[0]))
        open(sentinel, "w").close()
        assert ar.check_adlr_autoresume_termination(4, None, None, None, lambda *a: saved.append(a[0]))
        assert saved == [4] and obj.resume_requested and (tmp_path / "stop.resume").exists()
        assert ar.FileAutoResume(None, deadline=1.0).termination_requested()
    finally:
        ar.set_adlr_autoresume(None)


def test_spiky_value_detection_and_disabled_mode_validation():
    """is_unexpectedly_large (reference rerun_state_machine.py:655): 10x the largest of the first samples is rejected; with reruns
    disabled a rejected fatal result raises, a non-fatal one only warns."""
    import pytest
    from functools import partial

    m = RerunStateMachine(mode="disabled")
    big = partial(m.is_unexpectedly_large, threshold=10, context="loss", num_samples=5)
    assert not any(big(v) for v in (2.0, 3.0, 2.5, 1.0, 2.0))          # sampling phase
    assert not big(29.0) and big(31.0) and not big(float("nan"))
    assert not m.is_unexpectedly_large(1e9, threshold=10, context="grads", num_samples=5)      # contexts are independent
    m.validate_result(31.0, big, "Spiky loss", fatal=False)
    assert m.stats["rejections"] == 1
    with pytest.raises(RuntimeError, match="Spiky loss"):
        m.validate_result(31.0, big, "Spiky loss", fatal=True)
    m.validate_result(2.0, big, "Spiky loss", fatal=True)


def test_rank_monitor_sections_and_learned_timeouts(tmp_path):
    """ft_integration (reference training/ft_integration.py): sections are timed, --calc-ft-timeouts turns the observed maxima
    into timeouts (x SAFETY_FACTOR), persists them next to the checkpoints and the next setup() loads them."""
    import json
    import time
    from types import SimpleNamespace

    from megatronapp_b200.training import ft_integration as ft

    args = SimpleNamespace(save=str(tmp_path), calc_ft_timeouts=True, async_save=False)
    ft.setup(args)
    cli = ft.get_rank_monitor_client()
    assert cli.section == "setup" and cli.section_timeouts["step"] == 600.0
    ft.on_checkpoint_loaded(is_local_chkpt=False)
    for _ in range(ft.NUM_WARMUP_ITERS + ft.MIN_ITERS_FOR_STEP_TIMEOUT_UPDATE):
        ft.on_training_step_start()
        time.sleep(0.01)
        ft.on_training_step_end()
    assert cli.section is None and 0.005 < cli.observed["step"] < 1.0
    ft.on_checkpointing_start()
    time.sleep(0.02)
    ft.on_checkpointing_end()
    state = json.load(open(tmp_path / "ft_state.json"))["section_timeouts"]
    assert state["step"] < 10 and state["checkpointing"] < 10 and state["setup"] < 100 and state["out_of_section"] == 3600.0
    ft.shutdown()                                                     # a complete run also learns the out-of-section timeout
    assert json.load(open(tmp_path / "ft_state.json"))["section_timeouts"]["out_of_section"] < 100
    assert ft.get_rank_monitor_client() is None
    ft.setup(SimpleNamespace(save=str(tmp_path), calc_ft_timeouts=False, async_save=True))
    assert ft.get_rank_monitor_client().section_timeouts["step"] == state["step"]
    ft.shutdown()


def test_hung_rank_is_terminated_by_the_watchdog(tmp_path):
    """--enable-ft-package: a rank that stops making progress (--simulated-fault rank_hung) overstays its section, the watchdog
    dumps the stacks and exits with the restart code instead of idling forever."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    save = tmp_path / "ck"
    save.mkdir()
    json.dump({"section_timeouts": {"setup": 120.0, "step": 60.0, "checkpointing": 60.0, "out_of_section": 3.0}},
              open(save / "ft_state.json", "w"))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29877", RANK="0", WORLD_SIZE="1", LOCAL_RANK="0",
               CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2")
    cmd = [sys.executable, os.path.join(root, "pretrain_gpt.py"), "--num-layers", "1", "--hidden-size", "16", "--num-attention-heads", "2",
           "--seq-length", "16", "--max-position-embeddings", "16", "--micro-batch-size", "1", "--global-batch-size", "1",
           "--train-iters", "100000", "--lr", "1e-3", "--vocab-size", "64", "--mock-data", "--log-interval", "1000", "--eval-interval",
           "100000", "--eval-iters", "0", "--distributed-backend", "gloo", "--no-gradient-accumulation-fusion", "--save", str(save),
           "--save-interval", "100000", "--enable-ft-package", "--simulated-fault", "rank_hung:0:1"]
    p = subprocess.run(cmd, cwd=root, env=env, capture_output=True, text=True, timeout=300)
    out = p.stdout + p.stderr
    assert p.returncode == 18, out[-2000:]
    assert "FT: loaded timeouts" in out and "'out_of_section' exceeded its timeout of 3.0 s" in out
    assert "Thread 0x" in out or "Current thread" in out            # faulthandler stack dump
