import pytest

from megatronapp_b200.parallel.schedules import (Task, get_schedule_table, interleaved_warmup, schedule_1f1b,
                                                 schedule_interleaved, validate_interleaved)


def _check_complete(tasks, M, C):
    f = [(t.microbatch, t.chunk) for t in tasks if t.kind == "F"]
    b = [(t.microbatch, t.chunk) for t in tasks if t.kind == "B"]
    assert sorted(f) == sorted(b) == [(m, c) for m in range(M) for c in range(C)]
    done = set()
    for t in tasks:
        if t.kind == "F":
            done.add((t.microbatch, t.chunk))
        else:
            assert (t.microbatch, t.chunk) in done  # backward only after its forward


@pytest.mark.parametrize("pp,M", [(2, 4), (4, 8), (4, 3), (8, 16)])
def test_1f1b(pp, M):
    for r in range(pp):
        tasks = schedule_1f1b(pp, r, M)
        _check_complete(tasks, M, 1)
        warm = 0
        for t in tasks:
            if t.kind == "B":
                break
            warm += 1
        assert warm == min(pp - r - 1, M) + (1 if M > min(pp - r - 1, M) else 0)


def test_interleaved_matches_reference_trace_counts():
    # pp 4, C 4, M 8, G 4 -> rank 0 runs 18 warm-up forwards, rank 3 runs 12
    assert interleaved_warmup(4, 0, 8, 4, 4) == 18
    assert interleaved_warmup(4, 3, 8, 4, 4) == 12
    for r in range(4):
        tasks = schedule_interleaved(4, r, 8, 4, 4)
        _check_complete(tasks, 8, 4)
        lead = 0
        for t in tasks:
            if t.kind == "B":
                break
            lead += 1
        assert lead == interleaved_warmup(4, r, 8, 4, 4) + 1


def test_schedule_table_order():
    assert get_schedule_table(4, 2, 2) == [(0, 0), (1, 0), (0, 1), (1, 1), (2, 0), (3, 0), (2, 1), (3, 1)]


def test_interleaved_validation():
    with pytest.raises(ValueError):
        validate_interleaved(4, 8, 2, 3)   # G < pp
    with pytest.raises(ValueError):
        validate_interleaved(4, 10, 2, 8)  # 10 % 8 = 2 not in {0} U [4, 8)
    validate_interleaved(4, 12, 2, 8)      # 12 % 8 = 4 ok


def test_channel_order_is_consistent_between_neighbours():
    """Sender's F order on the last stage == receiver's F order on the first stage (chunk c -> c+1),
    which is what makes plain FIFO channels sufficient for the interleaved schedule."""
    pp, M, C, G = 4, 8, 3, 4
    last = [(t.microbatch, t.chunk + 1) for t in schedule_interleaved(pp, pp - 1, M, C, G)
            if t.kind == "F" and t.chunk < C - 1]
    first = [(t.microbatch, t.chunk) for t in schedule_interleaved(pp, 0, M, C, G) if t.kind == "F" and t.chunk > 0]
    assert last == first
    for r in range(pp - 1):
        a = [(t.microbatch, t.chunk) for t in schedule_interleaved(pp, r, M, C, G) if t.kind == "F"]
        b = [(t.microbatch, t.chunk) for t in schedule_interleaved(pp, r + 1, M, C, G) if t.kind == "F"]
        assert a == b
