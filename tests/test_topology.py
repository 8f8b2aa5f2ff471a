import pytest

from megatronapp_b200.parallel.topology import Topology, layer_offset


def test_default_order_tp_fastest():
    t = Topology(world_size=8, tp=2, pp=2)
    assert t.dp == 2
    assert t.groups("tp") == [[0, 1], [2, 3], [4, 5], [6, 7]]
    assert t.groups("dp") == [[0, 2], [1, 3], [4, 6], [5, 7]]
    assert t.groups("pp") == [[0, 4], [1, 5], [2, 6], [3, 7]]
    assert t.groups("mp") == [[0, 1, 4, 5], [2, 3, 6, 7]]
    assert t.coords(5) == {"tp": 1, "cp": 0, "ep": 0, "dp": 0, "pp": 1, "etp": 1, "edp": 0}


def test_every_rank_in_exactly_one_group_per_kind():
    t = Topology(world_size=16, tp=2, pp=2, cp=2, ep=2)
    for kind in ["tp", "pp", "dp", "cp", "dp_cp", "mp", "tp_cp", "ep", "etp", "edp", "ep_tp", "tp_ep_pp"]:
        seen = sorted(r for g in t.groups(kind) for r in g)
        assert seen == list(range(16)), kind


def test_expert_groups():
    t = Topology(world_size=8, tp=2, ep=4)
    assert t.edp == 1
    assert t.groups("ep") == [[0, 2, 4, 6], [1, 3, 5, 7]]
    assert t.groups("etp") == [[0, 1], [2, 3], [4, 5], [6, 7]]


def test_embedding_groups():
    t = Topology(world_size=8, pp=4)
    assert t.groups("embd") == [[0, 6], [1, 7]]
    assert t.groups("pos_embd") == [[0], [1]]


def test_fbd_pairs_and_half_dp():
    t = Topology(world_size=8, tp=1, pp=2, fbd=True)
    assert t.dp == 4 and t.effective_dp == 2
    assert t.groups("fb") == [[0, 1], [2, 3], [4, 5], [6, 7]]
    assert t.groups("half_dp") == [[0, 2], [1, 3], [4, 6], [5, 7]]
    assert t.is_forward_rank(0) and t.is_backward_rank(1) and t.twin_of(2) == 3 and t.twin_of(3) == 2


def test_invalid():
    with pytest.raises(ValueError):
        Topology(world_size=6, tp=4)
    with pytest.raises(ValueError):
        Topology(world_size=4, tp=2, fbd=True, pp=2)  # dp == 1 is odd
    with pytest.raises(ValueError):
        Topology(world_size=4, pp=1, vpp=2)


def test_layer_offsets_interleaved():
    # 16 layers, pp 4, vpp 2: stage p chunk v owns layers [(v*4+p)*2, +2)
    assert layer_offset(16, 4, 1, 2, 0) == (2, 2)
    assert layer_offset(16, 4, 1, 2, 1) == (10, 2)
    assert layer_offset(12, 3, 2) == (8, 4)
    assert layer_offset(10, 3, 0, first_stage_layers=2, last_stage_layers=2) == (0, 2)
    assert layer_offset(10, 3, 1, first_stage_layers=2, last_stage_layers=2) == (2, 6)
