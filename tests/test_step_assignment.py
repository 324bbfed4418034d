"""Step assignment: bit-exact against the reference's own outputs (tests/golden/step_assignment.json,
generated from /root/reference/src/pipeline/step_assignment.py) plus the uneven extension."""
import json
import os

import pytest

from vdpp_b200.pipeline import StepRange, assign_steps, assign_steps_uneven, stage_sizes

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "step_assignment.json")))


@pytest.mark.parametrize("key", sorted(GOLD["table"]))
def test_matches_reference_table(key):
    T, W = map(int, key.split("x"))
    for rank, want in enumerate(GOLD["table"][key]):
        if want == "ValueError":
            with pytest.raises(ValueError):
                assign_steps(T, W, rank)
        else:
            got = assign_steps(total_steps=T, world_size=W, rank=rank)
            assert [got.start, got.end] == want


@pytest.mark.parametrize("name,args", [("zero_steps", (0, 1, 0)), ("neg_steps", (-1, 1, 0)),
                                       ("zero_world", (28, 0, 0)), ("rank_hi", (28, 4, 4)),
                                       ("rank_neg", (28, 4, -1))])
def test_argument_errors_match_reference(name, args):
    assert GOLD["errors"][name] == "ValueError"
    with pytest.raises(ValueError):
        assign_steps(*args)
    with pytest.raises(ValueError):
        assign_steps_uneven(*args)


def test_step_range_semantics():
    sr = StepRange(start=5, end=10)
    assert sr.count == 5 and list(sr) == [5, 6, 7, 8, 9]
    assert StepRange(3, 3).count == 0
    for bad in ((-1, 5), (10, 5), (0, -1)):
        with pytest.raises(ValueError):
            StepRange(*bad)
    with pytest.raises(Exception):
        sr.start = 1  # frozen


@pytest.mark.parametrize("T,W", [(25, 2), (25, 4), (25, 7), (25, 8), (35, 8), (35, 7), (28, 4), (3, 5)])
def test_uneven_split_covers_schedule(T, W):
    ranges = [assign_steps_uneven(T, W, r) for r in range(W)]
    sizes = [r.count for r in ranges]
    assert sizes == stage_sizes(T, W)
    assert sum(sizes) == T and max(sizes) - min(sizes) <= 1
    assert sizes == sorted(sizes, reverse=True)          # the longer stages come first
    pos = 0
    for r in ranges:
        assert r.start == pos
        pos = r.end
    if T % W == 0:
        assert ranges == [assign_steps(T, W, r) for r in range(W)]


def test_baseline_config_sizes():
    assert stage_sizes(25, 4) == [7, 6, 6, 6]
    assert stage_sizes(25, 8) == [4, 3, 3, 3, 3, 3, 3, 3]
    assert stage_sizes(35, 8) == [5, 5, 5, 4, 4, 4, 4, 4]
    assert stage_sizes(25, 7) == [4, 4, 4, 4, 3, 3, 3]
