"""Host-side logic of the fused trainer that needs no device."""
import pytest


def test_plan_gen_groups():
    """trainer.plan_gen_groups: how the generator forwards of the n_critic iterations (train.py:193-196) are grouped
    into batched passes."""
    from music2dance_b200.trainer import GEN_GROUP_SEQUENCES, plan_gen_groups
    assert GEN_GROUP_SEQUENCES == 64
    assert plan_gen_groups(8, 7) == {0: 8}                          # default.yaml: 56 sequences, one pass
    assert plan_gen_groups(8, 16) == {0: 4, 4: 4}
    assert plan_gen_groups(8, 30) == {0: 2, 2: 2, 4: 2, 6: 2}
    assert plan_gen_groups(8, 33) == {i: 1 for i in range(8)}       # one pass per iteration from 33 sequences on
    assert plan_gen_groups(8, 512) == {i: 1 for i in range(8)}
    assert plan_gen_groups(3, 2) == {0: 3}
    assert plan_gen_groups(5, 20) == {0: 3, 3: 2}
    assert plan_gen_groups(8, 7, "1,7") == {0: 1, 1: 7}
    assert plan_gen_groups(8, 7, "1,1,2,4") == {0: 1, 1: 1, 2: 2, 4: 4}
    assert plan_gen_groups(2, 512, "2") == {0: 2}                   # an explicit plan overrides the size rule
    for bad in ("1,6", "9", "0,8", "4,4,1"):
        with pytest.raises(ValueError):
            plan_gen_groups(8, 7, bad)
    # every iteration is served exactly once, in order
    for nc in range(1, 10):
        for B in (1, 2, 7, 9, 31, 64):
            plan = plan_gen_groups(nc, B)
            served = [i + j for i, g in sorted(plan.items()) for j in range(g)]
            assert served == list(range(nc))
            assert all(g * B <= max(GEN_GROUP_SEQUENCES, B) for g in plan.values())
