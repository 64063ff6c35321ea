"""Host-side logic that needs no GPU: planning of the pipelined host step."""


def test_host_pipeline_plan():
    """chunk / slice plan of the pipelined host step (player_util.host_pipeline_plan): pure host logic"""
    from active_tracking_rl_b200.player_util import host_pipeline_plan as plan
    f32 = lambda E: E * 2 * 169 * 4  # noqa: E731
    assert plan(65536, f32(65536)) == (16, {8, 16})            # the bench size: 16 chunks of 5.5 MB, two halves
    assert plan(65536, f32(65536) // 4) == (10, {5, 10})       # uint8 observations
    assert plan(4096, f32(4096)) == (2, {2})                   # small batch: one forward over everything
    assert plan(1003, f32(1003)) == (1, {1})
    assert plan(65536, f32(65536), chunks=8, forward_slices=4) == (8, {2, 4, 6, 8})
    assert plan(65536, f32(65536), chunks=16, forward_slices=(8, 12)) == (16, {8, 12, 16})
    assert plan(65536, f32(65536), chunks=3, forward_slices=8) == (3, {1, 2, 3})
    for bad in (dict(chunks=0), dict(chunks=17), dict(forward_slices=(0,)), dict(chunks=4, forward_slices=(5,))):
        try:
            plan(65536, f32(65536), **bad)
        except ValueError:
            continue
        raise AssertionError(bad)
