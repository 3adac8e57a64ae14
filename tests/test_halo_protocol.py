"""conv_halo.cu's mbarrier protocol, checked by randomised simulation on the CPU
(tools/halo_protocol_sim.py): no deadlock, no over-arrival, every read sees the
chunk / tap it expects -- for every ring configuration the planner may choose."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import halo_protocol_sim as sim  # noqa: E402


def _shipped_sets():
    """Split sets per CTA of the compiled kernel variants:
    conv_halo_kernel<SETS, MIN_CTAS> instantiations in conv_halo.cu."""
    src = open(os.path.join(ROOT, 'preworld_b200', 'csrc', 'conv_halo.cu')).read()
    sets = {int(m) for m in re.findall(
        r'cudaLaunchKernelEx\(&cfg, conv_halo_kernel<(\d+), \d+>', src)}
    assert sets, 'no kernel launches found'
    return tuple(sorted(sets))


def test_protocol_is_safe_for_every_planned_ring():
    cases, bad = sim.sweep(_shipped_sets(), n=3, seed=1)
    assert cases > 500
    assert not bad, bad


def test_protocol_is_safe_for_persistent_ctas():
    """Three tiles per CTA: rings run on across tiles, sets whose stride
    overshoots a tile start the next one mid-way, the MMA warp waits for the
    epilogue (acc_empty) before it overwrites the accumulators."""
    cases, bad = sim.sweep(_shipped_sets(), n=2, seed=5, tiles=3)
    assert cases > 500
    assert not bad, bad


def test_protocol_is_safe_with_more_split_sets():
    _, bad = sim.sweep((3, 4), n=2, seed=2)
    assert not bad, bad


def test_sim_catches_early_release_of_a_skipped_chunk():
    # the rule the split warps follow (release a chunk they never read only
    # after its halo_full) is load-bearing: without it pointwise layers race
    _, bad = sim.sweep((2,), n=15, seed=3, fixed=False)
    assert bad and all(k[2] == 1 and k[1] == 1 for k in bad), bad


def test_sim_catches_a_ring_shallower_than_the_set_stride():
    _, bad = sim.sweep((3,), n=5, seed=4, only_allowed=False)
    assert bad and all(k[1] == 1 and k[5] == 2 for k in bad), bad
