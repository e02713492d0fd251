"""CPU check of the KERNEL SOURCE: csrc/pam_track.h compiled for the host with one "thread"
(tests/hostemu, test infrastructure only) must reproduce the oracle's decisions.  This is how the
algorithmic logic of the CUDA path is iterated on in the GPU-less build container; the real parity
tests are the -m gpu ones."""
import pytest

from tests import util
from pam_b200 import synth

CASES = [
    ("shelf", {}, 120),
    ("shelf", dict(enter_stagger=25, miss_prob=0.1, outlier_prob=0.05, absences=[(1, 60, 90)]), 200),
    ("campus", dict(miss_prob=0.05, outlier_prob=0.03), 150),
    ("shelf17", dict(miss_prob=0.03, outlier_prob=0.02), 80),
    ("panoptic", dict(miss_prob=0.05, outlier_prob=0.03), 60),
]


@pytest.mark.parametrize("shape,kw,T", CASES)
def test_kernel_source_on_host_matches_oracle(shape, kw, T):
    st = synth.make_stream(shape, 7, T, **kw)
    cfg = util.stream_config(st, max_tracks=12)
    out = util.run_hostemu([st], cfg)
    assert out["status"].tolist() == [0]
    oo, oa, _ = util.run_oracle(st)
    worst = util.compare_with_oracle(out, 0, st, oo, oa)
    assert out["count"].sum() > 0 and worst < 5e-4
