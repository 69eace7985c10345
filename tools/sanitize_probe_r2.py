"""Small end-to-end runs of the round-2 device paths for `compute-sanitizer` (memcheck / racecheck / synccheck):

  a  mixed-site 128-member blocks (two forcing streams per block) with events, segmented, fast + throughput policy
  b  the validation dump with the mass-balance rows and the message counters
  d  the one-GPU row summaries of a large row (key histograms + candidates sorted in shared memory) and of small rows
  c  the persistent grid (more block descriptors than resident CTA slots: items pulled from the atomic counter,
     per-block progress words) followed by the team summaries (lockstep radix select, one-rank team)

    compute-sanitizer --tool memcheck python tools/sanitize_probe_r2.py [a|b|c ...]

The sections print a checksum each, so a sanitizer run can also be compared with a plain run.
"""
import hashlib
import sys

sys.path.insert(0, '.')
import numpy as np
from sipnet_b200 import _abi as A, api, synth


def digest(*arrays):
    h = hashlib.sha1()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:12]


def section_a():
    sites, P, ms, flags = synth.config_c3(nsites=5, members_per_site=45, nyears=1)  # 225 members: two mixed blocks
    for math in (A.MATH_FAST, A.MATH_THROUGHPUT):
        with api.Ensemble(sites, P, ms, flags, outputs=A.OUT_FULL | A.OUT_EVENTS | A.OUT_MOMENTS, math=math,
                          summary_cols=[A.O["nee"]], max_event_records=256, block_threads=128) as ens:
            ens.run(0, 300)
            ens.run(300, 730)
            print("a", math, digest(ens.output(), ens.mean(), ens.status(), ens.event_counts()), flush=True)


def section_b():
    site = synth.synth_site(0, 1, "half-daily", with_events=True)
    P = synth.synth_params(70, stream=5)
    with api.Ensemble([site], P, None, synth.SYNTH_FLAGS, outputs=A.OUT_FULL | A.OUT_DEBUG | A.OUT_EVENTS,
                      math=A.MATH_VALIDATION, max_event_records=256) as ens:
        ens.run(0, 200)
        ens.run(200, 730)
        print("b", digest(ens.output(), ens.debug(), ens.balance(), ens.counters()), flush=True)


def section_c():
    # 148 SMs x 2 resident 128-member blocks = 296 slots: 300 block descriptors take the persistent grid, 600 steps
    # = three 256-step items per block (the second and third wait on the progress word of their predecessor)
    M = 300 * 128
    site = synth.synth_site(0, 1, "half-daily")
    P = synth.synth_params(M, stream=9)
    with api.Ensemble([site], P, None, synth.SYNTH_FLAGS, outputs=A.OUT_MOMENTS | A.OUT_QUANTILES, math=A.MATH_FAST,
                      summary_cols=[A.O["nee"], A.O["gpp"]], quantiles=[0.0, 0.05, 0.5, 0.95, 1.0], block_threads=128) as ens:  # 5 quantiles: two select groups
        ens.join_team(1, 0)
        ens.run(0, 600)
        ens.team_summaries()
        print("c", digest(ens.mean(), ens.variance(), ens.quantiles(), ens.status()), "levels", ens.team_last_levels(),
              flush=True)


def section_d():
    site = synth.synth_site(2, 1, "half-daily", with_events=True)
    P = synth.synth_params(3000, stream=13)
    P[:, 1500:] = P[:, :1500]  # every member twice: ties in every row
    with api.Ensemble([site], P, None, synth.SYNTH_FLAGS, outputs=A.OUT_MOMENTS | A.OUT_QUANTILES, math=A.MATH_FAST,
                      summary_cols=[A.O["nee"], A.O["gpp"]], quantiles=[0.0, 0.05, 0.5, 0.95, 1.0]) as ens:
        ens.run(0, 400)
        print("d", digest(ens.mean(), ens.variance(), ens.quantiles()), flush=True)


if __name__ == "__main__":
    want = sys.argv[1:] or ["a", "b", "c", "d"]
    for s in want:
        {"a": section_a, "b": section_b, "c": section_c, "d": section_d}[s]()
    print("done")
