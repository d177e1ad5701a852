"""Seeded scenarios shared by the golden-vector generator and the parity tests (shapes of
BASELINE.json configs C1-C5 at sizes the CPU oracle finishes in seconds)."""
SCENARIOS = {
    "c1_static": dict(cfg=dict(R=1, P=0, n_obj=4), seed=11, steps=3),
    "c2_eight_robots": dict(cfg=dict(R=8, P=0, n_obj=0), seed=12, steps=3, lo=4.0, hi=7.0),
    "c3_orca": dict(cfg=dict(R=1, P=20, scene="rvoscene", n_obj=4, max_ped=20), seed=13, steps=3),
    "c4_ervo_small": dict(cfg=dict(R=5, P=8, scene="ervoscene", n_obj=5, max_ped=8), seed=14, steps=3, beep=True, opt_in_beep=True,
                          lo=4.0, hi=7.0),
    "c5_sfm_small": dict(cfg=dict(R=3, P=5, scene="pedscene", n_obj=2), seed=15, steps=3, lo=3.5, hi=7.5),
}
