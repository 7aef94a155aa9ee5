"""baseline/<config>.counters.json (the algorithmic work both `roofline.achieved` and the
judge use) were written from the GPU's reference-order counting kernel. Config A is small enough
for the CPU oracle in its pruned mode to recount it here in full, every counter; config B's first
two samples of a strip of the frame are checked the same way on the GPU box by the parity tests."""
import json
import os

from conftest import ASSETS, ROOT


def test_config_a_counters_equal_the_cpu_oracle(rt, orc, default_scene, env_2k):
    with open(os.path.join(ROOT, "baseline", "A.counters.json")) as fh:
        want = json.load(fh)["counters_per_step"]
    ora = orc.Oracle(default_scene, env_2k, 640, 360, max_bounces=8, prune=True)
    _, got = ora.render(0, 64)
    assert {k: got[k] for k in want} == want


def test_baseline_counters_cover_the_configs_and_follow_the_rule():
    for name in "ABCD":
        with open(os.path.join(ROOT, "baseline", f"{name}.counters.json")) as fh:
            d = json.load(fh)
        c = d["counters_per_step"]
        ops = (18 * c["nodes_tested"] + 45 * c["tris_tested"] + 30 * c["spheres_tested"] + 40 * c["planes_tested"] +
               250 * c["shaded_bounces"])
        assert ops == d["algorithmic_ops_per_step"], name
        assert c["rays"] > 0 and d["algorithmic_bytes_per_step"] > 0
