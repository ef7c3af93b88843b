"""The bench.py output contract, checked on the committed evidence lines (profiles/) -- no GPU needed.  A line that loses a
key the driver / judge reads (roofline, cpu_baseline, e2e, clocks, gpu_launches) would otherwise only be noticed on the GPU box."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(path):
    return json.loads([l for l in open(path) if l.startswith("{")][-1])


def _latest(pattern):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)), key=os.path.getmtime)
    if not files:
        pytest.skip(f"no committed bench line matches {pattern}")
    return files[-1]


@pytest.mark.parametrize("pattern", ["r01_bench_c2_v*.json", "r01_bench_c3_v*.json", "r01_bench_c4_v*.json"])
def test_ours_line_has_the_contract_keys(pattern):
    d = _line(_latest(pattern))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["metric"] == "encode+decode images/sec" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic" and d["warmup"] >= 3
    assert "workload" in d["config"] and "model" not in d["config"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != d["value"]                          # measured separately, not a copy of the device-resident number
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert d["gpu_launches"] > 0 and d["clocks"]["sm_max_mhz"]
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert not (bad & set(d["clocks"]["reasons"]))


def test_reference_line_has_the_contract_keys():
    d = _line(_latest("r01_bench_ref_v*.json"))
    assert d["impl"] == "reference" and d["metric"] == "encode+decode images/sec" and d["unit"] == "images/s"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port" and d["gpu_launches"] == 0


def test_traffic_file_matches_the_capture_it_cites():
    t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["c2"]
    assert t["qkv_swiglu_gemm_dram_bytes_per_launch"] >= t["algorithmic_bytes_per_launch"] * 0.95
    assert t["qkv_swiglu_gemm_dram_bytes_per_launch"] <= t["algorithmic_bytes_per_launch"] * 1.25     # no wasted re-reads
    cited = t["source"].split(":")[0]
    assert os.path.exists(os.path.join(ROOT, cited)), cited


def test_flop_model_matches_the_survey_numbers():
    """bench.py's algorithmic FLOPs per image (verified there against FlopCounterMode): 350M @N=256
    189.01 GFLOP, 350M @N=1024 846.2, 5B-f16x64 @N=256 2592.6, @N=1024 10 795.5, 5B-f32x256 @N=1024 10 826.9."""
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "vitok-release_b200"))
    import bench
    from oracle import ae_oracle
    cases = [("Ld4-Ld24/1x16x64", 256, 189.01), ("Ld4-Ld24/1x16x64", 1024, 846.2), ("Td4-T/1x16x64", 256, 2592.6),
             ("Td4-T/1x16x64", 1024, 10795.5), ("Td4-T/1x32x256", 1024, 10826.9)]
    for variant, n, gflop in cases:
        got = bench.flops_per_image(ae_oracle.decode_variant(variant), n) / 1e9
        assert abs(got - gflop) / gflop < 1e-3, (variant, n, got, gflop)
    # valid-token accounting of the NaFlex workload: the sizes are seeded and every image fits the token budget
    sizes = bench.c3_sizes(64, 1234)
    assert len(sizes) == 64 and all(128 <= h <= 512 and 128 <= w <= 512 and -(-h // 16) * -(-w // 16) <= 1024 for h, w in sizes)
    assert sizes == bench.c3_sizes(64, 1234) and sizes != bench.c3_sizes(64, 1235)


def test_scaling_modes_and_shared_config():
    """c2 is BASELINE's "batch 64 ... batch-sharded to 2/4/8": strong scaling when N > 1 (64 / N images per rank), weak at N = 1 and
    for every other workload; --scaling forces one; both arms print the same `config` object."""
    import sys
    import types
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "vitok-release_b200"))
    import bench

    def args(**kw):
        d = dict(workload="c2", batch=0, scaling="auto", sw=0)
        d.update(kw)
        return types.SimpleNamespace(**d)

    assert bench.resolve_scaling(args(), 1) == ("weak", 64)
    assert [bench.resolve_scaling(args(), n) for n in (2, 4, 8)] == [("strong", 32), ("strong", 16), ("strong", 8)]
    assert bench.resolve_scaling(args(scaling="weak"), 8) == ("weak", 64)
    assert bench.resolve_scaling(args(scaling="strong"), 1) == ("strong", 64)
    assert bench.resolve_scaling(args(workload="c4"), 8) == ("weak", 8)
    assert bench.resolve_scaling(args(workload="c3"), 8) == ("weak", 64)
    assert bench.resolve_scaling(args(batch=8), 1) == ("weak", 8)
    with pytest.raises(SystemExit):
        bench.resolve_scaling(args(), 3)
    c = bench.make_config(args(), 8, "strong", 8)
    assert c["global_batch"] == 64 and c["batch_per_gpu"] == 8 and "sharded 8/GPU" in c["workload"] and "model" not in c
