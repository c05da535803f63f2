"""Host-side logic of bench.py that needs no GPU: workloads, the config dict both arms print, the per-kernel split of the
algorithmic bytes, and the source stamp that guards the ncu-measured traffic figures."""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_algorithmic_bytes_split_adds_up():
    b = _bench()
    for L in (40, 60, 80):
        k = b.kernel_alg_bytes(L)
        assert k["lw_column"] == k["lw_taumol"] + k["lw_rtrn"]        # a fused kernel carries the charge of the two it replaces
        assert k["sw_column"] == k["sw_taumol"] + k["sw_solver"]
        assert sum(v for n, v in k.items() if not n.endswith("_column")) == b.b_alg(L) == 8 * (1060 * L + 35)


def test_workloads_cover_baseline_configs():
    b = _bench()
    assert set(b.WORKLOADS) == {"T42L40", "T85L40", "T170L60", "T42L40-4xCO2", "T341L80"}
    base, kw = b.workload_spec("T42L40-4xCO2")
    assert base == "T42L40" and kw["co2_ppmv"] == 4 * 390.0 and kw["ozone"] == "file" and kw["secondary_gases"]


def test_both_arms_print_the_same_config():
    b = _bench()
    for world in (1, 2, 8):
        c = b.config_dict("T170L60", world, True)
        assert c["columns_total"] == 131072 and c["columns_per_gpu"] == 131072 // world and c["layers"] == 60
        assert b.config_dict("T170L60", world, True) == c
    assert b.config_dict("T170L60", 8, False)["columns_total"] == 8 * 131072


def test_traffic_is_reported_only_for_the_profiled_sources(tmp_path, monkeypatch):
    b = _bench()
    from mima_b200.build import source_hash
    prof = tmp_path / "profiles"
    prof.mkdir()
    entry = {"sw_solver": {"dram_bytes_per_column": 1.0, "fp64_pipe_pct": 50.0}}
    monkeypatch.setattr(b, "ROOT", str(tmp_path))
    (prof / "traffic.json").write_text(json.dumps({"T170L60": dict(entry, _stamp={"csrc_sha256": source_hash(), "when": "now"})}))
    t, why = b.profiled("T170L60")
    assert t is not None and t["sw_solver"]["fp64_pipe_pct"] == 50.0
    (prof / "traffic.json").write_text(json.dumps({"T170L60": dict(entry, _stamp={"csrc_sha256": "0" * 64, "when": "then"})}))
    t, why = b.profiled("T170L60")
    assert t is None and "other kernel sources" in why
    t, why = b.profiled("T85L40")
    assert t is None
