"""bench.py's JSON contract without a GPU: the whole N=1 arm (device-resident steps, e2e through the host-pointer C-ABI calls, per-kernel
profile spans, roofline, CPU baseline) runs against the host emulation of the kernels with a workload shrunk to seconds, and the
`--impl reference` arm runs as it is.  Only the shape of the lines is checked here -- the numbers of an emulated run mean nothing."""
import json
import os
import sys

import pytest

from test_emu_parity import EMU_SO, build_emu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(monkeypatch, capsys, argv):
    import bench
    monkeypatch.setattr(bench, "DBG_BITS", 1 << 26)
    monkeypatch.setattr(bench, "CBF_BYTES", 1 << 23)
    monkeypatch.setattr(bench, "host_filter_sizes", lambda: (1 << 26, 1 << 23, 0))
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    bench.main()
    lines = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line"
    return json.loads(lines[0])


@pytest.mark.parametrize("engine", ["sliced", "direct"])
def test_bench_line_has_the_contract_keys(monkeypatch, capsys, engine):
    from rnabloom_b200 import binding as B
    build_emu()
    monkeypatch.setattr(B, "_lib", B.bind(EMU_SO, allow_missing=True))
    monkeypatch.setenv("RB_SLICE_BITS_LOG2", "17")
    monkeypatch.setenv("RB_SLICE_BYTES_LOG2", "15")
    d = _run(monkeypatch, capsys, ["--engine", engine, "--reads-per-step", "600", "--steps", "2", "--warmup", "3"])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "k-mers/s" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"] and d["config"]["engine"] == engine
    assert d["gpu_launches"] >= (2 * 2 if engine == "direct" else 2 * 14)
    e = d["e2e"]
    assert e["unit"] == "k-mers/s" and e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] == 600 * 126 * 4
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "kernels_ms_per_step"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert any(n.startswith("ks_" if engine == "sliced" else "k_graph_") for n in r["kernels_ms_per_step"])
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == "k-mers/s" and "sample" in c


def test_reference_arm_line(monkeypatch, capsys):
    d = _run(monkeypatch, capsys, ["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "k-mers/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
