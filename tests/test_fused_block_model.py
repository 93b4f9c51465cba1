"""CPU checks of the experimental fused Upsample block (csrc/tc_block.cu): its window / slot / replicate-padding
indexing against a plain layer-by-layer evaluation, and its mbarrier protocol under randomised schedules.  The kernel
itself is exercised on the GPU by tools/fused_block_check.py (it is off by default)."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_window_indexing_matches_layerwise_block():
    assert _load("fused_block_model").main() == 0


def test_wrong_halo_is_detected():
    m = _load("fused_block_model")
    m.HALO = 30
    m.BS = m.BW - 2 * m.HALO
    assert m.main() == 1


def test_protocol_has_no_deadlock_or_hazard(monkeypatch):
    m = _load("fused_block_protocol_sim")
    monkeypatch.setattr(sys, "argv", ["sim", "60", "4"])
    assert m.main() == 0


def test_protocol_sim_catches_phase_aliasing(monkeypatch):
    m = _load("fused_block_protocol_sim")
    m.BUG = "naive_wait"
    monkeypatch.setattr(sys, "argv", ["sim", "30", "4"])
    assert m.main() == 1


def test_model_constants_match_the_kernel():
    import re

    src = open(os.path.join(ROOT, "tinyvc_b200", "csrc", "tc_block.cu")).read()

    def const(name):
        m = re.search(r"constexpr\s+(?:int|uint32_t)\s+" + name + r"\s*=\s*([^;]+);", src)
        assert m, name
        return m.group(1).split("//")[0].strip()

    m = _load("fused_block_model")
    assert int(const("kBM")) == m.BM and int(const("kBT")) == m.BT and int(const("kBHalo")) == m.HALO
    assert int(const("kBPad")) == m.PAD and const("kBSlots") == "kBW + 72" and m.SLOTS == m.BW + 72
    assert sum(m.DIL) == m.HALO and max(m.DIL) == int(const("kBMaxDil"))
    # shared-memory budget of the kernel's layout (bytes), recomputed here
    act = 2 * 3 * m.SLOTS * 16
    total = 2 * act + 2 * (2 * 3 * 17 * 128) + (4 * 3 * 4096 + 2 * 8192 + 4096) + 256
    assert total == 208128 and total <= 227 * 1024
