"""Timer (mineral_b200/common/timer.py) against the reference's metric keys and semantics (mineral/common/timer.py:8-85)."""
import time

from mineral_b200.common.timer import Timer


class _Obj:
    def work(self, dt):
        time.sleep(dt)
        return dt


def test_timer_keys_and_counts_match_reference_semantics():
    t = Timer()
    o = _Obj()
    t.wrap("agent", o, ["work"])
    assert o.work(0.01) == 0.01 and o.work(0.02) == 0.02
    with t.scope("env.step"):
        time.sleep(0.005)
    m = t.stats(step=100, total_names=("agent.work",), reset=False)
    for col in Timer.COLUMNS:
        assert f"agent.work/{col}" in m, col
    assert m["agent.work/count"] == 2 and m["env.step/count"] == 1
    assert 0.03 <= m["agent.work/sum"] < 0.2 and m["agent.work/min"] <= m["agent.work/avg"] <= m["agent.work/max"]
    assert abs(m["agent.work/totalrate"] - 100 / m["agent.work/total"]) < 1e-9
    assert m["total"] == m["agent.work/total"] and "totalrate" in m and "lastrate" in m and "duration" in m
    # reset=False keeps the scopes' durations (timer.py:81-84: only the running totals are cleared)
    m2 = t.stats(step=200)
    assert m2["agent.work/count"] == 2
    m3 = t.stats(step=300)
    assert "agent.work/avg" not in m3 and m3["agent.work/count"] == 0
    t.start("x")
    t.end("x")
    assert t.stats()["x/count"] == 1
