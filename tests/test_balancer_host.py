"""Host logic of the screen-partition feedback (streetunveiler_b200/sharded.py: _Balancer) -- pure Python, CPU."""
from streetunveiler_b200.sharded import _Balancer


def _simulate(cost_per_unit, steps=14, gain=0.7):
    """Ranks whose blend time is share_k * cost_per_unit[k] (a sparse range costs more per unit of the histogram than a
    saturated one).  Measurements arrive two steps late, like on the GPU."""
    G = len(cost_per_unit)
    b = _Balancer(G, gain=gain)
    history = []
    for _ in range(steps):
        step = b.begin_step()
        used = list(b.records[step]["shares"])
        times = [max(int(1e6 * s * c), 1) for s, c in zip(used, cost_per_unit)]
        history.append((step, times))
        if len(history) >= 3:                      # the measurement of step - 2 is applied now
            old_step, old_times = history[-3]
            b.update(old_step, old_times)
    final = [s * c for s, c in zip(b.shares, cost_per_unit)]
    return b, max(final) / (sum(final) / G)


def test_shares_converge_to_equal_times():
    b, spread = _simulate([3.0, 1.0, 0.6, 0.5, 0.5, 0.6, 1.0, 3.5])
    assert abs(sum(b.shares) - 1.0) < 1e-9 and min(b.shares) > 0
    assert spread < 1.05, spread                   # from 2.6x (equal shares) to within 5 % of the mean
    assert b.shares[0] < b.shares[3] and b.shares[7] < b.shares[4]


def test_missing_or_zero_measurements_change_nothing():
    b = _Balancer(4)
    s = b.begin_step()
    before = list(b.shares)
    b.update(s, [100, 0, 100, 100])                # one rank had nothing to report
    assert b.shares == before
    b.update(s + 7, [100, 200, 100, 100])          # a step nobody remembers
    assert b.shares == before
    assert b.report() == (s - 2, 0)                # nothing measured two steps ago
    b.update(s, [100, 400, 100, 100])
    assert b.shares[1] < before[1] and abs(sum(b.shares) - 1.0) < 1e-9


def test_floor_keeps_every_rank_alive():
    b = _Balancer(8, floor=0.02)
    for _ in range(30):
        s = b.begin_step()
        b.update(s, [10_000_000, 1, 1, 1, 1, 1, 1, 1])
    assert min(b.shares) >= 0.02 / 8 * 0.99
