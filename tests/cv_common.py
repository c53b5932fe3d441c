"""Conduction velocity of a planar wave on a strip, the measurement behind the reference's only
published result table, diff_conduction_velcoty.dat (CV in cm/s vs `diff`; the grid spacing is not
documented -- SURVEY.md Appendix B.1 finds dx = 0.0302 cm/cell for the 4v column)."""
import numpy as np

# diff_conduction_velcoty.dat:3-14 (values restated, the file is not copied)
CV_TABLE_4V = {0.4: 45.9, 0.5: 52.8, 0.6: 59.3, 0.7: 64.8, 0.8: 70.1, 0.9: 75.7, 1.0: 80.0,
               1.1: 84.4, 1.25: 90.9, 1.5: 101.0}
CV_TABLE_BR = {0.4: 30.2, 0.5: 33.8, 0.6: 37.6, 0.7: 41.2, 0.8: 44.7, 0.9: 47.7, 1.0: 50.9,
               1.1: 53.7, 1.25: 57.7, 1.5: 64.0, 1.75: 68.8, 2.0: 75.3}


def strip_config(diff, width, cheby=False, skip=False):
    return {'width': width, 'height': 5, 'dt': 0.1, 'dt_per_plot': 10, 'diff': diff, 'duration': 1,
            'timeline': False, 'timeline_name': 'x', 'save_graph': False, 'skip': skip,
            'cheby': cheby, 'ultra_slow': False}


def crossing_time(times, trace, level):
    t = np.asarray(trace, np.float64)
    i = int(np.argmax(t >= level))
    if i == 0 or t[i] < level:
        return None
    return times[i - 1] + (level - t[i - 1]) / (t[i] - t[i - 1]) * (times[i] - times[i - 1])


def measure_cv(model, level, c1, c2, dt_iter, max_iters, read_row):
    """Runs model.iterate() until the front passed column c2; CV in cells/ms from the
    linear-interpolated threshold-crossing times at columns c1 and c2 of the middle row.
    read_row(model) -> 1-D array of the transmembrane variable along the middle row."""
    times, a, b = [], [], []
    for i in range(max_iters):
        model.iterate()
        row = read_row(model)
        times.append((i + 1) * dt_iter)
        a.append(row[c1])
        b.append(row[c2])
        if row[c2] >= level and i > 2 and b[-2] >= level:
            break
    t1, t2 = crossing_time(times, a, level), crossing_time(times, b, level)
    assert t1 is not None and t2 is not None, 'the wave never reached the probes'
    return (c2 - c1) / (t2 - t1)


def fit_dx(cv_cells_per_ms, table):
    """Least-squares grid spacing (cm/cell) mapping cells/ms -> the table's cm/s, and the worst
    relative residual over the rows."""
    d = sorted(cv_cells_per_ms)
    x = np.array([cv_cells_per_ms[k] * 1000.0 for k in d])       # cells/s
    y = np.array([table[k] for k in d])
    dx = float((x * y).sum() / (x * x).sum())
    return dx, float(np.max(np.abs(x * dx - y) / y))
